#!/bin/bash
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests/ -m gpu -q 2>&1 | tail -15) > gpurun_out/r02_pytest13.log 2>&1
tail -6 gpurun_out/r02_pytest13.log | cut -c1-200
timeout 900 python tools/decode_bench.py > gpurun_out/r02_decode_bench4.json 2> gpurun_out/r02_decode_bench4.err
tail -c 300 gpurun_out/r02_decode_bench4.err; cat gpurun_out/r02_decode_bench4.json
