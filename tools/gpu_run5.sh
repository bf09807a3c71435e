#!/bin/bash
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_sampler_v2_gpu.py -m gpu -q 2>&1 | tail -5) > gpurun_out/r02_pytest5.log 2>&1
tail -2 gpurun_out/r02_pytest5.log
timeout 300 python tools/sampler_sweep.py 20 > gpurun_out/r02_sampler_sweep5.log 2>&1
grep "'v2'\|'generic', \|'exact'" gpurun_out/r02_sampler_sweep5.log | cut -c1-160
timeout 900 python tools/decode_bench.py > gpurun_out/r02_decode_bench.json 2> gpurun_out/r02_decode_bench.err
tail -c 600 gpurun_out/r02_decode_bench.err; cat gpurun_out/r02_decode_bench.json
