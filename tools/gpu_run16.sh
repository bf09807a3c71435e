#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_llama_gpu.py tests/test_generate_gpu.py tests/test_mm_interleaved_gpu.py tests/test_unet_gpu.py -m gpu -q 2>&1 | tail -8) > gpurun_out/r02_pytest16.log 2>&1
tail -3 gpurun_out/r02_pytest16.log | cut -c1-200
timeout 900 python tools/decode_bench.py > gpurun_out/r02_decode_bench6.json 2> gpurun_out/r02_decode_bench6.err
tail -c 300 gpurun_out/r02_decode_bench6.err; cat gpurun_out/r02_decode_bench6.json
