#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_linear_skinny_gpu.py tests/test_attn_gpu.py tests/test_generate_gpu.py tests/test_llama_gpu.py -m gpu -q -x 2>&1 | tail -25) > gpurun_out/r02_pytest10.log 2>&1
tail -25 gpurun_out/r02_pytest10.log | cut -c1-220
timeout 300 python tools/skinny_ab.py > gpurun_out/r02_skinny_ab.log 2>&1; tail -8 gpurun_out/r02_skinny_ab.log
