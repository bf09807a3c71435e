#!/bin/bash
mkdir -p gpurun_out
{
timeout 120 python tools/skinny_debug.py 27648 5120 1; timeout 120 python tools/skinny_debug.py 5120 13824; timeout 120 python tools/skinny_debug.py 5120 5120
timeout 300 python -m pytest tests/test_linear_skinny_gpu.py -m gpu -q 2>&1 | tail -12 | cut -c1-200
SK_MODE=2 timeout 300 python tools/skinny_ab.py 2>&1 | tail -7
SK_MODE=1 timeout 300 python tools/skinny_ab.py 2>&1 | tail -7
} > gpurun_out/r02_skinny21.log 2>&1
cat gpurun_out/r02_skinny21.log | cut -c1-200
