#!/bin/bash
mkdir -p gpurun_out
{
python tools/skinny_debug.py 27648 5120; python tools/skinny_debug.py 5120 13824; python tools/skinny_debug.py 5120 5120
timeout 300 python -m pytest tests/test_linear_skinny_gpu.py -m gpu -q 2>&1 | tail -12 | cut -c1-200
for kc in 0 1280 512; do KC_MAX=$kc timeout 300 python tools/skinny_ab.py 2>&1 | tail -6; done
} > gpurun_out/r02_skinny11.log 2>&1
cat gpurun_out/r02_skinny11.log
