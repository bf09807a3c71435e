#!/bin/bash
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_msda_bwd_gpu.py tests/test_capi.py tests/test_msda_gpu.py -m gpu -q 2>&1 | tail -15) > gpurun_out/r02_pytest7.log 2>&1
tail -4 gpurun_out/r02_pytest7.log
timeout 300 python tools/bwd_ab.py > gpurun_out/r02_bwd_ab.log 2>&1; tail -3 gpurun_out/r02_bwd_ab.log
