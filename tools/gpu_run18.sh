#!/bin/bash
mkdir -p gpurun_out
(timeout 180 python -m pytest tests/test_attn_gpu.py -m gpu -q -x -k "tc_attention" 2>&1 | tail -15) > gpurun_out/r02_pytest18.log 2>&1
tail -15 gpurun_out/r02_pytest18.log | cut -c1-200
if grep -q "passed" gpurun_out/r02_pytest18.log && ! grep -q "failed\|Timeout\|error" gpurun_out/r02_pytest18.log; then
  (timeout 120 python tools/attn_sweep.py 2>&1 | tail -10; MMFS_ATTN_PERSISTENT=0 timeout 60 python tools/attn_one.py llama_cfg3 9; timeout 60 python tools/attn_one.py llama_cfg3 9; timeout 60 python tools/attn_one.py sd_b16 9; timeout 60 python tools/attn_one.py llama_nc 9) > gpurun_out/r02_attn_persistent.log 2>&1
  cat gpurun_out/r02_attn_persistent.log
fi
