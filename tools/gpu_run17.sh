#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_attn_gpu.py tests/test_llama_gpu.py tests/test_visual_tokenizer_gpu.py tests/test_unet_gpu.py -m gpu -q 2>&1 | tail -8) > gpurun_out/r02_pytest17.log 2>&1
tail -3 gpurun_out/r02_pytest17.log | cut -c1-200
for s in llama_cfg3 llama_nc llama_cfg2 sd_b16 clip; do timeout 120 python tools/attn_one.py $s 9; done > gpurun_out/r02_attn_one17.log 2>&1
cat gpurun_out/r02_attn_one17.log
