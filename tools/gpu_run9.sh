#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_attn_gpu.py tests/test_generate_gpu.py tests/test_llama_gpu.py tests/test_mm_interleaved_gpu.py -m gpu -q 2>&1 | tail -15) > gpurun_out/r02_pytest9.log 2>&1
tail -4 gpurun_out/r02_pytest9.log
for s in llama_cfg3 llama_nc llama_cfg2 sd_b16; do timeout 120 python tools/attn_one.py $s 9; done > gpurun_out/r02_attn_one9.log 2>&1
cat gpurun_out/r02_attn_one9.log
timeout 900 python tools/decode_bench.py > gpurun_out/r02_decode_bench3.json 2> gpurun_out/r02_decode_bench3.err
tail -c 300 gpurun_out/r02_decode_bench3.err; cat gpurun_out/r02_decode_bench3.json
timeout 300 python tools/decode_profile.py 2>&1 | grep -v Warn | head -14 > gpurun_out/r02_decode_profile2.txt; cat gpurun_out/r02_decode_profile2.txt | cut -c1-150
