#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_attn_gpu.py tests/test_generate_gpu.py tests/test_llama_gpu.py -m gpu -q 2>&1 | tail -8) > gpurun_out/r02_pytest14.log 2>&1
tail -4 gpurun_out/r02_pytest14.log | cut -c1-200
{ echo staged; python tools/decode_attn_ab.py 2>&1 | tail -2; echo batched; MMFS_DECODE_ATTN=batched python tools/decode_attn_ab.py 2>&1 | tail -2; } > gpurun_out/r02_decode_attn_ab.log 2>&1
cat gpurun_out/r02_decode_attn_ab.log
ncu --set full --clock-control none -k regex:attn_decode_staged128 -s 8 -c 1 --page details python tools/decode_attn_ab.py > gpurun_out/r02_decode_attn_staged_ncu.txt 2>&1
grep -E "Duration|DRAM Throughput|Achieved Occupancy|Executed Ipc Active|No Eligible|Warp Cycles Per Issued|Registers Per" gpurun_out/r02_decode_attn_staged_ncu.txt | head
timeout 900 python tools/decode_bench.py > gpurun_out/r02_decode_bench5.json 2> gpurun_out/r02_decode_bench5.err
tail -c 300 gpurun_out/r02_decode_bench5.err; cat gpurun_out/r02_decode_bench5.json
