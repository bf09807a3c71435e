#!/bin/bash
mkdir -p gpurun_out
{
python tools/skinny_debug.py 27648 5120 1; python tools/skinny_debug.py 15360 5120 1; python tools/skinny_debug.py 5120 13824
timeout 300 python -m pytest tests/test_linear_skinny_gpu.py -m gpu -q 2>&1 | tail -12 | cut -c1-200
timeout 300 python tools/skinny_ab.py 2>&1 | tail -6
} > gpurun_out/r02_skinny12.log 2>&1
cat gpurun_out/r02_skinny12.log
ncu --set full --clock-control none -k regex:linear_skinny -s 14 -c 1 --page details python tools/skinny_ab.py > gpurun_out/r02_skinny_ncu.txt 2>&1
grep -E "Duration|DRAM Throughput|Executed Ipc Active|No Eligible|Warp Cycles Per Issued|Registers Per|Issue Slots Busy|L2 Hit|Grid Size|Dynamic Shared" gpurun_out/r02_skinny_ncu.txt | head -14
