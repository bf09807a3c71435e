#!/bin/bash
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_sampler_v2_gpu.py tests/test_mmfs_gpu.py tests/test_generate_gpu.py tests/test_llama_gpu.py tests/test_cache_safety.py -m gpu -q 2>&1 | tail -30) > gpurun_out/r02_pytest4.log 2>&1
tail -3 gpurun_out/r02_pytest4.log
timeout 300 python tools/sampler_sweep.py 20 > gpurun_out/r02_sampler_sweep4.log 2>&1
grep "'v2'\|'generic', \|'exact'" gpurun_out/r02_sampler_sweep4.log | cut -c1-160
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:mmfs_sampler_v2 -c 1 -f -o /tmp/r02_sampler_v2b python tools/sampler_one.py 4 2 masked v2 0 3 > gpurun_out/ncu_v2b.log 2>&1
ncu -i /tmp/r02_sampler_v2b.ncu-rep --page details > gpurun_out/r02_sampler_v2b_ncu_details.txt 2>&1
ncu -i /tmp/r02_sampler_v2b.ncu-rep --page source --csv > gpurun_out/r02_sampler_v2b_ncu_source.csv 2>&1
timeout 900 python tools/decode_bench.py > gpurun_out/r02_decode_bench.json 2> gpurun_out/r02_decode_bench.err
tail -c 1500 gpurun_out/r02_decode_bench.err; cat gpurun_out/r02_decode_bench.json
