#!/bin/bash
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_attn_gpu.py tests/test_generate_gpu.py tests/test_llama_gpu.py tests/test_mm_interleaved_gpu.py -m gpu -q 2>&1 | tail -15) > gpurun_out/r02_pytest8.log 2>&1
tail -4 gpurun_out/r02_pytest8.log
timeout 900 python tools/decode_bench.py > gpurun_out/r02_decode_bench2.json 2> gpurun_out/r02_decode_bench2.err
tail -c 400 gpurun_out/r02_decode_bench2.err; cat gpurun_out/r02_decode_bench2.json
