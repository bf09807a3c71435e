#!/bin/bash
mkdir -p gpurun_out
CS="compute-sanitizer --tool memcheck --error-exitcode 9 --launch-timeout 0"
{
echo "== skinny linear"; timeout 600 $CS python -m pytest tests/test_linear_skinny_gpu.py -m gpu -q -x -k "matches_fp64 and (4- or 5- or 6-) or rejects" 2>&1 | tail -6
echo "== rope append + decode attention"; timeout 600 $CS python -m pytest tests/test_llama_gpu.py tests/test_attn_gpu.py -m gpu -q -x -k "rope_append or decode_attention" 2>&1 | tail -6
echo "== persistent attention (small)"; timeout 900 $CS python -m pytest tests/test_attn_gpu.py -m gpu -q -x -k "tc_attention and (14- or 4- or 0-)" 2>&1 | tail -6
} > gpurun_out/r02_sanitizer.log 2>&1
grep -v "^$" gpurun_out/r02_sanitizer.log | cut -c1-200 | tail -30
