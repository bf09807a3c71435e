#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_unet_gpu.py tests/test_unet_oracle.py tests/test_mmfsnet.py tests/test_mm_interleaved_gpu.py tests/test_cache_safety.py -m gpu -q 2>&1 | tail -12) > gpurun_out/r02_pytest23.log 2>&1
tail -12 gpurun_out/r02_pytest23.log | cut -c1-200
timeout 900 python bench.py --steps 3 --warmup 2 --no-secondary --workload sd_cfg4 > gpurun_out/r02_bench_cfg4_graph.json 2> gpurun_out/r02_bench_cfg4_graph.err || tail -5 gpurun_out/r02_bench_cfg4_graph.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r02_bench_cfg4_graph.json'))
    print(d['value'], d['ms_per_step'], d.get('images_per_s'), d['roofline']['frac'], d['roofline']['attention']['frac'], d['e2e']['value'])
except Exception as e:
    print('parse failed', e)
PY
