#!/bin/bash
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4) > gpurun_out/r02_pytest24.log 2>&1; tail -2 gpurun_out/r02_pytest24.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-80
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench6.json 2> gpurun_out/r02_bench6.err; tail -2 gpurun_out/r02_bench6.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench6_reference.json 2> gpurun_out/r02_bench6_reference.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench6.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','sequences_per_s')}, d['e2e']['value'], d['clocks'])
print('sampler', d['roofline']['frac'], d['roofline']['avg_launch_us'], 'attn', d['roofline']['attention']['frac'])
for k,v in d.get('secondary',{}).items():
    print(k, v.get('value'), v.get('ms_per_step'), v.get('error'), {kk:v.get(kk) for kk in ('images_per_s','decode_ms_per_token','generate_texts_ms','generate_images_ms','sequences_per_s')})
PY
