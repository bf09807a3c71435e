#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q --maxfail=20 2>&1 | tail -30) > gpurun_out/r02_pytest22.log 2>&1
tail -3 gpurun_out/r02_pytest22.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke5.log 2>&1; tail -2 gpurun_out/r02_smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench5.json 2> gpurun_out/r02_bench5.err
tail -3 gpurun_out/r02_bench5.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench5_reference.json 2> gpurun_out/r02_bench5_reference.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench5.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','sequences_per_s')}, d['e2e']['value'], d['clocks'])
print('sampler', d['roofline']['frac'], d['roofline']['avg_launch_us'], 'attn', d['roofline']['attention']['frac'])
for k,v in d.get('secondary',{}).items():
    print(k, v.get('value'), v.get('ms_per_step'), v.get('error'), {kk:v.get(kk) for kk in ('images_per_s','decode_ms_per_token','generate_texts_ms','generate_images_ms','sequences_per_s')})
PY
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 2500 --csv --log-file gpurun_out/r02_launches_step_cfg3_final3.csv python tools/step_for_ncu.py > gpurun_out/r02_step_for_ncu.log 2>&1; tail -2 gpurun_out/r02_step_for_ncu.log; wc -l gpurun_out/r02_launches_step_cfg3_final3.csv
