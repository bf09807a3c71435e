#!/bin/bash
# GPU session script (kept in-tree so the exact commands behind profiles/ are reproducible)
mkdir -p gpurun_out
timeout 300 python tools/sampler_sweep.py 20 > gpurun_out/r02_sampler_sweep2.log 2>&1
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:mmfs_sampler_v2 -c 1 -f -o /tmp/r02_sampler_v2 python tools/sampler_one.py 4 2 masked v2 0 3 > gpurun_out/ncu_v2.log 2>&1
timeout 300 $NCU -k regex:mmfs_sampler_kernel -c 1 -f -o /tmp/r02_sampler_generic python tools/sampler_one.py 4 2 masked generic > gpurun_out/ncu_generic.log 2>&1
for n in r02_sampler_v2 r02_sampler_generic; do
  ncu -i /tmp/$n.ncu-rep --page details > gpurun_out/${n}_ncu_details.txt 2>&1
  ncu -i /tmp/$n.ncu-rep --page source --csv > gpurun_out/${n}_ncu_source.csv 2>&1
  ncu -i /tmp/$n.ncu-rep --page raw --csv > gpurun_out/${n}_ncu_raw.csv 2>&1
done
cp /tmp/r02_sampler_v2.ncu-rep gpurun_out/ 2>/dev/null
timeout 600 python tools/decode_bench.py > gpurun_out/r02_decode_bench.json 2> gpurun_out/r02_decode_bench.err
tail -c 3000 gpurun_out/r02_decode_bench.err
timeout 300 python tools/gemm_ab.py 20 > gpurun_out/r02_gemm_ab.log 2>&1
timeout 600 $NCU -k regex:"conv_igemm|groupnorm|attn_fwd" -c 40 -f -o /tmp/r02_unet_kernels python tools/unet_one.py 1 > gpurun_out/ncu_unet.log 2>&1
ncu -i /tmp/r02_unet_kernels.ncu-rep --page raw --csv > gpurun_out/r02_unet_kernels_raw.csv 2>&1
ncu -i /tmp/r02_unet_kernels.ncu-rep --page details > gpurun_out/r02_unet_kernels_details.txt 2>&1
timeout 1200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_step_cfg3.csv python tools/step_for_ncu.py interleaved_cfg3 > gpurun_out/r02_step_under_ncu.log 2>&1
du -sh gpurun_out; ls -la gpurun_out | tail -25
