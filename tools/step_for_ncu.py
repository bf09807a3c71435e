"""One bench step of a workload inside a cudaProfilerStart/Stop window, after two warm steps: the ncu target for the
per-step launch list (`ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none ...`).
usage: python tools/step_for_ncu.py [interleaved_cfg3|interleaved_cfg2|sd_cfg4|generate_cfg5]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from benchmarks import workloads  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "interleaved_cfg3"
wl = workloads.make(name, rank=0, world=1, local_batch=0)
wl.setup()
for _ in range(2):
    wl.flush_l2()
    wl.step_device()
torch.cuda.synchronize()
wl.flush_l2()
torch.cuda.synchronize()
torch.cuda.profiler.start()
wl.step_device()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
wl.teardown()
print("ok")
