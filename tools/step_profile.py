"""Kernel-time breakdown (torch.profiler, CUDA activities) of one bench step of a workload.
usage: python tools/step_profile.py [interleaved_cfg3|sd_cfg4] [rows]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from benchmarks import workloads  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "interleaved_cfg3"
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 30
lb = os.environ.get("LOCAL_BATCH")
wl = workloads.make(name, rank=0, world=1, local_batch=int(lb) if lb else 0)
wl.setup()
for _ in range(2):
    wl.step_device()
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CUDA]) as prof:
    wl.step_device()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=rows, max_name_column_width=90))
