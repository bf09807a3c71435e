"""Kernel-time breakdown of greedy decode steps of the 13B decoder (after a 4 x 2048-token prefill)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from benchmarks import workloads  # noqa: E402

wl = workloads.make("interleaved_cfg3", rank=0, world=1, local_batch=4)
wl.setup()
ids, img = wl.dev[0], wl.dev[1]
from torch.profiler import ProfilerActivity, profile  # noqa: E402

with torch.no_grad():
    wl.tok_in.copy_(img)
    wl.tok_graph.replay()
    vis = wl.tok_out
    wl.model.generate_texts(ids, vis, wl.nimg, wl.N_IMG, max_new_tokens=2, eos_token_id=None)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        wl.model.generate_texts(ids, vis, wl.nimg, wl.N_IMG, max_new_tokens=5, eos_token_id=None)
        torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=80))
