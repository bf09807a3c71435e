"""Kernel-time breakdown (torch.profiler, CUDA activities) of greedy decode steps of the 13B decoder after a B x 2048-token
prefill: eager in-place-cache loop, 1 + N tokens minus 1 token isolates the N decode steps."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from benchmarks import workloads  # noqa: E402
from mm_interleaved_b200.mm_interleaved import InterleavedForward  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

B, n_new = int(os.environ.get("LOCAL_BATCH", 4)), int(os.environ.get("N_NEW", 8))
wl = workloads.InterleavedCfg3(0, 1, B)
wl.make_host_inputs(pin=False)
model = workloads.full_model(with_image_decoder=False)
ids, img, nimg = (t.cuda() for t in wl.host)
with torch.no_grad():
    vis = model._tokenize(img)
    gen = lambda n: InterleavedForward.generate_texts(model, ids, vis, nimg, wl.N_IMG, max_new_tokens=n, eos_token_id=None)
    gen(2)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        gen(1 + n_new)
        torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
tot = sum(e.device_time_total for e in rows)
print(f"total device time {tot / 1e3:.1f} ms for prefill + {n_new + 1} tokens (batch {B}); kernels with count multiple of {n_new + 1} are per-token")
for e in rows[:int(os.environ.get('TOP', 45))]:
    print(f"{e.device_time_total / 1e3:9.2f} ms  n={e.count:5d}  avg={e.device_time_total / max(e.count, 1):8.1f} us  {e.key[:100]}")
