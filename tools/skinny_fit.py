"""linear_skinny vs cuBLAS over N at fixed K (plain prologue, no residual): time = c0 + bytes / BW.  One CUDA graph of
NW calls over rotating weight copies, device time / NW."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mm_interleaved_b200 import _lib, ops  # noqa: E402

M, K = 4, int(os.environ.get("K", 5120))
MODE = int(os.environ.get("SK_MODE", 0))
assert _lib.lib().mmfs_linear_skinny_set_tuning(MODE) == 0
g = torch.Generator(device="cuda").manual_seed(0)


def timed(fn, nw, reps=15):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for i in range(nw):
            fn(i)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for i in range(nw):
            fn(i)
    graph.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); graph.replay(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / nw)
    ts.sort()
    return ts[len(ts) // 2]


print(f"mode {MODE}, K = {K}, rows = {M}")
with torch.no_grad():
    for N in (1280, 2560, 5120, 10240, 20480, 40960):
        nw = max(4, min(48, int(400e6 // (N * K * 2)) + 1))                 # > L2 in total
        ws = [torch.randn((N, K), device="cuda", dtype=torch.bfloat16, generator=g) * K ** -0.5 for _ in range(nw)]
        x = torch.randn((M, K), device="cuda", dtype=torch.bfloat16, generator=g)
        out = torch.empty((M, N), device="cuda", dtype=torch.bfloat16)
        ta = timed(lambda i: ops.linear_skinny(x, ws[i], out=out), nw)
        tb = timed(lambda i: torch.mm(x, ws[i].t(), out=out), nw)
        mb = N * K * 2 / 1e6
        print(f"N={N:6d} ({mb:6.1f} MB, {nw:2d} copies): linear_skinny {ta:6.1f} us | cuBLAS {tb:6.1f} us")
        del ws
