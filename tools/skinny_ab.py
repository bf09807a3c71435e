"""Decode-step linears at batch 4: this repo's linear_skinny (with the RMSNorm / SwiGLU in front folded in) against
cuBLAS (F.linear / addmm_) + the stand-alone norm / activation kernel, per shape of the 13 B decoder.  Each variant is
one CUDA graph over 6 rotating weight copies (> L2), replayed; time per call = device time / 6."""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mm_interleaved_b200 import ops  # noqa: E402

M = int(os.environ.get("ROWS", 4))
H, I = 5120, 13824
SHAPES = [  # name, N, K, prologue, residual
    ("qkv      (rmsnorm ->)", 3 * H, H, 1, False),
    ("o_proj   (+ residual)", H, H, 0, True),
    ("gate|up  (rmsnorm ->)", 2 * I, H, 1, False),
    ("down     (swiglu ->, + residual)", H, I, 2, True),
]
NW = 6
g = torch.Generator(device="cuda").manual_seed(0)


def timed(fn, reps=15):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for i in range(NW):
            fn(i)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for i in range(NW):
            fn(i)
    graph.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); graph.replay(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / NW)
    ts.sort()
    return ts[len(ts) // 2]


from mm_interleaved_b200 import _lib  # noqa: E402
MODE = int(os.environ.get("SK_MODE", 0))
assert _lib.lib().mmfs_linear_skinny_set_tuning(MODE) == 0
print(f"linear_skinny mode {MODE} (0 default, 1 per-lane cp.async, 2 tensor-map TMA)")
tot_a = tot_b = 0.0
with torch.no_grad():
    for name, N, K, pro, with_res in SHAPES:
        ws = [(torch.randn((N, K), device="cuda", dtype=torch.bfloat16, generator=g) * K ** -0.5) for _ in range(NW)]
        x = torch.randn((M, K * (2 if pro == 2 else 1)), device="cuda", dtype=torch.bfloat16, generator=g)
        nw = torch.ones((K,), device="cuda", dtype=torch.bfloat16)
        res = torch.randn((M, N), device="cuda", dtype=torch.bfloat16, generator=g)
        out = torch.empty((M, N), device="cuda", dtype=torch.bfloat16)

        def ours(i):
            ops.linear_skinny(x, ws[i], residual=res if with_res else None, out=res if with_res else out,
                              norm_weight=nw if pro == 1 else None, eps=1e-6, swiglu=pro == 2)

        def cublas(i):
            a = ops.rmsnorm(x, nw, 1e-6) if pro == 1 else (ops.swiglu(x) if pro == 2 else x)
            if with_res:
                res.addmm_(a, ws[i].t())
            else:
                torch.mm(a, ws[i].t(), out=out)

        ta, tb = timed(ours), timed(cublas)
        tot_a += ta; tot_b += tb
        nbytes = N * K * 2
        print(f"{name:34s} N={N:6d} K={K:6d}: linear_skinny {ta:6.1f} us = {nbytes / ta / 1e6:5.2f} TB/s | "
              f"cuBLAS + prologue kernel {tb:6.1f} us = {nbytes / tb / 1e6:5.2f} TB/s")
print(f"per layer: {tot_a:.1f} us vs {tot_b:.1f} us; x 40 layers = {tot_a * 40 / 1e3:.2f} ms vs {tot_b * 40 / 1e3:.2f} ms per token (rows = {M})")
