"""Repeatability / accuracy probe of linear_skinny on one shape: prints how many outputs differ between two runs and
where, and the error against an fp64 statement."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mm_interleaved_b200 import ops  # noqa: E402

M, N, K = 4, int(sys.argv[1]) if len(sys.argv) > 1 else 27648, int(sys.argv[2]) if len(sys.argv) > 2 else 5120
PRO = int(sys.argv[3]) if len(sys.argv) > 3 else 0
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn((M, K), device="cuda", dtype=torch.bfloat16, generator=g)
w = torch.randn((N, K), device="cuda", dtype=torch.bfloat16, generator=g) * K ** -0.5
nw = (1.0 + 0.1 * torch.randn((K,), device="cuda", generator=g)).to(torch.bfloat16) if PRO == 1 else None
with torch.no_grad():
    ys = [ops.linear_skinny(x, w, norm_weight=nw, eps=1e-6).float() for _ in range(6)]
    xa = x if PRO == 0 else (nw * (x.float() * torch.rsqrt(x.float().pow(2).mean(-1, keepdim=True) + 1e-6)).to(torch.bfloat16))
    ref = (xa.double() @ w.double().t()).float()
for i in range(1, 6):
    d = (ys[i] != ys[0])
    cols = d.any(0).nonzero().flatten()
    print(f"run {i} vs 0: {int(d.sum())} differing outputs, rows {d.any(1).nonzero().flatten().tolist()}, blocks {sorted(set((cols // 32).tolist()))[:12]}, "
          f"max diff {float((ys[i] - ys[0]).abs().max()):.3g}, nan {int(torch.isnan(ys[i]).sum())}")
err = (ys[0] - ref).abs()
print(f"max |err| {float(err.max()):.4g} at col {int(err.max(0).values.argmax())}, ref max {float(ref.abs().max()):.3g}; "
      f"cols with err > 0.05: {(err.max(0).values > 0.05).nonzero().flatten()[:16].tolist()}")
