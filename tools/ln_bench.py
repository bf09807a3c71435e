"""LayerNorm kernel timing on the shapes of the path (CLIP 1024, Q-Former 768, UNet 320 / 640 / 1280 columns)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mm_interleaved_b200 import ops  # noqa: E402

for rows, cols in [(16 * 257, 1024), (16 * 1029, 1024), (16 * 64, 768), (16 * 4096, 320), (16 * 1024, 640), (16 * 256, 1280), (8192, 5120)]:
    x = torch.randn((rows, cols), device="cuda", dtype=torch.bfloat16)
    w = torch.ones(cols, device="cuda", dtype=torch.bfloat16); b = torch.zeros_like(w)
    for _ in range(3):
        ops.layernorm(x, w, b, 1e-5)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        y = ops.layernorm(x, w, b, 1e-5)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 20
    ref = torch.nn.functional.layer_norm(x.float(), (cols,), w.float(), b.float(), 1e-5)
    print(f"{rows:6d} x {cols:5d}: {us:7.1f} us  {2 * x.numel() * 2 / us / 1e6:6.2f} TB/s  max err {(y.float() - ref).abs().max().item():.3e}")
