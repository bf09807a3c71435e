"""RoPE in place on the qkv buffer at the cfg-3 shape (8192 tokens x 40 heads x 128): CUDA-graph timing over 3 buffers."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mm_interleaved_b200 import ops  # noqa: E402
from mm_interleaved_b200.llama_mmfs import rotary_tables  # noqa: E402

B, T, H, hd = 4, 2048, 40, 128
bufs = [torch.randn((B, T, 3, H, hd), device="cuda", dtype=torch.bfloat16) for _ in range(3)]
cos, sin = rotary_tables(hd, 2048, device="cuda")
pos = torch.arange(T, device="cuda")
fn = lambda i: ops.rope_qk_(bufs[i][:, :, 0], bufs[i][:, :, 1], cos, sin, pos)
for i in range(3):
    fn(i)
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for i in range(3): fn(i)
torch.cuda.current_stream().wait_stream(s)
with torch.cuda.graph(g):
    for i in range(3): fn(i)
g.replay(); torch.cuda.synchronize()
ts = []
for _ in range(15):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3 / 3)
t = sorted(ts)[7]
nbytes = B * T * 2 * H * hd * 2 * 2
print(f"rope_qk_ cfg3: {t:.1f} us = {nbytes / t / 1e6:.2f} TB/s (q, k read + written)")
