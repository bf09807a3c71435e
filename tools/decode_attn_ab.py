"""Decode attention (q_len = 1) at the cfg-3 cache: KV layout A/B.  (B, T, H, hd) -- the projection GEMM's own layout,
256-byte key rows 10 KB apart -- against the same bytes laid out (B*H, T, 1, hd) (each head's keys contiguous)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mm_interleaved_b200 import ops  # noqa: E402

B, H, T, hd = 4, 40, 2080, 128
g = torch.Generator(device="cuda").manual_seed(0)


def bench(q, ks, vs, reps=20):
    """One CUDA graph of len(ks) * 2 calls over rotating cache buffers (3 x 170 MB > L2), replayed: device time per call
    without the host-side launch path."""
    n = len(ks) * 2
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for i in range(n):
            ops.attention(q, ks[i % len(ks)], vs[i % len(ks)], causal=True, past=T - 1)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for i in range(n):
            ops.attention(q, ks[i % len(ks)], vs[i % len(ks)], causal=True, past=T - 1)
    graph.replay()
    torch.cuda.synchronize()
    ts = []
    for i in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); graph.replay(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / n)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


from mm_interleaved_b200 import _lib  # noqa: E402
W = int(os.environ.get("DEC_WARPS", 0))
assert _lib.lib().mmfs_attn_decode_set_tuning(W) == 0
print(f"decode attention: {W or 4} warps per 256-key CTA")
nbytes = 2 * B * T * H * hd * 2
for name, shape, qshape in (("B,T,H,hd", (B, T, H, hd), (B, 1, H, hd)), ("B*H,T,1,hd", (B * H, T, 1, hd), (B * H, 1, 1, hd))):
    ks = [torch.randn(shape, device="cuda", dtype=torch.bfloat16, generator=g) for _ in range(3)]   # 3 x 170 MB > L2
    vs = [torch.randn(shape, device="cuda", dtype=torch.bfloat16, generator=g) for _ in range(3)]
    q = torch.randn(qshape, device="cuda", dtype=torch.bfloat16, generator=g)
    med, best = bench(q, ks, vs)
    print(f"{name:12s}: median {med:6.1f} us  best {best:6.1f} us  {nbytes / med / 1e6:6.2f} TB/s (K+V bytes / median)")
