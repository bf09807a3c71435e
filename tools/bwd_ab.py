"""MSDA backward: deterministic (64-bit fixed-point integer atomics) vs float-atomic path, cfg-3 layer shape, CUDA events."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mm_interleaved_b200 as m  # noqa: E402
from oracle import make_msda_inputs  # noqa: E402

rows = []
for name, N, shapes, Lq in (("cfg3_layer_1seq", 1, [(32, 32), (16, 16), (8, 8)] * 4, 2048), ("sd_Lq4096", 1, [(64, 64), (32, 32), (16, 16), (8, 8)], 4096)):
    v, s, st, loc, a = make_msda_inputs(N, shapes, 16, 64, Lq, 8, seed=1, loc_mode="clustered", dtype=torch.bfloat16)
    go = torch.randn((N, Lq, 1024)).to(torch.bfloat16)
    args = [v.cuda().to(torch.bfloat16), s.cuda(), st.cuda(), loc.cuda().to(torch.bfloat16), a.cuda().to(torch.bfloat16), go.cuda()]
    out = {}
    for det in (True, False):
        for _ in range(3):
            m.ms_deform_attn_backward(*args, 64, deterministic=det)
        ts = []
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); m.ms_deform_attn_backward(*args, 64, deterministic=det); e1.record()
            torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        out["deterministic_us" if det else "float_atomics_us"] = round(sorted(ts)[5] * 1e3, 1)
    rows.append(dict(shape=name, **out))
    print(rows[-1], flush=True)
print(json.dumps(rows))
