"""Run the MSDA forward a few times on one BASELINE shape (target for ncu captures)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mm_interleaved_b200 as m  # noqa: E402
from oracle import make_msda_inputs  # noqa: E402
from tools.msda_sweep import SHAPES  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg3_llm_L12_Lq2048"
dtype = {"bf16": torch.bfloat16, "f32": torch.float32, "f16": torch.float16}[sys.argv[2] if len(sys.argv) > 2 else "bf16"]
loc_mode = sys.argv[3] if len(sys.argv) > 3 else "clustered"
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
N, shapes, M, D, Lq, P = SHAPES[name]
v, s, st, loc, a = make_msda_inputs(N, shapes, M, D, Lq, P, seed=0, loc_mode=loc_mode, dtype=dtype)
args = [v.to("cuda", dtype), s.cuda(), st.cuda(), loc.to("cuda", dtype), a.to("cuda", dtype)]
for _ in range(reps):
    out = m.ms_deform_attn_forward(*args, 64)
torch.cuda.synchronize()
print(name, out.float().abs().mean().item())
