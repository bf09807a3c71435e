"""Helpers to load the committed golden fixtures (tests/golden/*.npz)."""
from __future__ import annotations

import glob
import hashlib
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")


def msda_case_names():
    return sorted(os.path.basename(p)[len("msda_core_"):-4] for p in glob.glob(os.path.join(GOLDEN, "msda_core_*.npz")))


def load_msda_case(name):
    """Returns dict(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, out_fp32, out_fp64)."""
    from oracle import make_msda_inputs
    z = np.load(os.path.join(GOLDEN, f"msda_core_{name}.npz"))
    N, M, D, Lq, P, seed = [int(v) for v in z["params"]]
    shapes = torch.from_numpy(z["spatial_shapes"])
    if "value" in z.files:
        value, loc, attn = (torch.from_numpy(z[k]) for k in ("value", "sampling_loc", "attn_weight"))
    else:
        value, _, _, loc, attn = make_msda_inputs(N, shapes, M, D, Lq, P, seed=seed, loc_mode=str(z["loc_mode"]))
    h = hashlib.sha1()
    for t in (value, loc, attn):
        h.update(t.contiguous().numpy().tobytes())
    assert h.hexdigest() == str(z["input_sha1"]), f"golden inputs of '{name}' do not reproduce (RNG drift?)"
    return dict(value=value, spatial_shapes=shapes, level_start_index=torch.from_numpy(z["level_start_index"]),
                sampling_loc=loc, attn_weight=attn, out_fp32=torch.from_numpy(z["out_fp32"]),
                out_fp64=torch.from_numpy(z["out_fp64"]))
