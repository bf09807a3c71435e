"""Generate the committed golden fixtures from the UNMODIFIED reference, executed in the build
container (``/root/reference`` is absent on the GPU box, so the vectors travel instead).

    python tests/golden/make_golden.py            # writes tests/golden/*.npz

The reference ships no golden vectors for this path (the pickle that
ops/tests/compare_with_data.py:112-114 expects is not in the repository), so these fixtures
are outputs of the reference's own code run here:

  msda_core_*.npz   ms_deform_attn_core_pytorch (ops/functions/ms_deform_attn_func.py:47-67),
                    fp32 and fp64, on seeded inputs that are stored alongside the outputs
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import make_msda_inputs  # noqa: E402
from oracle import ref_loader  # noqa: E402

# name -> (N, spatial_shapes, M, D, Lq, P, seed, loc_mode)
MSDA_CASES = {
    # the reference test-script shape (ops/tests/forward_backward_error.py:175-206)
    "ref_script": (1, [(6, 4), (3, 2)], 2, 64, 2, 2, 0, "uniform"),
    "tiny_edges": (2, [(8, 8), (4, 6), (3, 2)], 4, 32, 37, 4, 1, "edges"),
    "llm_like": (1, [(32, 32), (16, 16), (8, 8)] * 2, 16, 64, 24, 8, 2, "clustered"),
    "sd_like": (1, [(64, 64), (32, 32), (16, 16), (8, 8)], 16, 64, 16, 8, 3, "uniform"),
    "adapter_like": (2, [(16, 16)], 16, 32, 40, 4, 4, "clustered"),
    "odd_dims": (1, [(5, 7), (2, 3)], 3, 24, 9, 3, 5, "edges"),
}


def input_digest(value, loc, attn) -> str:
    import hashlib
    h = hashlib.sha1()
    for t in (value, loc, attn):
        h.update(t.contiguous().numpy().tobytes())
    return h.hexdigest()


def main():
    ref = ref_loader.load()
    core = ref.func.ms_deform_attn_core_pytorch
    for name, (N, shapes, M, D, Lq, P, seed, mode) in MSDA_CASES.items():
        value, ss, starts, loc, attn = make_msda_inputs(N, shapes, M, D, Lq, P, seed=seed, loc_mode=mode)
        out32 = core(value, ss, loc, attn)
        out64 = core(value.double(), ss, loc.double(), attn.double())
        path = os.path.join(HERE, f"msda_core_{name}.npz")
        arrays = dict(spatial_shapes=ss.numpy(), level_start_index=starts.numpy(),
                      out_fp32=out32.numpy(), out_fp64=out64.numpy(),
                      params=np.array([N, M, D, Lq, P, seed], dtype=np.int64), loc_mode=np.array(mode),
                      input_sha1=np.array(input_digest(value, loc, attn)))
        if value.numel() + loc.numel() + attn.numel() <= 200_000:
            # small cases carry their inputs; large ones are regenerated from the seed by
            # oracle.make_msda_inputs and verified against input_sha1
            arrays.update(value=value.numpy(), sampling_loc=loc.numpy(), attn_weight=attn.numpy())
        np.savez_compressed(path, **arrays)
        print(f"{path}: {os.path.getsize(path) / 1024:.0f} KiB  out {tuple(out32.shape)}")


if __name__ == "__main__":
    torch.manual_seed(0)
    main()
