"""oracle/msda.py -- CPU checker for the native multi-scale deformable attention op.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py for who may import it).

Two independent statements of the op are provided:

* ``msda_forward_ref``  -- ctypes wrapper around oracle/msda_ref.c, the scalar
  restatement of the reference CUDA kernel
  (ops/src/cuda/ms_deform_im2col_cuda.cuh:36-87, 240-302).  This one also emits the
  integer index stream, which only a restatement of the .cuh can pin bit-exactly.
* ``msda_core_pytorch`` -- restatement of the reference's own pure-PyTorch core
  ``ms_deform_attn_core_pytorch`` (ops/functions/ms_deform_attn_func.py:47-67):
  per level, ``grid_sample(bilinear, zeros, align_corners=False)`` on ``2*loc-1``,
  then the attention-weighted sum over (level, point).  BASELINE.json configs[0]
  names this as the reference's CPU-runnable path, so it is also what
  ``bench.py --impl reference`` / ``cpu_baseline`` time.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle_msda.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile oracle/msda_ref.c into oracle/liboracle_msda.so (gcc, see Makefile)."""
    src = os.path.join(_HERE, "msda_ref.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle_msda.so"],
                              stdout=subprocess.DEVNULL)
    return _LIB_PATH


def _load():
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(_LIB_PATH)
        i, l, p = ctypes.c_int, ctypes.c_long, ctypes.c_void_p
        lib.msda_ref_forward_f32.argtypes = [p, p, p, p, p, p, p, i, i, i, i, i, i, i, l, l]
        lib.msda_ref_forward_f32.restype = None
        lib.msda_ref_forward_f64.argtypes = [p, p, p, p, p, p, i, i, i, i, i, i, i, l, l]
        lib.msda_ref_forward_f64.restype = None
        lib.msda_ref_idx_fields.restype = ctypes.c_int
        _lib = lib
    return _lib


def level_start_index(spatial_shapes) -> torch.Tensor:
    """Prefix sum of H*W per level (reference: modeling_llama_mmfs.py:304-305)."""
    ss = torch.as_tensor(spatial_shapes, dtype=torch.long).reshape(-1, 2)
    return torch.cat((ss.new_zeros((1,)), ss.prod(1).cumsum(0)[:-1]))


def round_to_dtype(x: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """Round an fp32/fp64 tensor to ``dtype`` storage and widen back (what the device
    load ``opmath_t x = data[i]`` sees, cuh:283-285)."""
    if dtype in (torch.float32, torch.float64):
        return x.to(dtype)
    return x.to(dtype).to(torch.float32)


def msda_forward_ref(value, spatial_shapes, level_start, sampling_loc, attn_weight,
                     want_index_stream: bool = False, threads: int | None = None):
    """Scalar restatement of the reference kernel (oracle/msda_ref.c).

    value (N,S,M,D); spatial_shapes (L,2) int64 [H,W]; level_start (L,) int64;
    sampling_loc (N,Lq,M,L,P,2) last dim (x,y); attn_weight (N,Lq,M,L,P).
    Inputs are CPU tensors in fp32 (already rounded to the storage dtype under test)
    or fp64.  Returns out (N,Lq,M*D) in the opmath type (un-rounded accumulator) and,
    if asked, the int32 index stream (N,Lq,M,L,P,8) =
    [in_range, h_low, w_low, valid_mask(bit0..3 = v1..v4), ptr1..ptr4] for channel 0.
    """
    lib = _load()
    f64 = value.dtype == torch.float64
    dt = torch.float64 if f64 else torch.float32
    value = value.detach().to("cpu", dt).contiguous()
    loc = sampling_loc.detach().to("cpu", dt).contiguous()
    attn = attn_weight.detach().to("cpu", dt).contiguous()
    shapes = torch.as_tensor(spatial_shapes).to("cpu", torch.long).contiguous()
    starts = torch.as_tensor(level_start).to("cpu", torch.long).contiguous()
    N, S, M, D = value.shape
    _, Lq, _, L, P, _ = loc.shape
    assert shapes.shape == (L, 2) and starts.shape == (L,)
    assert attn.shape == (N, Lq, M, L, P)
    assert int((shapes[:, 0] * shapes[:, 1]).sum()) <= S
    out = torch.zeros((N, Lq, M * D), dtype=dt)  # cu:55  at::zeros
    idx = None
    if want_index_stream:
        assert not f64
        idx = torch.zeros((N, Lq, M, L, P, lib.msda_ref_idx_fields()), dtype=torch.int32)
    total = N * Lq
    nthreads = max(1, min(threads or (os.cpu_count() or 1), total))
    bounds = np.linspace(0, total, nthreads + 1).astype(np.int64)

    def run(k):
        a, b = int(bounds[k]), int(bounds[k + 1])
        if a == b:
            return
        if f64:
            lib.msda_ref_forward_f64(value.data_ptr(), shapes.data_ptr(), starts.data_ptr(),
                                     loc.data_ptr(), attn.data_ptr(), out.data_ptr(),
                                     N, S, M, D, L, Lq, P, a, b)
        else:
            lib.msda_ref_forward_f32(value.data_ptr(), shapes.data_ptr(), starts.data_ptr(),
                                     loc.data_ptr(), attn.data_ptr(), out.data_ptr(),
                                     idx.data_ptr() if idx is not None else None,
                                     N, S, M, D, L, Lq, P, a, b)

    if nthreads == 1:
        run(0)
    else:
        with ThreadPoolExecutor(nthreads) as ex:  # ctypes releases the GIL
            list(ex.map(run, range(nthreads)))
    return (out, idx) if want_index_stream else out


def msda_core_pytorch(value, spatial_shapes, sampling_loc, attn_weight):
    """Restatement of ``ms_deform_attn_core_pytorch``
    (ops/functions/ms_deform_attn_func.py:47-67): the reference's CPU-runnable path.

    Level by level, the (N,H*W,M,D) slab becomes an (N*M, D, H, W) image, the
    normalised locations become a grid in [-1,1] and ``grid_sample`` (bilinear, zero
    padding, align_corners=False) fetches (N*M, D, Lq, P); the (L*P) samples are then
    combined with the attention weights.  Returns (N, Lq, M*D) in value's dtype.
    """
    N, S, M, D = value.shape
    _, Lq, _, L, P, _ = sampling_loc.shape
    hw = [(int(h), int(w)) for h, w in torch.as_tensor(spatial_shapes).tolist()]
    per_level = value.split([h * w for h, w in hw], dim=1)
    grid = 2 * sampling_loc - 1                                            # func.py:53
    sampled = []
    for lvl, (h, w) in enumerate(hw):
        img = per_level[lvl].flatten(2).transpose(1, 2).reshape(N * M, D, h, w)   # func.py:57
        g = grid[:, :, :, lvl].transpose(1, 2).flatten(0, 1)                     # (N*M,Lq,P,2)
        sampled.append(F.grid_sample(img, g, mode="bilinear", padding_mode="zeros",
                                     align_corners=False))                        # func.py:61-62
    aw = attn_weight.transpose(1, 2).reshape(N * M, 1, Lq, L * P)          # func.py:65
    out = (torch.stack(sampled, dim=-2).flatten(-2) * aw).sum(-1).view(N, M * D, Lq)
    return out.transpose(1, 2).contiguous()


def error_metrics(out: torch.Tensor, ref: torch.Tensor, eps: float = 1e-6):
    """The reference's acceptance metric (ops/tests/forward_backward_error.py:164-169):
    max-abs, max-rel and mean-rel error with relative errors taken where |ref| > eps."""
    out = out.detach().double().cpu()
    ref = ref.detach().double().cpu()
    abs_err = (out - ref).abs()
    mask = ref.abs() > eps
    rel = abs_err[mask] / ref[mask].abs()
    return {
        "max_abs": float(abs_err.max()) if abs_err.numel() else 0.0,
        "max_rel": float(rel.max()) if rel.numel() else 0.0,
        "mean_rel": float(rel.mean()) if rel.numel() else 0.0,
    }


def make_msda_inputs(N, spatial_shapes, M, D, Lq, P, seed=0, loc_mode="uniform",
                     dtype=torch.float32):
    """Seeded synthetic inputs (reference convention
    ops/tests/create_data.py:18-21): value ~ U[0,1), weights = (U+1e-5) normalised
    over (L,P).  ``loc_mode``:
      'uniform'   loc ~ U[0,1) drawn in float64 then rounded -> FULL-MANTISSA fp32
                  coordinates (torch.rand(float32) only yields multiples of 2^-24,
                  which makes the index-parity test vacuous, SURVEY.md 8a')
      'clustered' loc ~ N(0.5, 0.15) -> a few percent fall outside [0,1]: exercises
                  the in-range predicate and the border validity bits
      'edges'     loc concentrated within +-1.5 px of the map borders and on exact
                  cell boundaries (k/W, (k+0.5)/W)
    All tensors are generated on the CPU in fp32/fp64 and rounded to ``dtype`` storage.
    """
    g = torch.Generator().manual_seed(seed)
    shapes = torch.as_tensor(spatial_shapes, dtype=torch.long).reshape(-1, 2)
    L = shapes.shape[0]
    S = int((shapes[:, 0] * shapes[:, 1]).sum())
    starts = level_start_index(shapes)
    value = torch.rand((N, S, M, D), generator=g, dtype=torch.float32)
    if loc_mode == "uniform":
        loc = torch.rand((N, Lq, M, L, P, 2), generator=g, dtype=torch.float64)
    elif loc_mode == "clustered":
        loc = 0.5 + 0.15 * torch.randn((N, Lq, M, L, P, 2), generator=g, dtype=torch.float64)
        far = torch.rand((N, Lq, M, L, P, 1), generator=g) < 0.05
        loc = torch.where(far, loc * 3.0 - 1.0, loc)
    elif loc_mode == "edges":
        wh = torch.stack([shapes[:, 1], shapes[:, 0]], -1).double()      # (L,2) = (W,H)
        u = torch.rand((N, Lq, M, L, P, 2), generator=g, dtype=torch.float64)
        k = torch.floor(u * (wh[None, None, None, :, None, :] + 3)) - 1   # -1 .. W+1
        frac = torch.randint(0, 4, (N, Lq, M, L, P, 2), generator=g).double() * 0.5
        jitter = (torch.rand((N, Lq, M, L, P, 2), generator=g, dtype=torch.float64) - 0.5) * 1e-6
        jitter = jitter * (torch.rand((N, Lq, M, L, P, 2), generator=g) < 0.5)
        loc = (k + frac) / wh[None, None, None, :, None, :] + jitter
    else:
        raise ValueError(loc_mode)
    attn = torch.rand((N, Lq, M, L, P), generator=g, dtype=torch.float32) + 1e-5
    attn = attn / attn.sum(-1, keepdim=True).sum(-2, keepdim=True)
    if dtype == torch.float64:
        return value.double(), shapes, starts, loc, attn.double()
    loc = loc.float()
    return (round_to_dtype(value, dtype), shapes, starts,
            round_to_dtype(loc, dtype), round_to_dtype(attn, dtype))
