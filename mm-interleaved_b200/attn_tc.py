"""Dispatch to the tcgen05 flash-attention kernel (csrc/attn_fwd_sm100.cu)."""
from __future__ import annotations

import torch

from . import _lib

ENABLED = hasattr(_lib, "HAS_ATTN_TC") and _lib.HAS_ATTN_TC


def supported(q, k, v, Tq, Tkv, hd) -> bool:
    return False


def forward(q, k, v, out, key_mask, causal, past, scale):
    raise RuntimeError("tensor-core attention kernel is not built")
