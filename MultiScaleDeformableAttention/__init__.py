"""Drop-in for the reference's native extension module.

The reference does ``import MultiScaleDeformableAttention as MSDA`` and calls
``MSDA.ms_deform_attn_forward(value, spatial_shapes, level_start_index, sampling_loc,
attn_weight, im2col_step)`` / ``MSDA.ms_deform_attn_backward(..., grad_output, im2col_step)``
(ops/functions/ms_deform_attn_func.py:18-21,29-30,41-42 and its twin under
encoders/vit_adapter/ops/functions/).  This package exposes the same two names, backed by
libmmfs_b200.so, so both call sites run unchanged once this directory is on ``sys.path``.
"""
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
if _root not in _sys.path:
    _sys.path.insert(0, _root)

from mm_interleaved_b200.msda import ms_deform_attn_backward, ms_deform_attn_forward  # noqa: E402,F401
