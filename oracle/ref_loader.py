"""oracle/ref_loader.py -- import the UNMODIFIED reference modules in the build container.

TEST INFRASTRUCTURE ONLY.  ``/root/reference`` exists only in the build container
(never on the GPU box); everything here is used by ``tests/golden/make_golden.py`` to
generate the committed fixtures and by ``-m "not gpu"`` tests that are skipped when
the reference tree is absent.

``import mm_interleaved.models`` fails under transformers 5.x (SURVEY.md preamble), so
the parent packages are registered as empty stubs and only the leaf modules on the hot
path are executed from their files:

  ops/functions/ms_deform_attn_func.py   MSDeformAttnFunction, ms_deform_attn_core_pytorch
  ops/modules/mmfs.py                    MMFS
  decoders/modeling_llama_mmfs.py        LlamaModel, LlamaDecoderLayer, LlamaMMFSAttention ...
  decoders/sd_mmfs.py                    MMFSBlock, MMFSNet
  utils/pos_embed.py

The reference CUDA op cannot run here (no GPU, and the op has no CPU implementation,
ops/src/ms_deform_attn.h:38), so ``MSDeformAttnFunction.apply`` is routed to the
reference's own ``ms_deform_attn_core_pytorch`` -- the substitution SURVEY.md 8c
describes.  Nothing in the reference tree is modified or copied.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("MMI_REFERENCE_ROOT", "/root/reference")
_PKG = "mm_interleaved"


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, _PKG, "models", "utils", "ops"))


def _stub(name: str, path: str):
    if name in sys.modules:
        return sys.modules[name]
    mod = types.ModuleType(name)
    mod.__path__ = [path]
    mod.__package__ = name
    sys.modules[name] = mod
    return mod


def _load(name: str, file: str):
    if name in sys.modules and getattr(sys.modules[name], "__file__", None) == file:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, file)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


class _CoreFunction:
    """Stands in for MSDeformAttnFunction (func.py:24-44) with the reference's own
    PyTorch core (func.py:47-67) as the body."""

    core = None

    @classmethod
    def apply(cls, value, shapes, level_start_index, sampling_locations, attention_weights, im2col_step):
        return cls.core(value, shapes, sampling_locations, attention_weights)


def load():
    """Returns a namespace with the reference leaf modules (func, mmfs, llama, sd_mmfs, pos_embed)."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    root = os.path.join(REF_ROOT, _PKG)
    _stub(_PKG, root)
    _stub(f"{_PKG}.models", os.path.join(root, "models"))
    _stub(f"{_PKG}.models.utils", os.path.join(root, "models", "utils"))
    _stub(f"{_PKG}.models.decoders", os.path.join(root, "models", "decoders"))
    ops = os.path.join(root, "models", "utils", "ops")
    _stub(f"{_PKG}.models.utils.ops", ops)
    fpkg = _stub(f"{_PKG}.models.utils.ops.functions", os.path.join(ops, "functions"))
    mpkg = _stub(f"{_PKG}.models.utils.ops.modules", os.path.join(ops, "modules"))

    func = _load(f"{_PKG}.models.utils.ops.functions.ms_deform_attn_func",
                 os.path.join(ops, "functions", "ms_deform_attn_func.py"))
    _CoreFunction.core = staticmethod(func.ms_deform_attn_core_pytorch)
    fpkg.MSDeformAttnFunction = _CoreFunction          # what `from ..functions import` sees
    fpkg.ms_deform_attn_core_pytorch = func.ms_deform_attn_core_pytorch
    mmfs = _load(f"{_PKG}.models.utils.ops.modules.mmfs", os.path.join(ops, "modules", "mmfs.py"))
    mpkg.MMFS = mmfs.MMFS
    sys.modules[f"{_PKG}.models.utils.ops"].modules = mpkg
    pos_embed = _load(f"{_PKG}.models.utils.pos_embed",
                      os.path.join(root, "models", "utils", "pos_embed.py"))
    ns = types.SimpleNamespace(func=func, mmfs=mmfs, pos_embed=pos_embed, llama=None, sd_mmfs=None)
    try:
        ns.llama = _load(f"{_PKG}.models.decoders.modeling_llama_mmfs",
                         os.path.join(root, "models", "decoders", "modeling_llama_mmfs.py"))
    except Exception as e:  # pragma: no cover - depends on the transformers version
        ns.llama_error = e
    try:
        ns.sd_mmfs = _load(f"{_PKG}.models.decoders.sd_mmfs",
                           os.path.join(root, "models", "decoders", "sd_mmfs.py"))
    except Exception as e:  # pragma: no cover
        ns.sd_mmfs_error = e
    try:   # ViT-Adapter building blocks (need a timm stub: DropPath is never active at drop_path = 0)
        if "timm" not in sys.modules:
            timm = types.ModuleType("timm"); timm.models = types.ModuleType("timm.models")
            timm.models.layers = types.ModuleType("timm.models.layers")
            timm.models.layers.DropPath = type("DropPath", (torch_nn().Identity,), {"__init__": lambda self, p=0.0: torch_nn().Identity.__init__(self)})
            sys.modules.update({"timm": timm, "timm.models": timm.models, "timm.models.layers": timm.models.layers})
        enc = os.path.join(root, "models", "encoders")
        _stub(f"{_PKG}.models.encoders", enc)
        va = os.path.join(enc, "vit_adapter")
        _stub(f"{_PKG}.models.encoders.vit_adapter", va)
        vops = os.path.join(va, "ops")
        _stub(f"{_PKG}.models.encoders.vit_adapter.ops", vops)
        vf = _stub(f"{_PKG}.models.encoders.vit_adapter.ops.functions", os.path.join(vops, "functions"))
        vfunc = _load(f"{_PKG}.models.encoders.vit_adapter.ops.functions.ms_deform_attn_func",
                      os.path.join(vops, "functions", "ms_deform_attn_func.py"))
        vf.MSDeformAttnFunction = _CoreFunction
        vm = _stub(f"{_PKG}.models.encoders.vit_adapter.ops.modules", os.path.join(vops, "modules"))
        vmod = _load(f"{_PKG}.models.encoders.vit_adapter.ops.modules.ms_deform_attn",
                     os.path.join(vops, "modules", "ms_deform_attn.py"))
        vmod.MSDeformAttnFunction = _CoreFunction
        vm.MSDeformAttn = vmod.MSDeformAttn
        ns.adapter = _load(f"{_PKG}.models.encoders.vit_adapter.adapter_modules", os.path.join(va, "adapter_modules.py"))
        ns.adapter_msda = vmod
    except Exception as e:  # pragma: no cover
        ns.adapter = None
        ns.adapter_error = e
    return ns


def torch_nn():
    import torch.nn as nn
    return nn
