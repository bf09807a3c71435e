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


class AttrDict(dict):
    """dict with attribute access (what the reference's OmegaConf configs look like to its constructors)."""
    __getattr__ = dict.__getitem__


def _xformers_stub():
    """xformers 0.0.20 is not in this image.  The reference calls exactly one function of it on the encoder path,
    ``xformers.ops.memory_efficient_attention(q, k, v, attn_bias)`` with (B, T, H, hd) operands (xattn.py:70-72), whose
    published semantics are softmax(q k^T / sqrt(hd) + bias) v -- also spelled out in the reference's own commented
    eager code (xattn.py:75-135).  That formula is what the stub computes."""
    import torch
    xf, xo = types.ModuleType("xformers"), types.ModuleType("xformers.ops")

    class LowerTriangularMask:  # noqa: D401 - marker type only
        pass

    def memory_efficient_attention(q, k, v, attn_bias=None, p=0.0, scale=None):
        s = torch.einsum("bqhd,bkhd->bhqk", q, k) * (q.shape[-1] ** -0.5 if scale is None else scale)
        if attn_bias is not None:
            T, S = q.shape[1], k.shape[1]
            s = s.masked_fill(~torch.tril(torch.ones(T, S, dtype=torch.bool)), float("-inf"))
        return torch.einsum("bhqk,bkhd->bqhd", torch.softmax(s, -1), v)

    xo.memory_efficient_attention, xo.LowerTriangularMask, xf.ops = memory_efficient_attention, LowerTriangularMask, xo
    return xf, xo


def load_visual():
    """The reference's visual tokenizer stack, importable here with three shims (none touches /root/reference):
    (1) the docstring decorators / constants ``vit_adapter_hf.py`` imports from transformers 4.31 that 5.x dropped are
    supplied as no-ops; (2) ``xformers`` is the stub above; (3) ``timm.DropPath`` is Identity (never active).  The CLIP
    encoder layers and the Q-Former glue then come from the transformers version in this image (5.x) as a stand-in
    for the pinned 4.31; the attention classes (``CLIPXAttention``, the qk-norm ``Blip2QFormerMultiHeadAttention``), the
    adapter, ``PerceiverResampler`` and ``VisualTokenizer`` are the reference's own files.  Returns a namespace."""
    ns = load()                                    # registers the package stubs + adapter_modules (timm stub)
    import transformers.models.clip.modeling_clip as mc
    import transformers.utils as tu
    for name in ("CLIP_VISION_INPUTS_DOCSTRING", "CLIP_START_DOCSTRING"):
        if not hasattr(mc, name):
            setattr(mc, name, "")

    def _passthrough(*a, **k):
        return lambda fn: fn

    for name in ("add_start_docstrings", "add_start_docstrings_to_model_forward", "replace_return_docstrings"):
        if not hasattr(tu, name):
            setattr(tu, name, _passthrough)
    if "xformers" not in sys.modules:
        xf, xo = _xformers_stub()
        sys.modules["xformers"], sys.modules["xformers.ops"] = xf, xo
    root = os.path.join(REF_ROOT, _PKG, "models")
    va = os.path.join(root, "encoders", "vit_adapter")
    pk = f"{_PKG}.models.encoders.vit_adapter"
    mp = os.path.join(root, "utils", "monkey_patch")
    _stub(f"{_PKG}.models.utils.monkey_patch", mp)
    qk = _load(f"{_PKG}.models.utils.monkey_patch.blip2_qknorm_monkey_patch", os.path.join(mp, "blip2_qknorm_monkey_patch.py"))
    qk.replace_blip2_attn_with_qknorm_attn()       # inference.py:19
    ns.qknorm = qk
    ns.xattn = _load(pk + ".xattn", os.path.join(va, "xattn.py"))
    ns.clip_vit = _load(pk + ".clip_vit_hf", os.path.join(va, "clip_vit_hf.py"))
    ns.vit_adapter = _load(pk + ".vit_adapter_hf", os.path.join(va, "vit_adapter_hf.py"))
    sys.modules[pk].clip_vit_adapter_hf = ns.vit_adapter.clip_vit_adapter_hf
    ns.perceiver = _load(f"{_PKG}.models.decoders.perceiver", os.path.join(root, "decoders", "perceiver.py"))
    ns.visual_tokenizer = _load(f"{_PKG}.models.encoders.visual_tokenizer", os.path.join(root, "encoders", "visual_tokenizer.py"))
    return ns
