"""The reference's ``replace_*()`` start-up mechanism (inference.py:10-24) pointed at the B200 classes: after
``replace_*_b200()`` the REFERENCE's own ``LlamaModel`` / ``LlamaMMFSAttention`` / ``MMFSNet`` constructors build B200
modules, and a state dict saved from the unpatched reference loads into them with ``strict=True``.  CPU-only
(construction + state-dict load; compute needs the GPU).  Skipped where /root/reference is absent (the GPU box)."""
import sys

import pytest
import torch

from oracle import ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present (GPU box)")


def _ref_llama_config():
    from transformers import LlamaConfig
    from tests.golden.make_golden import LLAMA_TINY
    extra = ("cross_attention_frequency", "spatial_shapes", "image_embed_dim")
    cfg = LlamaConfig(**{k: v for k, v in LLAMA_TINY.items() if k not in extra}, hidden_act="silu")
    for k in extra:
        setattr(cfg, k, LLAMA_TINY[k])
    return cfg


def test_replace_functions_rebind_the_reference_and_checkpoints_load():
    import mm_interleaved_b200 as b200
    from mm_interleaved_b200 import patch
    ns = ref_loader.load()
    if ns.llama is None or ns.sd_mmfs is None:
        pytest.skip("reference decoders do not import in this environment")
    ref_mmfs_cls, ref_layer_cls, ref_net_cls = ns.mmfs.MMFS, ns.llama.LlamaDecoderLayer, ns.sd_mmfs.MMFSNet
    cfg = _ref_llama_config()
    ref_sd = ns.llama.LlamaModel(cfg).state_dict()                       # unpatched reference: the checkpoint layout
    from tests.golden.make_golden import MMFSNET_TINY
    ref_net_sd = ref_net_cls(**MMFSNET_TINY).state_dict()
    try:
        b200.replace_mmfs_b200()
        assert ns.mmfs.MMFS is b200.MMFS and ns.sd_mmfs.MMFS is b200.MMFS   # (the llama file imports it lazily, :325)
        assert sys.modules["mm_interleaved.models.utils.ops.modules"].MMFS is b200.MMFS
        assert sys.modules["MultiScaleDeformableAttention"].ms_deform_attn_forward is b200.ms_deform_attn_forward
        assert ns.func.MSDA is sys.modules["MultiScaleDeformableAttention"]
        # reference LlamaMMFSAttention (still the reference class) now builds the B200 MMFS inside
        xattn = ns.llama.LlamaMMFSAttention(cfg, layer_idx=0)
        assert type(xattn.attn) is b200.MMFS and type(xattn).__module__.startswith("mm_interleaved.")
        xattn.load_state_dict({k[len("layers.0.llama_cross_attn."):]: v for k, v in ref_sd.items()
                               if k.startswith("layers.0.llama_cross_attn.")}, strict=True)

        b200.replace_llama_b200()
        for name in ("LlamaRMSNorm", "LlamaMLP", "LlamaAttention", "LlamaMMFSAttention", "LlamaDecoderLayer"):
            assert getattr(ns.llama, name) is getattr(b200, name)
        model = ns.llama.LlamaModel(cfg)                                  # the reference's own LlamaModel glue
        assert type(model).__module__.startswith("mm_interleaved.")
        assert all(type(l) is b200.LlamaDecoderLayer for l in model.layers)
        assert type(model.layers[0].llama_cross_attn.attn) is b200.MMFS and type(model.norm) is b200.LlamaRMSNorm
        missing, unexpected = model.load_state_dict(ref_sd, strict=True)
        assert not missing and not unexpected
        assert set(model.state_dict().keys()) == set(ref_sd.keys())

        b200.replace_visual_b200()
        assert ns.sd_mmfs.MMFSNet is b200.MMFSNet and ns.sd_mmfs.MMFSBlock is b200.MMFSBlock
        net = ns.sd_mmfs.MMFSNet(**MMFSNET_TINY)
        net.load_state_dict(ref_net_sd, strict=True)

        b200.replace_all_b200()                                           # idempotent
        assert ns.llama.LlamaDecoderLayer is b200.LlamaDecoderLayer
    finally:
        patch.restore_reference()
    assert ns.mmfs.MMFS is ref_mmfs_cls and ns.sd_mmfs.MMFS is ref_mmfs_cls and ns.llama.LlamaDecoderLayer is ref_layer_cls
    assert ns.sd_mmfs.MMFSNet is ref_net_cls


def test_b200_top_level_model_keeps_the_reference_state_dict_layout():
    """``MMInterleaved`` (B200) exposes the reference's parameter names for the parts whose reference classes import
    here (decoder + text head + glue): every key of the reference ``LlamaModel`` / ``TextDecoder`` layout exists."""
    import mm_interleaved_b200 as b200
    from tests.golden.make_golden import LLAMA_TINY
    ns = ref_loader.load()
    if ns.llama is None:
        pytest.skip("reference decoder does not import")
    ref_keys = {"mm_decoder." + k for k in ns.llama.LlamaModel(_ref_llama_config()).state_dict().keys()}
    tiny_vt = dict(clip_config=b200.visual_tokenizer.CLIPVisionConfigLite(hidden_size=64, intermediate_size=128,
                                                                        num_hidden_layers=4, num_attention_heads=4,
                                                                        image_size=32, patch_size=8),
                   perceiver_config=dict(num_queries=3, hidden_size=48, encoder_hidden_size=64, num_hidden_layers=2,
                                         num_attention_heads=4, intermediate_size=96, cross_attention_frequency=1,
                                         qk_normalization=True), grid_size=4)
    try:
        model = b200.MMInterleaved(llm_config=dict(LLAMA_TINY, vocab_size=62), txt_vocab_size=64, seq_len=32,
                                   visual_tokenizer_config=tiny_vt, image_embed_dim=LLAMA_TINY["image_embed_dim"],
                                   cross_attention_frequency=2, spatial_shapes=LLAMA_TINY["spatial_shapes"])
    except Exception as e:                                                # tiny tokenizer dims unsupported on this build
        pytest.skip(f"tiny visual tokenizer config not constructible: {e}")
    keys = set(model.state_dict().keys())
    assert ref_keys <= keys, sorted(ref_keys - keys)[:5]
    for k in ("text_decoder.head.weight", "text_decoder.head.bias", "text_decoder.head_new.weight",
              "text_decoder.head_new.bias", "soi_token", "context_feat_proj.weight", "context_feat_proj.bias"):
        assert k in keys, k
    assert model.text_decoder.head.weight.shape[0] == 64 and model.text_decoder.head_new.weight.shape[0] == 2
    assert any(k.startswith("visual_tokenizer.encoder.") for k in keys)
