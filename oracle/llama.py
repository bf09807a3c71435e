"""oracle/llama.py -- CPU restatement of the reference's Llama-MMFS decoder.  TEST INFRASTRUCTURE ONLY.

Follows /root/reference/mm_interleaved/models/decoders/modeling_llama_mmfs.py:
  _make_causal_mask / _expand_mask :21-50, LlamaRMSNorm :53-70, rotary tables :119-151,
  rotate_half / apply_rotary_pos_emb :158-172, LlamaMLP :175-189, LlamaAttention.forward :217-280,
  LlamaMMFSAttention.forward :346-367, LlamaDecoderLayer.forward :385-450, LlamaModel.forward :623-752.
Parameters come as a flat dict with the reference's state-dict names.  Pinned against the committed
outputs of the reference LlamaModel (tests/golden/llama_*.npz) and the live reference in the build
container (tests/test_oracle_llama.py).  Also the body of bench.py's CPU baseline for the decoder.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .mmfs import llama_mmfs_attention_ref, rms_norm_ref


def rotary_tables_ref(dim, max_pos, base=10000.0):
    inv_freq = 1.0 / (base ** (torch.arange(0, dim, 2).float() / dim))           # :125
    t = torch.arange(max_pos, dtype=torch.float32)
    freqs = torch.outer(t, inv_freq)
    emb = torch.cat((freqs, freqs), dim=-1)                                      # :138
    return emb.cos(), emb.sin()


def _rotate_half(x):
    x1, x2 = x[..., : x.shape[-1] // 2], x[..., x.shape[-1] // 2:]
    return torch.cat((-x2, x1), dim=-1)                                          # :158-162


def additive_mask_ref(attention_mask, q_len, past, dtype):
    """causal (+ key padding) additive mask (B,1,q_len,kv_len): :21-50, :599-620."""
    B, kv_len = attention_mask.shape
    neg = torch.finfo(dtype).min
    mask = None
    if q_len > 1:
        causal = torch.full((q_len, q_len), neg, dtype=dtype)
        causal = causal.masked_fill(torch.arange(q_len)[None, :] <= torch.arange(q_len)[:, None], 0)
        if past > 0:
            causal = torch.cat([torch.zeros(q_len, past, dtype=dtype), causal], dim=-1)
        mask = causal[None, None].expand(B, 1, q_len, kv_len)
    inverted = 1.0 - attention_mask[:, None, None, :].to(dtype).expand(B, 1, q_len, kv_len)
    expanded = inverted.masked_fill(inverted.to(torch.bool), neg)
    return expanded if mask is None else expanded + mask


def llama_attention_ref(p, prefix, x, add_mask, position_ids, n_heads, past_kv=None):
    B, T, C = x.shape
    hd = C // n_heads
    q = F.linear(x, p[f"{prefix}q_proj.weight"]).view(B, T, n_heads, hd).transpose(1, 2)
    k = F.linear(x, p[f"{prefix}k_proj.weight"]).view(B, T, n_heads, hd).transpose(1, 2)
    v = F.linear(x, p[f"{prefix}v_proj.weight"]).view(B, T, n_heads, hd).transpose(1, 2)
    kv_len = T + (past_kv[0].shape[2] if past_kv is not None else 0)
    cos, sin = rotary_tables_ref(hd, max(kv_len, int(position_ids.max()) + 1))
    cos = cos.to(x.dtype)[position_ids][:, None]                                 # gather by position, :165-169
    sin = sin.to(x.dtype)[position_ids][:, None]
    q = q * cos + _rotate_half(q) * sin
    k = k * cos + _rotate_half(k) * sin
    if past_kv is not None:
        k = torch.cat([past_kv[0], k], dim=2)
        v = torch.cat([past_kv[1], v], dim=2)
    w = torch.matmul(q * (hd ** -0.5), k.transpose(2, 3))                        # :246
    if add_mask is not None:
        w = w + add_mask
        w = torch.max(w, torch.tensor(torch.finfo(w.dtype).min))                 # :258
    w = F.softmax(w, dim=-1, dtype=torch.float32).to(q.dtype)                    # :261
    o = torch.matmul(w, v).transpose(1, 2).reshape(B, T, C)
    return F.linear(o, p[f"{prefix}o_proj.weight"]), (k, v)


def llama_layer_ref(p, idx, x, vision, cross_mask, add_mask, position_ids, cfg, past_kv=None):
    pre = f"layers.{idx}."
    h = rms_norm_ref(x, p[pre + "input_layernorm.weight"], cfg["eps"])
    a, kv = llama_attention_ref(p, pre + "self_attn.", h, add_mask, position_ids, cfg["n_heads"], past_kv)
    x = x + a
    if pre + "llama_cross_attn.gate" in p and vision is not None:                # :427
        pc = {k[len(pre + "llama_cross_attn."):]: t for k, t in p.items() if k.startswith(pre + "llama_cross_attn.")}
        x = x + llama_mmfs_attention_ref(pc, x, vision, cross_mask, spatial_shapes=cfg["spatial_shapes"],
                                         eps=cfg["eps"], sampler=cfg.get("sampler"))
    h = rms_norm_ref(x, p[pre + "post_attention_layernorm.weight"], cfg["eps"])
    gate = F.silu(F.linear(h, p[pre + "mlp.gate_proj.weight"]))
    x = x + F.linear(gate * F.linear(h, p[pre + "mlp.up_proj.weight"]), p[pre + "mlp.down_proj.weight"])
    return x, kv


def llama_model_ref(p, inputs_embeds, attention_mask, position_ids, vision, cross_mask, cfg, past=None, layers=None):
    """LlamaModel.forward on inputs_embeds; returns (last_hidden_state, [per-layer (k, v)])."""
    B, T, _ = inputs_embeds.shape
    past_len = past[0][0].shape[2] if past is not None else 0
    if attention_mask is None:
        attention_mask = torch.ones((B, past_len + T))
    add_mask = additive_mask_ref(attention_mask, T, past_len, inputs_embeds.dtype)
    if position_ids is None:
        position_ids = torch.arange(past_len, past_len + T)[None].expand(B, T)
    x = inputs_embeds
    kvs = []
    n_layers = cfg["n_layers"] if layers is None else layers
    for i in range(n_layers):
        x, kv = llama_layer_ref(p, i, x, vision, cross_mask, add_mask, position_ids, cfg,
                                past[i] if past is not None else None)
        kvs.append(kv)
    return rms_norm_ref(x, p["norm.weight"], cfg["eps"]), kvs
