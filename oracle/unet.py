"""oracle/unet.py -- independent fp32 restatement of the SD-2.1-base UNet forward with the MMFS hook.
TEST INFRASTRUCTURE ONLY.  **Parity unpinned**: the arithmetic lives in diffusers==0.20.0 (requirements.txt:9), a
third-party dependency that is neither under /root/reference nor in this image, and the reference ships no fixture
for it.  What the reference DOES own is the forward's control flow -- its patched ``UNet2DConditionModel.forward``
(mm_interleaved/models/utils/monkey_patch/sd_unet_forward_monkey_patch.py:17-371) -- which this file follows line by
line; the block algorithms restate diffusers 0.20.0's published modules, cited per function.

Written against a flat state dict with diffusers' parameter names (the keys of a reference checkpoint's
``image_decoder.decoder.unet.*``); the block structure is DERIVED FROM THE KEYS (which blocks carry ``attentions``, how
many ``resnets``, whether a ``downsamplers`` / ``upsamplers`` entry exists), not from any module of the product, and
uses only torch.nn.functional ops.
"""
from __future__ import annotations

import math
import re

import torch
import torch.nn.functional as F


def _count(sd, prefix, what):
    idx = {int(m.group(1)) for k in sd for m in [re.match(re.escape(prefix) + what + r"\.(\d+)\.", k)] if m}
    return max(idx) + 1 if idx else 0


def timesteps_ref(t, dim, max_period=10000.0):
    """diffusers ``get_timestep_embedding(flip_sin_to_cos=True, downscale_freq_shift=0)`` (embeddings.py): frequencies
    exp(-ln(10000) * i / half); the [sin | cos] halves swapped to [cos | sin].  Always fp32 (forward :121-126)."""
    half = dim // 2
    exponent = -math.log(max_period) * torch.arange(half, dtype=torch.float32) / (half - 0.0)
    emb = t[:, None].float() * torch.exp(exponent)[None, :]
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
    return torch.cat([emb[:, half:], emb[:, :half]], dim=-1)


def _lin(sd, p, x):
    return F.linear(x, sd[p + ".weight"], sd.get(p + ".bias"))


def _conv(sd, p, x, stride=1, padding=1):
    return F.conv2d(x, sd[p + ".weight"], sd.get(p + ".bias"), stride=stride, padding=padding)


def resnet_ref(sd, p, x, temb, groups=32, eps=1e-5):
    """``ResnetBlock2D.forward`` (resnet.py): GN -> SiLU -> conv1 -> + time_emb_proj(SiLU(temb)) -> GN -> SiLU -> conv2,
    1x1 ``conv_shortcut`` when the channel count changes, output_scale_factor 1."""
    h = F.silu(F.group_norm(x, groups, sd[p + ".norm1.weight"], sd[p + ".norm1.bias"], eps))
    h = _conv(sd, p + ".conv1", h)
    h = h + _lin(sd, p + ".time_emb_proj", F.silu(temb))[:, :, None, None]
    h = F.silu(F.group_norm(h, groups, sd[p + ".norm2.weight"], sd[p + ".norm2.bias"], eps))
    h = _conv(sd, p + ".conv2", h)
    if p + ".conv_shortcut.weight" in sd:
        x = _conv(sd, p + ".conv_shortcut", x, padding=0)
    return x + h


def attention_ref(sd, p, x, ctx, heads):
    """``Attention`` + ``AttnProcessor`` (attention_processor.py): q/k/v without bias, softmax(q k^T * d^-1/2) v per head,
    ``to_out.0`` with bias.  (The reference enables the xformers processor, sd.py:64-65: same function.)"""
    ctx = x if ctx is None else ctx
    B, T, _ = x.shape
    q, k, v = _lin(sd, p + ".to_q", x), _lin(sd, p + ".to_k", ctx), _lin(sd, p + ".to_v", ctx)
    d = q.shape[-1] // heads
    q, k, v = (t.view(B, -1, heads, d).transpose(1, 2) for t in (q, k, v))
    w = torch.softmax(torch.matmul(q, k.transpose(-1, -2)) * d ** -0.5, dim=-1)
    o = torch.matmul(w, v).transpose(1, 2).reshape(B, T, heads * d)
    return _lin(sd, p + ".to_out.0", o)


def transformer_block_ref(sd, p, x, ctx, heads):
    """``BasicTransformerBlock.forward`` (attention.py): LN -> self-attn (+), LN -> cross-attn (+), LN -> GEGLU FF (+)."""
    ln = lambda n, t: F.layer_norm(t, (t.shape[-1],), sd[f"{p}.{n}.weight"], sd[f"{p}.{n}.bias"], 1e-5)
    x = x + attention_ref(sd, p + ".attn1", ln("norm1", x), None, heads)
    x = x + attention_ref(sd, p + ".attn2", ln("norm2", x), ctx, heads)
    h, gate = _lin(sd, p + ".ff.net.0.proj", ln("norm3", x)).chunk(2, dim=-1)        # GEGLU: value * gelu(gate)
    return x + _lin(sd, p + ".ff.net.2", h * F.gelu(gate))


def transformer2d_ref(sd, p, x, ctx, heads=None, groups=32, head_dim=64):
    """``Transformer2DModel.forward`` with ``use_linear_projection`` (SD 2.x; transformer_2d.py): GN(eps 1e-6) ->
    (B, HW, C) -> proj_in -> blocks -> proj_out -> (B, C, H, W) + residual."""
    B, C, H, W = x.shape
    inner = sd[p + ".proj_in.weight"].shape[0]
    if sd[p + ".proj_in.weight"].dim() != 2:
        raise NotImplementedError("conv projections (SD 1.x) are not the SD-2.1 configuration")
    h = F.group_norm(x, groups, sd[p + ".norm.weight"], sd[p + ".norm.bias"], 1e-6)
    h = h.permute(0, 2, 3, 1).reshape(B, H * W, C)
    h = _lin(sd, p + ".proj_in", h)
    for i in range(_count(sd, p + ".", "transformer_blocks")):
        h = transformer_block_ref(sd, f"{p}.transformer_blocks.{i}", h, ctx, heads if heads else inner // head_dim)
    h = _lin(sd, p + ".proj_out", h)
    return h.reshape(B, H, W, C).permute(0, 3, 1, 2) + x


def unet_forward_ref(sd, sample, timestep, encoder_hidden_states, mmfs_features=None, mmfs_mask=None, mmfs_module=None,
                     attention_head_dim=(5, 10, 20, 20)):
    """Patched forward, sd_unet_forward_monkey_patch.py: time embedding :103-128, conv_in :236, down blocks :253-283
    (``down_block_res_samples`` collects conv_in's output and every resnet(+attention) / downsampler output), mid block
    :301-311, the MMFS hook :316-326 (``mmfs_module(sample, down_block_res_samples, mmfs_features, mmfs_mask)`` -- any
    callable with that signature, e.g. the MMFSNet oracle), up blocks :329-362 (each pops ``len(resnets)`` skips),
    conv_norm_out -> SiLU -> conv_out :365-368.  fp32 CPU tensors.  ``attention_head_dim`` is the UNet config entry of
    that name, which diffusers 0.20 uses as the NUMBER of heads per down block (unet_2d_condition.py: ``num_attention_heads
    = num_attention_heads or attention_head_dim``; SD-2.1: 5/10/20/20 heads of 64); the mid block takes the last entry,
    the up blocks the reversed list."""
    heads = list(attention_head_dim)
    sd = {k: v.float() for k, v in sd.items()}
    t = torch.as_tensor(timestep).reshape(-1).expand(sample.shape[0])
    temb_dim = sd["time_embedding.linear_1.weight"].shape[1]
    emb = _lin(sd, "time_embedding.linear_2", F.silu(_lin(sd, "time_embedding.linear_1", timesteps_ref(t, temb_dim))))
    sample = _conv(sd, "conv_in", sample.float())
    ctx = encoder_hidden_states.float()
    res = (sample,)
    for b in range(_count(sd, "", "down_blocks")):
        p = f"down_blocks.{b}"
        has_attn = _count(sd, p + ".", "attentions") > 0
        for i in range(_count(sd, p + ".", "resnets")):
            sample = resnet_ref(sd, f"{p}.resnets.{i}", sample, emb)
            if has_attn:
                sample = transformer2d_ref(sd, f"{p}.attentions.{i}", sample, ctx, heads[b])
            res += (sample,)
        if _count(sd, p + ".", "downsamplers") > 0:              # Downsample2D: 3x3 conv, stride 2, padding 1
            sample = _conv(sd, f"{p}.downsamplers.0.conv", sample, stride=2, padding=1)
            res += (sample,)
    # UNetMidBlock2DCrossAttn: resnets[0], then (attention, resnet) pairs
    sample = resnet_ref(sd, "mid_block.resnets.0", sample, emb)
    for i in range(_count(sd, "mid_block.", "attentions")):
        sample = transformer2d_ref(sd, f"mid_block.attentions.{i}", sample, ctx, heads[-1])
        sample = resnet_ref(sd, f"mid_block.resnets.{i + 1}", sample, emb)
    if mmfs_module is not None:                                   # MODIFICATION START / END of the reference patch
        sample, res = mmfs_module(sample, res, mmfs_features, mmfs_mask)
    res = tuple(res)
    for b in range(_count(sd, "", "up_blocks")):
        p = f"up_blocks.{b}"
        n = _count(sd, p + ".", "resnets")
        skips, res = res[-n:], res[:-n]
        has_attn = _count(sd, p + ".", "attentions") > 0
        for i in range(n):
            sample = torch.cat([sample, skips[-1 - i]], dim=1)    # res_hidden_states_tuple[-1] popped per resnet
            sample = resnet_ref(sd, f"{p}.resnets.{i}", sample, emb)
            if has_attn:
                sample = transformer2d_ref(sd, f"{p}.attentions.{i}", sample, ctx, heads[::-1][b])
        if _count(sd, p + ".", "upsamplers") > 0:                 # Upsample2D: nearest x2, then 3x3 conv
            sample = _conv(sd, f"{p}.upsamplers.0.conv", F.interpolate(sample, scale_factor=2.0, mode="nearest"))
    sample = F.silu(F.group_norm(sample, 32, sd["conv_norm_out.weight"], sd["conv_norm_out.bias"], 1e-5))
    return _conv(sd, "conv_out", sample)
