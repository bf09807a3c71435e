"""oracle/sd_mmfs.py -- CPU restatement of the reference's MMFSBlock / MMFSNet forward
(mm_interleaved/models/decoders/sd_mmfs.py:99-145, 230-272).  TEST INFRASTRUCTURE ONLY.
Parameters: flat dict with the reference's state-dict names.  Pinned by tests/golden/mmfsnet_tiny.npz."""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from .mmfs import mmfs_forward_ref


def _abs_pos(abs_pos, tgt_len):                                                       # utils/pos_embed.py:16-40
    src, tgt = int(math.sqrt(abs_pos.shape[0])), int(math.sqrt(tgt_len))
    if src == tgt:
        return abs_pos
    x = abs_pos.float().reshape(1, src, src, -1).permute(0, 3, 1, 2)
    x = F.interpolate(x, size=(tgt, tgt), mode="bicubic", align_corners=False)
    return x.permute(0, 2, 3, 1).flatten(0, 2).to(abs_pos.dtype)


def _ref_points(h, w):                                                                 # sd_mmfs.py:15-28
    ry, rx = torch.meshgrid(torch.linspace(0.5, h - 0.5, h), torch.linspace(0.5, w - 0.5, w), indexing="ij")
    return torch.stack((rx.reshape(-1)[None] / w, ry.reshape(-1)[None] / h), -1)[:, :, None]


def mmfs_block_ref(p, prefix, sample, ms_feat, mask, spatial_shapes, *, n_heads=16, n_points=8, base_spatial_shape):
    B, C, H, W = sample.shape
    n_img = mask.shape[-1]
    ss = torch.tensor(list(spatial_shapes) * n_img, dtype=torch.long)
    starts = torch.cat((ss.new_zeros((1,)), ss.prod(1).cumsum(0)[:-1]))
    q = sample.flatten(2).transpose(1, 2)
    q = F.layer_norm(q, (C,), p[prefix + "query_norm.weight"], p[prefix + "query_norm.bias"], 1e-6)
    q = q + _abs_pos(p[prefix + "pos_embed"], H * W)
    feat = F.layer_norm(ms_feat, (ms_feat.shape[-1],), p[prefix + "feat_norm.weight"], p[prefix + "feat_norm.bias"], 1e-6)
    pm = {k[len(prefix + "mmfs."):]: v for k, v in p.items() if k.startswith(prefix + "mmfs.")}
    scale = torch.tensor([s[0] / base_spatial_shape for s in spatial_shapes])
    out = mmfs_forward_ref(pm, q, _ref_points(H, W), feat, ss, starts, mask, n_heads=n_heads,
                           n_levels=len(spatial_shapes), n_points=n_points, scale_ratios=scale)
    out = out.transpose(1, 2).reshape(B, C, H, W)
    return F.conv2d(out, p[prefix + "conv.weight"], p[prefix + "conv.bias"])


def mmfsnet_ref(p, sample, down_res, mmfs_features, mask, *, downsample_factor, n_down):
    shapes = [(int(f.shape[-2]), int(f.shape[-1])) for f in mmfs_features]
    feats = torch.cat([f.flatten(3).transpose(2, 3) for f in mmfs_features], dim=2)
    sd_shapes = [s[0] // downsample_factor for s in shapes]
    new = []
    for i, res in enumerate(down_res):
        r = mmfs_block_ref(p, f"mmfs_down_blocks.{i}.", res, feats, mask, shapes, base_spatial_shape=sd_shapes[i // 3])
        new.append(res + r)
    mid = mmfs_block_ref(p, "mmfs_mid_block.", sample, feats, mask, shapes, base_spatial_shape=sd_shapes[-1])
    return sample + mid, new
