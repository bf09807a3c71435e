"""oracle/glue.py -- loop-for-loop restatement of the reference's batch-preparation helpers
(mm_interleaved/models/mm_interleaved.py:121-252).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import torch


def prepare_mm_embeds_ref(text_embeds, text_ids, image_embeds, soi_token, image_token_id, soi_token_id):
    B, L, C = text_embeds.shape                                                     # :131-170
    px, py = (text_ids == image_token_id).nonzero(as_tuple=True)
    pos = (px * L + py)[:, None].expand(-1, C)
    flat = text_embeds.reshape(B * L, C).to(image_embeds.dtype)
    mm = torch.scatter(flat, 0, pos, image_embeds.reshape(-1, C))
    sx, sy = (text_ids == soi_token_id).nonzero(as_tuple=True)
    spos = (sx * L + sy)[:, None].expand(-1, C)
    mm = torch.scatter_add(mm, 0, spos, soi_token.repeat(spos.shape[0], 1).to(mm.dtype))
    return mm.view(B, L, C)


def cross_attention_mask_ref(text_ids, num_image_per_seq, bos_token_id, soi_token_id):
    B, L = text_ids.shape                                                           # :192-221
    max_num_image = int(num_image_per_seq.max())
    soi_pos = (text_ids == soi_token_id).nonzero(as_tuple=True)[1]
    image_token_pos = -1 * torch.ones(B, max_num_image).type_as(soi_pos)
    start = 0
    for i in range(B):
        n = int(num_image_per_seq[i])
        image_token_pos[i, :n] = soi_pos[start:start + n] + 1
        start += n
    image_token_pos = image_token_pos[..., None].repeat(1, 1, L)
    idx = torch.arange(L)[None, :].repeat(B, 1)
    nearest_bos = idx.masked_fill(text_ids != bos_token_id, -1).cummax(dim=1).values
    mask = (image_token_pos > nearest_bos[:, None, :]) * (image_token_pos <= torch.arange(L)[None, None, :]) * \
           (image_token_pos != -1)
    return mask.transpose(-1, -2).float()


def pack_mmfs_features_ref(multiscale_features, spatial_shapes, num_image_per_seq):
    B = num_image_per_seq.shape[0]                                                  # :223-250
    max_num_image = int(num_image_per_seq.max())
    feats = [f for f in multiscale_features if int(f.shape[-1]) in spatial_shapes]
    out = []
    for f in feats:
        buf = torch.zeros(B, max_num_image, *f.shape[1:], dtype=f.dtype)
        start = 0
        for i in range(B):
            n = int(num_image_per_seq[i])
            buf[i, :n] = f[start:start + n]
            start += n
        out.append(buf.flatten(3).transpose(2, 3))                                  # b n c h w -> b n (h w) c
    return torch.cat(out, dim=2)
