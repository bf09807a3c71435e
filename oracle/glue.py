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


def context_features_for_image_decoder_ref(context_features, text_ids, soi_token_id, proj_weight, proj_bias, seq_len,
                                           nearest_bos_idxs=None):
    """mm_interleaved.py:254-304, loop for loop (numpy table of utils/pos_embed.py:77-95)."""
    import numpy as np
    image_start_token_idx = (text_ids == soi_token_id).nonzero(as_tuple=True)[-1]
    if nearest_bos_idxs is None:
        nearest_bos_idxs = torch.zeros_like(image_start_token_idx)
    row_ids = (text_ids == soi_token_id).nonzero(as_tuple=True)[0]
    B_I, C = image_start_token_idx.shape[0], context_features.shape[-1]
    lengths = image_start_token_idx - nearest_bos_idxs + 1
    L_max = int(max(lengths))
    per_image = torch.zeros((B_I, L_max, C)).type_as(context_features)
    mask = torch.zeros((B_I, L_max)).type_as(image_start_token_idx)
    for i in range(B_I):
        f = context_features[row_ids[i], nearest_bos_idxs[i]: image_start_token_idx[i] + 1, :].flip(dims=(0,))
        per_image[i, : lengths[i], :] = f
        mask[i, : lengths[i]] = 1
    omega = np.arange(C // 2, dtype=np.float32)
    omega /= C / 2.0
    omega = 1.0 / 10000 ** omega
    out = np.einsum("m,d->md", np.arange(seq_len, dtype=np.float32), omega)
    pos = torch.from_numpy(np.concatenate([np.sin(out), np.cos(out)], axis=1)).type_as(context_features)
    per_image = torch.nn.functional.linear(per_image, proj_weight, proj_bias) + pos[None, :L_max]
    return per_image, mask


def mmfs_features_for_image_decoder_ref(multiscale_features, text_ids, soi_token_id, nearest_bos_idxs=None):
    """mm_interleaved.py:306-340, loop for loop."""
    L = text_ids.shape[1]
    B_I = multiscale_features[0].shape[0]
    ix, iy = (text_ids == soi_token_id).nonzero(as_tuple=True)
    idx = ix * L + iy
    if nearest_bos_idxs is None:
        nearest_bos_idxs = torch.zeros_like(idx)
    nb = ix * L + nearest_bos_idxs
    m = nb[:, None] <= idx[None, :]
    m = torch.triu(torch.tril(m, diagonal=-1), diagonal=-1)
    feats = [torch.zeros_like(f)[:, None] for f in multiscale_features]
    mask = torch.zeros((B_I, 1), dtype=torch.long)
    for i in range(B_I):
        sel = m[i].nonzero(as_tuple=True)[-1]
        for src, dst in zip(multiscale_features, feats):
            dst[i, : len(sel)] = src[sel]
        mask[i, : len(sel)] = 1
    return feats, mask
