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


def _soi_list(text_ids, soi_token_id):
    """(row, column) of every <soi> token in row-major order."""
    return [(r, c) for r in range(text_ids.shape[0]) for c in range(text_ids.shape[1]) if int(text_ids[r, c]) == soi_token_id]


def context_features_for_image_decoder_ref(context_features, text_ids, soi_token_id, proj_weight, proj_bias, seq_len,
                                           nearest_bos_idxs=None):
    """What mm_interleaved.py:254-304 produces, image by image and token by token: image i owns the decoder states from
    its context start (``nearest_bos_idxs[i]``, default 0) to its <soi>, walked BACKWARDS from the <soi>; rows are
    zero-padded to the longest context, projected (padding included) and offset by the 1-D sin-cos table
    (utils/pos_embed.py:77-95) of the first L_max positions."""
    import math
    soi = _soi_list(text_ids, soi_token_id)
    starts = [0] * len(soi) if nearest_bos_idxs is None else [int(v) for v in nearest_bos_idxs]
    lens = [c - s0 + 1 for (_, c), s0 in zip(soi, starts)]
    L_max, C = max(lens), context_features.shape[-1]
    rows = torch.zeros((len(soi), L_max, C), dtype=context_features.dtype)
    mask = torch.zeros((len(soi), L_max), dtype=torch.long)
    for i, ((r, c), n) in enumerate(zip(soi, lens)):
        for t in range(n):
            rows[i, t] = context_features[r, c - t]
            mask[i, t] = 1
    half = C // 2
    table = torch.zeros((L_max, C), dtype=torch.float32)
    import numpy as np
    omega = 1.0 / 10000 ** (np.arange(half, dtype=np.float32) / np.float32(C / 2.0))          # float32 like the numpy original
    for pos in range(L_max):
        ang = np.float32(pos) * omega
        table[pos, :half] = torch.from_numpy(np.sin(ang))
        table[pos, half:] = torch.from_numpy(np.cos(ang))
    assert L_max <= seq_len and not math.isnan(float(table.sum()))
    out = torch.nn.functional.linear(rows, proj_weight, proj_bias) + table.to(rows.dtype)[None]
    return out, mask


def mmfs_features_for_image_decoder_ref(multiscale_features, text_ids, soi_token_id, nearest_bos_idxs=None):
    """What mm_interleaved.py:306-340 produces: the lower-triangular mask restricted to the first sub-diagonal leaves one
    candidate per image, its predecessor in row-major order, which counts iff its flattened <soi> position is not before
    the image's own flattened context start."""
    L = text_ids.shape[1]
    soi = _soi_list(text_ids, soi_token_id)
    n = len(soi)
    starts = [0] * n if nearest_bos_idxs is None else [int(v) for v in nearest_bos_idxs]
    feats = [torch.zeros((n, 1) + tuple(f.shape[1:]), dtype=f.dtype) for f in multiscale_features]
    mask = torch.zeros((n, 1), dtype=torch.long)
    for i in range(1, n):
        (r, _), (pr, pc) = soi[i], soi[i - 1]
        if r * L + starts[i] <= pr * L + pc:
            for src, dst in zip(multiscale_features, feats):
                dst[i, 0] = src[i - 1]
            mask[i, 0] = 1
    return feats, mask


def text_head_ref(sd, hidden, orig_vocab, prefix="text_decoder."):
    """TextDecoder.forward (decoders/decoder_text.py:152-157): ``head`` over the whole vocabulary, ``head_new`` added on
    the columns of the new ids; both linears carry a bias (:43-46)."""
    logits = hidden @ sd[prefix + "head.weight"].t() + sd[prefix + "head.bias"]
    new = hidden @ sd[prefix + "head_new.weight"].t() + sd[prefix + "head_new.bias"]
    logits[..., orig_vocab:] = logits[..., orig_vocab:] + new
    return logits


def gt_text_ids_ref(text_ids, attention_mask, st, ignore_prompt_token_offset=0):
    """``_prepare_gt_text_ids`` (mm_interleaved.py:342-406) for the default branch (no ``gt_text_ids``, dataset not in
    ``dataset_to_ignore_noimage_cond_loss``), written per element."""
    B, L = text_ids.shape
    gt = torch.full((B, L - 1), -100, dtype=text_ids.dtype)
    for b in range(B):
        off = ignore_prompt_token_offset if isinstance(ignore_prompt_token_offset, int) else ignore_prompt_token_offset[b]
        for t in range(1, L):
            tok, prev = int(text_ids[b, t]), int(text_ids[b, t - 1])
            if t < off:                                                               # :352-359
                continue
            if tok in (st["pad_token_id"], st["image_token_id"], st["bos_token_id"]):  # :389-394, :402-404
                continue
            if int(attention_mask[b, t]) == 0:                                        # :395
                continue
            if prev == st["bos_token_id"] and tok == st["soi_token_id"]:              # :397-400
                continue
            gt[b, t - 1] = tok
    return gt
