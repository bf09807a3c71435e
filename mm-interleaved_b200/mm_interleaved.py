"""Top-level glue of the interleaved forward: embed splice, image-visibility mask, MMFS feature
packing, decoder prefill and text head -- the body of ``MMInterleaved.forward`` up to the logits
(mm_interleaved/models/mm_interleaved.py:121-252, 408-455) and ``TextDecoder.forward``
(models/decoders/decoder_text.py:140-163).

Same semantics as the reference helpers, but written for the device: no Python loops over the batch,
no ``.nonzero()`` / ``.max()`` host synchronisations, no per-sample slicing -- everything is a handful
of tensor ops whose shapes are known from the (static) maximum image count.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.nn.functional as F
from torch import nn

from .llama_mmfs import LlamaMMFSConfig, LlamaModel

# special-token convention of the reference tokenizer (custom_datasets/wds_utils.py:186-215): the two
# added ids follow the 32000 Llama ids
DEFAULT_SPECIAL_TOKENS = dict(bos_token_id=1, image_token_id=32000, soi_token_id=32001)


def splice_image_embeds(text_embeds, text_ids, image_embeds, soi_token, image_token_id, soi_token_id):
    """Steps 3 of ``_prepare_mm_embeds`` (mm_interleaved.py:144-170): the k-th ``<image>`` slot (row-major over the
    batch) receives the k-th row of ``image_embeds``; the learnable ``soi_token`` is added at every ``<soi>``."""
    is_img = (text_ids == image_token_id).unsqueeze(-1)
    out = text_embeds.to(image_embeds.dtype).masked_scatter(is_img, image_embeds.reshape(-1, image_embeds.shape[-1]))
    is_soi = (text_ids == soi_token_id).unsqueeze(-1).to(out.dtype)
    return out + is_soi * soi_token.to(out.dtype).view(1, 1, -1)


def cross_attention_mask_from_ids(text_ids, max_num_image: int, bos_token_id: int, soi_token_id: int,
                                  num_image_per_seq: Optional[torch.Tensor] = None):
    """(B, L, N) float 0/1: image n of a sequence is visible to token t iff ``soi_n + 1 > nearest_bos(t)`` and
    ``soi_n + 1 <= t`` (mm_interleaved.py:192-221).  Slots past a sequence's image count are never visible."""
    B, L = text_ids.shape
    ar = torch.arange(L, device=text_ids.device)
    soi_pos = torch.where(text_ids == soi_token_id, ar[None, :], L + 1).sort(dim=1).values[:, :max_num_image]
    if soi_pos.shape[1] < max_num_image:
        soi_pos = torch.nn.functional.pad(soi_pos, (0, max_num_image - soi_pos.shape[1]), value=L + 1)
    valid = soi_pos <= L
    if num_image_per_seq is not None:
        valid = valid & (torch.arange(max_num_image, device=text_ids.device)[None, :] < num_image_per_seq[:, None])
    img_pos = torch.where(valid, soi_pos + 1, torch.full_like(soi_pos, -1))            # (B, N)
    nearest_bos = torch.where(text_ids == bos_token_id, ar[None, :], -1).cummax(dim=1).values   # (B, L)
    vis = (img_pos[:, None, :] > nearest_bos[:, :, None]) & (img_pos[:, None, :] <= ar[None, :, None]) & \
          (img_pos[:, None, :] != -1)
    return vis.float()


def pack_mmfs_features(multiscale_features: Sequence[torch.Tensor], spatial_shapes: Sequence[int],
                       num_image_per_seq: torch.Tensor, max_num_image: int):
    """(B, N, sum(h*w), C): the maps whose side is in ``spatial_shapes``, zero-padded per sequence to ``max_num_image``
    images and flattened level by level (mm_interleaved.py:223-250)."""
    feats = [f for f in multiscale_features if int(f.shape[-1]) in spatial_shapes]
    B = num_image_per_seq.shape[0]
    first = torch.cumsum(num_image_per_seq, 0) - num_image_per_seq                      # first image of each sequence
    n_tot = feats[0].shape[0]
    img = torch.arange(n_tot, device=feats[0].device)
    seq_of = torch.bucketize(img, torch.cumsum(num_image_per_seq, 0), right=True)
    dest = seq_of * max_num_image + (img - first[seq_of])
    packed = []
    for f in feats:
        n, c, h, w = f.shape
        flat = f.flatten(2).transpose(1, 2)                                              # (n, hw, C)
        buf = flat.new_zeros((B * max_num_image, h * w, c))
        buf.index_copy_(0, dest, flat)
        packed.append(buf.view(B, max_num_image, h * w, c))
    return torch.cat(packed, dim=2)


class TextHead(nn.Module):
    """``TextDecoder`` (decoders/decoder_text.py): ``head`` over the original vocabulary plus ``head_new`` for the
    added ids, summed on the tail columns (:155-157).  State-dict names match the reference."""

    def __init__(self, hidden_size: int, vocab_size: int, orig_vocab_size: int):
        super().__init__()
        self.orig_txt_vocab_size = orig_vocab_size
        self.head = nn.Linear(hidden_size, vocab_size, bias=False)
        self.head_new = nn.Linear(hidden_size, vocab_size - orig_vocab_size, bias=False)

    _PAD = 128   # a vocabulary of 32002+ rows is not a multiple of 8: cuBLAS drops to an unaligned legacy kernel (5x slower)

    def _fused_weight(self):
        """head + head_new folded into one matrix, rows zero-padded to a multiple of 128 (inference only)."""
        key = tuple((w.data_ptr(), w._version, w.dtype, w.device) for w in (self.head.weight, self.head_new.weight))
        if getattr(self, "_fused", None) is None or self._fused[0] != key:
            V, C = self.head.weight.shape
            Vp = (V + self._PAD - 1) // self._PAD * self._PAD
            w = self.head.weight.new_zeros((Vp, C))
            w[:V] = self.head.weight.detach()
            w[self.orig_txt_vocab_size:V] += self.head_new.weight.detach()
            self._fused = (key, w)
        return self._fused[1]

    def forward(self, hidden_states):
        if torch.is_grad_enabled() and (self.head.weight.requires_grad or self.head_new.weight.requires_grad):
            logits = self.head(hidden_states)
            logits[..., self.orig_txt_vocab_size:] += self.head_new(hidden_states)
            return logits
        return F.linear(hidden_states, self._fused_weight())[..., :self.head.weight.shape[0]]


class InterleavedForward(nn.Module):
    """``mm_decoder`` + ``text_decoder`` + ``soi_token`` of ``MMInterleaved`` with the forward path of
    ``MMInterleaved.forward`` up to the text logits.  Image embeddings / multi-scale maps come from the visual
    tokenizer (``visual_output`` dict with ``vis_embed`` and ``multiscale_features``, visual_tokenizer.py:96-101)."""

    def __init__(self, config: LlamaMMFSConfig, special_tokens=None, orig_vocab_size: int = 32000):
        super().__init__()
        self.config = config
        self.special_token_dict = dict(DEFAULT_SPECIAL_TOKENS if special_tokens is None else special_tokens)
        self.mm_decoder = LlamaModel(config)
        self.text_decoder = TextHead(config.hidden_size, config.vocab_size, orig_vocab_size)
        self.soi_token = nn.Parameter(torch.zeros(1, config.hidden_size))
        self.spatial_shapes = list(config.spatial_shapes)

    def prepare(self, text_ids, visual_output, num_image_per_seq, max_num_image: int):
        st = self.special_token_dict
        embeds = self.mm_decoder.embed_tokens(text_ids)
        mm_embeds = splice_image_embeds(embeds, text_ids, visual_output["vis_embed"], self.soi_token,
                                        st["image_token_id"], st["soi_token_id"])
        cross = cross_attention_mask_from_ids(text_ids, max_num_image, st["bos_token_id"], st["soi_token_id"],
                                              num_image_per_seq)
        feats = pack_mmfs_features(visual_output["multiscale_features"], self.spatial_shapes, num_image_per_seq,
                                   max_num_image)
        return mm_embeds, cross, feats

    def forward(self, text_ids, visual_output, num_image_per_seq, max_num_image: int, attention_mask=None):
        mm_embeds, cross, feats = self.prepare(text_ids, visual_output, num_image_per_seq, max_num_image)
        out = self.mm_decoder(inputs_embeds=mm_embeds, attention_mask=attention_mask, vision_hidden_states=feats,
                              cross_attention_mask=cross, use_cache=False, return_dict=True)
        return self.text_decoder(out.last_hidden_state)

    @torch.no_grad()
    def generate_texts(self, text_ids, visual_output, num_image_per_seq, max_num_image: int, attention_mask=None,
                       max_new_tokens: int = 30, eos_token_id: Optional[int] = 2, pad_token_id: int = 0):
        """Greedy text continuation over the interleaved context -- the deterministic setting (num_beams=1,
        do_sample=False) of ``MMInterleaved.generate_texts`` (mm_interleaved.py:598-664), which drives HF ``generate``
        through ``CascadeLlamaForCausalLMWrapper`` (models/utils/causal_lm_cascade.py:91-204): prefill on
        ``inputs_embeds`` with the image features, then one token per step over the KV cache, the last row of the
        cross-attention mask serving every new token (mmfs.py:161-162), ``position_ids = cumsum(mask) - 1``
        (causal_lm_cascade.py:179-185).  Batches are expected left-padded (collator.py:337).  Returns (B, n_new) ids."""
        B, L = text_ids.shape
        if attention_mask is None:
            attention_mask = torch.ones((B, L), dtype=torch.long, device=text_ids.device)
        mm_embeds, cross, feats = self.prepare(text_ids, visual_output, num_image_per_seq, max_num_image)
        position_ids = (attention_mask.long().cumsum(-1) - 1).clamp(min=0)
        out = self.mm_decoder(inputs_embeds=mm_embeds, attention_mask=attention_mask, position_ids=position_ids,
                              vision_hidden_states=feats, cross_attention_mask=cross, use_cache=True, return_dict=True)
        past = out.past_key_values
        logits = self.text_decoder(out.last_hidden_state[:, -1:])
        new_ids = []
        finished = torch.zeros((B,), dtype=torch.bool, device=text_ids.device)
        mask = attention_mask
        last_cross = cross[:, -1:, :]
        pos = position_ids[:, -1:]
        for _ in range(max_new_tokens):
            nxt = logits[:, -1].argmax(-1)
            if eos_token_id is not None:
                nxt = torch.where(finished, torch.full_like(nxt, pad_token_id), nxt)
                finished = finished | (nxt == eos_token_id)
            new_ids.append(nxt)
            mask = torch.cat([mask, torch.ones((B, 1), dtype=mask.dtype, device=mask.device)], dim=1)
            pos = pos + 1
            step = self.mm_decoder(inputs_embeds=self.mm_decoder.embed_tokens(nxt[:, None]), attention_mask=mask,
                                   position_ids=pos, past_key_values=past, vision_hidden_states=feats,
                                   cross_attention_mask=last_cross, use_cache=True, return_dict=True)
            past = step.past_key_values
            logits = self.text_decoder(step.last_hidden_state)
        return torch.stack(new_ids, dim=1)
