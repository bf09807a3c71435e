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


def sincos_pos_embed_1d(embed_dim: int, length: int) -> torch.Tensor:
    """(length, embed_dim) [sin | cos] table of ``get_1d_sincos_pos_embed_from_grid`` (utils/pos_embed.py:77-95) for
    positions 0..length-1 (float32 arithmetic like the numpy original)."""
    import numpy as np
    omega = np.arange(embed_dim // 2, dtype=np.float32)
    omega /= embed_dim / 2.0
    omega = 1.0 / 10000 ** omega
    out = np.einsum("m,d->md", np.arange(length, dtype=np.float32), omega)
    return torch.from_numpy(np.concatenate([np.sin(out), np.cos(out)], axis=1))


def soi_positions(text_ids: torch.Tensor, soi_token_id: int, n_images: int):
    """Row / column of the first ``n_images`` ``<soi>`` tokens in row-major order, without ``nonzero`` (no host sync:
    the image count is known from the image tensor)."""
    B, L = text_ids.shape
    flat = torch.where((text_ids == soi_token_id).reshape(-1), torch.arange(B * L, device=text_ids.device), B * L)
    flat = flat.sort().values[:n_images]
    return flat // L, flat % L


def context_features_for_image_decoder(context_features: torch.Tensor, text_ids: torch.Tensor, soi_token_id: int,
                                       context_feat_proj: nn.Module, seq_len: int, n_images: int,
                                       nearest_bos_idxs: Optional[torch.Tensor] = None, pad_to: Optional[int] = None):
    """``_prepare_context_features_for_image_decoder`` (mm_interleaved.py:254-304): for every image, the decoder hidden
    states from its nearest ``<bos>`` (default: position 0) up to and including its ``<soi>``, in REVERSED order (the
    ``<soi>`` state first), zero-padded to the longest context, through ``context_feat_proj`` (padding rows included,
    as in the reference) plus the 1-D sin-cos table.  Returns (features (B_I, L_max, C), mask (B_I, L_max) int64).
    ``pad_to`` fixes L_max (no host sync); None reproduces the reference's data-dependent ``max(context_lengths)``."""
    rows, cols = soi_positions(text_ids, soi_token_id, n_images)
    bos = torch.zeros_like(cols) if nearest_bos_idxs is None else nearest_bos_idxs.to(cols.dtype)
    lengths = cols - bos + 1
    L_max = int(lengths.max()) if pad_to is None else int(pad_to)
    t = torch.arange(L_max, device=text_ids.device)
    src = cols[:, None] - t[None, :]                                   # reversed walk from the <soi> position
    valid = t[None, :] < lengths[:, None]
    gathered = context_features[rows[:, None], src.clamp(min=0)]       # (B_I, L_max, C)
    per_image = torch.where(valid[..., None], gathered, torch.zeros((), dtype=gathered.dtype, device=gathered.device))
    pos = sincos_pos_embed_1d(context_features.shape[-1], seq_len).to(device=per_image.device, dtype=per_image.dtype)
    per_image = context_feat_proj(per_image) + pos[None, :L_max]
    return per_image, valid.to(cols.dtype)


def mmfs_features_for_image_decoder(multiscale_features: Sequence[torch.Tensor], text_ids: torch.Tensor, soi_token_id: int,
                                    nearest_bos_idxs: Optional[torch.Tensor] = None):
    """``_prepare_mmfs_features_for_image_decoder`` (mm_interleaved.py:306-340): the tril/triu pair keeps exactly one
    candidate per image -- the image right before it in row-major order -- and it is used iff its ``<soi>`` lies at or
    after the current image's context start (``row * L + nearest_bos``).  Returns ([ (B_I, 1, C, h, w) ], (B_I, 1))."""
    n = multiscale_features[0].shape[0]
    L = text_ids.shape[1]
    rows, cols = soi_positions(text_ids, soi_token_id, n)
    start = rows * L + (torch.zeros_like(cols) if nearest_bos_idxs is None else nearest_bos_idxs.to(cols.dtype))
    flat = rows * L + cols
    prev = torch.arange(n, device=text_ids.device) - 1
    use = (prev >= 0) & (start <= flat[prev.clamp(min=0)])             # image_context_mask[i, i-1]
    feats = []
    for f in multiscale_features:
        g = f[prev.clamp(min=0)] * use.view(-1, 1, 1, 1).to(f.dtype)
        feats.append(g[:, None])
    return feats, use.to(torch.long)[:, None]


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


class ImageDecoder(nn.Module):
    """``ImageDecoder`` (decoders/decoder_image.py:9-156) on this repo's modules: ``perceiver_resampler`` (Q-Former, 77
    queries of width 1024 over the per-image LLM context), ``neg_prompt_embeds`` and ``decoder`` = the SD UNet with
    its MMFS network (decoders/sd.py:20-140).  The VAE and the noise scheduler are diffusers objects that are not part
    of this repository: ``generate_images`` returns the denoised LATENTS (``output_type="latent"`` of the patched
    pipeline, sd.py:196-211) and applies ``vae_decode`` if the caller supplies one."""

    def __init__(self, perceiver_config=None, seq_len=77, embed_dim=1024, unet=None, mmfs_module=None, image_size=512,
                 base_seed=0):
        super().__init__()
        from .visual_tokenizer import PerceiverResampler
        self.perceiver_resampler = PerceiverResampler(**(perceiver_config or dict(num_queries=seq_len, hidden_size=embed_dim)))
        self.neg_prompt_embeds = nn.Parameter(torch.zeros(1, seq_len, embed_dim).normal_(0, 0.02))
        self.unet, self.mmfs_module = unet, mmfs_module
        self.image_size, self.base_seed = image_size, base_seed

    @torch.no_grad()
    def generate_images(self, context_features, context_attention_mask=None, mmfs_features=None, mmfs_mask=None,
                        num_inference_steps=30, guidance_scale=7.5, latents=None, vae_decode=None, **_):
        from .unet_sd import denoise_loop
        text_embeds = self.perceiver_resampler(encoder_hidden_states=context_features,
                                               encoder_attention_mask=context_attention_mask)[0]        # decoder_image.py:132-136
        neg = self.neg_prompt_embeds.to(text_embeds.dtype).expand_as(text_embeds)                       # :141-143
        n = text_embeds.shape[0]
        if latents is None:
            g = torch.Generator(device=text_embeds.device).manual_seed(self.base_seed)                  # sd.py:166-169
            side = self.image_size // 8
            latents = torch.randn((n, 4, side, side), generator=g, device=text_embeds.device, dtype=text_embeds.dtype)
        if latents.is_cuda:
            latents = latents.contiguous(memory_format=torch.channels_last)
        lat = denoise_loop(self.unet, latents, text_embeds, neg, mmfs_features, mmfs_mask, self.mmfs_module,
                           num_steps=num_inference_steps, guidance=guidance_scale)
        out = {"latents": lat}
        if vae_decode is not None:                                                                      # sd.py:212-216
            out["image"] = (vae_decode(lat.float() / 0.18215) / 2 + 0.5).clamp(0, 1)
        return out


class InterleavedForward(nn.Module):
    """``mm_decoder`` + ``text_decoder`` + ``soi_token`` of ``MMInterleaved`` with the forward path of
    ``MMInterleaved.forward`` up to the text logits.  Image embeddings / multi-scale maps come from the visual
    tokenizer (``visual_output`` dict with ``vis_embed`` and ``multiscale_features``, visual_tokenizer.py:96-101)."""

    def __init__(self, config: LlamaMMFSConfig, special_tokens=None, orig_vocab_size: int = 32000, seq_len: int = 2048,
                 image_decoder: Optional[nn.Module] = None):
        super().__init__()
        self.config = config
        self.special_token_dict = dict(DEFAULT_SPECIAL_TOKENS if special_tokens is None else special_tokens)
        self.mm_decoder = LlamaModel(config)
        self.text_decoder = TextHead(config.hidden_size, config.vocab_size, orig_vocab_size)
        self.soi_token = nn.Parameter(torch.zeros(1, config.hidden_size))
        self.spatial_shapes = list(config.spatial_shapes)
        self.context_feat_proj = nn.Linear(config.hidden_size, config.hidden_size)       # mm_interleaved.py:99
        self.seq_len = seq_len
        self.image_decoder = image_decoder                                                # ImageDecoder or None

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
    def generate_images(self, text_ids, visual_output, num_image_per_seq, max_num_image: int, attention_mask=None,
                        target_image_idxs=None, **kwargs):
        """``MMInterleaved.generate_images`` (mm_interleaved.py:520-596): decoder prefill over the interleaved context,
        per-image reversed context features (:254-304) and previous-image MMFS features (:306-340), optional selection
        of target images, then ``ImageDecoder.generate_images`` (Q-Former -> CFG denoise loop with the MMFS network)."""
        if self.image_decoder is None:
            raise RuntimeError("generate_images needs an image_decoder (ImageDecoder with a UNet and an MMFSNet)")
        st = self.special_token_dict
        mm_embeds, cross, feats = self.prepare(text_ids, visual_output, num_image_per_seq, max_num_image)
        hidden = self.mm_decoder(inputs_embeds=mm_embeds, attention_mask=attention_mask, vision_hidden_states=feats,
                                 cross_attention_mask=cross, use_cache=False, return_dict=True).last_hidden_state
        ms = visual_output["multiscale_features"]
        n_img = ms[0].shape[0]
        mmfs_features, mmfs_mask = mmfs_features_for_image_decoder(ms, text_ids, st["soi_token_id"])
        ctx, ctx_mask = context_features_for_image_decoder(hidden, text_ids, st["soi_token_id"], self.context_feat_proj,
                                                           self.seq_len, n_img, pad_to=kwargs.pop("context_pad_to", None))
        if target_image_idxs is not None:
            ctx, ctx_mask, mmfs_mask = (torch.index_select(t, 0, target_image_idxs) for t in (ctx, ctx_mask, mmfs_mask))
            mmfs_features = [torch.index_select(f, 0, target_image_idxs) for f in mmfs_features]
        out = self.image_decoder.generate_images(context_features=ctx, context_attention_mask=ctx_mask,
                                                 mmfs_features=mmfs_features, mmfs_mask=mmfs_mask, **kwargs)
        out.update(context_features=ctx, context_attention_mask=ctx_mask, mmfs_mask=mmfs_mask)
        return out

    @torch.no_grad()
    def generate_texts(self, text_ids, visual_output, num_image_per_seq, max_num_image: int, attention_mask=None,
                       max_new_tokens: int = 30, eos_token_id=2, pad_token_id: int = 0, static_cache: bool = True,
                       min_length: int = 0, repetition_penalty: float = 1.0, use_nucleus_sampling: bool = False,
                       top_p: float = 0.9, temperature: float = 1.0, generator: Optional[torch.Generator] = None,
                       num_beams: int = 1, length_penalty: float = 1.0, num_return_sequences: int = 1):
        """Text continuation over the interleaved context -- ``MMInterleaved.generate_texts``
        (mm_interleaved.py:598-664), which drives HF ``generate`` through ``CascadeLlamaForCausalLMWrapper``
        (models/utils/causal_lm_cascade.py:91-204): prefill on ``inputs_embeds`` with the image features, then one
        token per step over the KV cache, the last row of the cross-attention mask serving every new token
        (mmfs.py:161-162), ``position_ids = cumsum(mask) - 1`` (causal_lm_cascade.py:179-185).  Batches are expected
        left-padded (collator.py:337).  Greedy by default (num_beams=1, do_sample=False: the release inference
        config); the reference's other knobs that do not need beams are honoured with HF's semantics:
        ``repetition_penalty`` (scores of already generated ids divided / multiplied), ``min_length`` (every eos id
        is suppressed while fewer than ``min_length`` tokens were generated), several ``eos_token_id`` values (the
        reference passes [eos, soi]), and ``use_nucleus_sampling`` = temperature + top-p sampling.  ``num_beams > 1``
        runs HF-style beam search (``_beam_search`` below; the reference's captioning default is 5 beams) and returns
        (B * num_return_sequences, <= max_new_tokens) padded ids; otherwise (B, max_new_tokens) ids."""
        if num_beams > 1:
            if use_nucleus_sampling:
                raise NotImplementedError("beam-sample (num_beams > 1 with sampling) is not implemented")
            return self._beam_search(text_ids, visual_output, num_image_per_seq, max_num_image, attention_mask, max_new_tokens,
                                     eos_token_id, pad_token_id, min_length, repetition_penalty, num_beams, length_penalty,
                                     num_return_sequences)
        B, L = text_ids.shape
        if attention_mask is None:
            attention_mask = torch.ones((B, L), dtype=torch.long, device=text_ids.device)
        eos_ids = [] if eos_token_id is None else ([int(eos_token_id)] if isinstance(eos_token_id, int) else [int(e) for e in eos_token_id])
        mm_embeds, cross, feats = self.prepare(text_ids, visual_output, num_image_per_seq, max_num_image)
        position_ids = (attention_mask.long().cumsum(-1) - 1).clamp(min=0)
        # pre-allocated per-layer caches appended in place (the reference's cat-per-token re-copies every layer's cache)
        past = self.mm_decoder.static_cache(B, L + max_new_tokens, dtype=mm_embeds.dtype, device=mm_embeds.device) if static_cache else None
        out = self.mm_decoder(inputs_embeds=mm_embeds, attention_mask=attention_mask, position_ids=position_ids,
                              past_key_values=past, vision_hidden_states=feats, cross_attention_mask=cross, use_cache=True,
                              return_dict=True)
        past = out.past_key_values
        logits = self.text_decoder(out.last_hidden_state[:, -1:])
        new_ids = []
        finished = torch.zeros((B,), dtype=torch.bool, device=text_ids.device)
        mask = attention_mask
        last_cross = cross[:, -1:, :]
        pos = position_ids[:, -1:]
        for step_idx in range(max_new_tokens):
            scores = logits[:, -1].float()
            if repetition_penalty != 1.0 and new_ids:                    # HF RepetitionPenaltyLogitsProcessor
                prev = torch.stack(new_ids, dim=1)
                picked = scores.gather(1, prev)
                scores = scores.scatter(1, prev, torch.where(picked < 0, picked * repetition_penalty, picked / repetition_penalty))
            if step_idx < min_length and eos_ids:                        # HF MinLengthLogitsProcessor
                scores[:, eos_ids] = float("-inf")
            if use_nucleus_sampling:                                     # temperature, then top-p (HF warper order)
                scores = scores / temperature
                srt, idx = scores.sort(dim=-1, descending=False)
                drop = srt.softmax(-1).cumsum(-1) <= (1.0 - top_p)
                drop[:, -1] = False                                      # always keep the most likely token
                scores = scores.masked_fill(drop.scatter(1, idx, drop), float("-inf"))
                nxt = torch.multinomial(scores.softmax(-1), 1, generator=generator).squeeze(1)
            else:
                nxt = scores.argmax(-1)
            if eos_ids:
                nxt = torch.where(finished, torch.full_like(nxt, pad_token_id), nxt)
                for e in eos_ids:
                    finished = finished | (nxt == e)
            new_ids.append(nxt)
            mask = torch.cat([mask, torch.ones((B, 1), dtype=mask.dtype, device=mask.device)], dim=1)
            pos = pos + 1
            step = self.mm_decoder(inputs_embeds=self.mm_decoder.embed_tokens(nxt[:, None]), attention_mask=mask,
                                   position_ids=pos, past_key_values=past, vision_hidden_states=feats,
                                   cross_attention_mask=last_cross, use_cache=True, return_dict=True)
            past = step.past_key_values
            logits = self.text_decoder(step.last_hidden_state)
        return torch.stack(new_ids, dim=1)

    @torch.no_grad()
    def _beam_search(self, text_ids, visual_output, num_image_per_seq, max_num_image, attention_mask, max_new_tokens,
                     eos_token_id, pad_token_id, min_length, repetition_penalty, num_beams, length_penalty, num_return):
        """Beam search with the bookkeeping of HF ``GenerationMixin.beam_search`` + ``BeamSearchScorer`` (transformers
        4.31, the version the reference pins; ``early_stopping=False``, one beam group): log-softmax scores, logits
        processors on the log-probabilities, top 2*num_beams candidates per sequence, finished hypotheses ranked by
        ``sum_logprobs / len(generated) ** length_penalty``, a sequence is done once ``num_beams`` hypotheses are all
        at least as good as the best running beam could become.  The prompt is prefilled ONCE per sequence and its
        cache rows are replicated per beam; every step re-gathers the cache rows by beam index (``_reorder_cache``)."""
        from .llama_mmfs import StaticKV
        B, L = text_ids.shape
        nb, dev = num_beams, text_ids.device
        if attention_mask is None:
            attention_mask = torch.ones((B, L), dtype=torch.long, device=dev)
        eos_ids = [] if eos_token_id is None else ([int(eos_token_id)] if isinstance(eos_token_id, int) else [int(e) for e in eos_token_id])
        mm_embeds, cross, feats = self.prepare(text_ids, visual_output, num_image_per_seq, max_num_image)
        position_ids = (attention_mask.long().cumsum(-1) - 1).clamp(min=0)
        pre = self.mm_decoder.static_cache(B, L, dtype=mm_embeds.dtype, device=dev)
        out = self.mm_decoder(inputs_embeds=mm_embeds, attention_mask=attention_mask, position_ids=position_ids,
                              past_key_values=pre, vision_hidden_states=feats, cross_attention_mask=cross, use_cache=True,
                              return_dict=True)
        rep = torch.arange(B, device=dev).repeat_interleave(nb)                    # beam row -> sequence
        past = self.mm_decoder.static_cache(B * nb, L + max_new_tokens, dtype=mm_embeds.dtype, device=dev)
        for dst, src in zip(past, pre):
            dst.k[:, :L].copy_(src.k.index_select(0, rep)); dst.v[:, :L].copy_(src.v.index_select(0, rep)); dst.length = L
        del pre
        logits = self.text_decoder(out.last_hidden_state[:, -1:]).index_select(0, rep)
        feats_b, last_cross = feats.index_select(0, rep), cross[:, -1:, :].index_select(0, rep)
        mask, pos = attention_mask.index_select(0, rep), position_ids[:, -1:].index_select(0, rep)

        beam_scores = torch.zeros((B, nb), dtype=torch.float32, device=dev)
        beam_scores[:, 1:] = -1e9
        beam_scores = beam_scores.view(-1)
        seqs = torch.zeros((B * nb, 0), dtype=torch.long, device=dev)              # generated ids per beam row
        hyps = [[] for _ in range(B)]                                              # per sequence: (score, ids list)
        worst = [1e9] * B
        done = [False] * B

        def add_hyp(b, ids, sum_logprobs):
            score = sum_logprobs / (max(len(ids), 1) ** length_penalty)
            if len(hyps[b]) < nb or score > worst[b]:
                hyps[b].append((score, ids))
                if len(hyps[b]) > nb:
                    hyps[b].remove(min(hyps[b], key=lambda h: h[0]))
                worst[b] = min(h[0] for h in hyps[b])

        for step_idx in range(max_new_tokens):
            scores = torch.log_softmax(logits[:, -1].float(), dim=-1)
            if repetition_penalty != 1.0 and seqs.shape[1] > 0:
                picked = scores.gather(1, seqs)
                scores = scores.scatter(1, seqs, torch.where(picked < 0, picked * repetition_penalty, picked / repetition_penalty))
            if step_idx < min_length and eos_ids:
                scores[:, eos_ids] = float("-inf")
            V = scores.shape[-1]
            cand = (scores + beam_scores[:, None]).view(B, nb * V)
            top_s, top_i = cand.topk(2 * nb, dim=1, largest=True, sorted=True)
            top_s_h, top_i_h, seqs_h = top_s.tolist(), top_i.tolist(), seqs.tolist()   # one host round trip per step
            cur_len = seqs.shape[1] + 1
            nxt_scores = [[0.0] * nb for _ in range(B)]
            nxt_tokens = [[pad_token_id] * nb for _ in range(B)]
            nxt_rows = [[b * nb] * nb for b in range(B)]
            for b in range(B):
                if done[b]:
                    continue
                k = 0
                for rank, (sc, idx) in enumerate(zip(top_s_h[b], top_i_h[b])):
                    row, tok = b * nb + idx // V, idx % V
                    if tok in eos_ids:
                        if rank >= nb:
                            continue
                        add_hyp(b, seqs_h[row], sc)
                    else:
                        nxt_scores[b][k], nxt_tokens[b][k], nxt_rows[b][k] = sc, tok, row
                        k += 1
                    if k == nb:
                        break
                if len(hyps[b]) >= nb and worst[b] >= top_s_h[b][0] / (cur_len ** length_penalty):
                    done[b] = True
            beam_scores = torch.tensor(nxt_scores, dtype=torch.float32, device=dev).view(-1)
            tok_t = torch.tensor(nxt_tokens, dtype=torch.long, device=dev).view(-1)
            row_t = torch.tensor(nxt_rows, dtype=torch.long, device=dev).view(-1)
            seqs = torch.cat([seqs.index_select(0, row_t), tok_t[:, None]], dim=1)
            if all(done) or step_idx == max_new_tokens - 1:
                break
            for c in past:                                                        # _reorder_cache
                n = c.length
                c.k[:, :n].copy_(c.k.index_select(0, row_t)[:, :n]); c.v[:, :n].copy_(c.v.index_select(0, row_t)[:, :n])
            mask = torch.cat([mask.index_select(0, row_t), torch.ones((B * nb, 1), dtype=mask.dtype, device=dev)], dim=1)
            pos = pos.index_select(0, row_t) + 1
            step = self.mm_decoder(inputs_embeds=self.mm_decoder.embed_tokens(tok_t[:, None]), attention_mask=mask, position_ids=pos,
                                   past_key_values=past, vision_hidden_states=feats_b, cross_attention_mask=last_cross,
                                   use_cache=True, return_dict=True)
            logits = self.text_decoder(step.last_hidden_state)

        # finalize: running beams of unfinished sequences become hypotheses; best `num_return` per sequence
        seqs_h, bs_h = seqs.tolist(), beam_scores.tolist()
        for b in range(B):
            if not done[b]:
                for j in range(nb):
                    add_hyp(b, seqs_h[b * nb + j], bs_h[b * nb + j])
        best = []
        for b in range(B):
            ranked = sorted(hyps[b], key=lambda h: h[0])
            for _ in range(num_return):
                best.append(ranked.pop()[1])
        width = min(max(len(x) for x in best) + 1, max_new_tokens)
        out_ids = torch.full((len(best), width), pad_token_id, dtype=torch.long)
        for i, x in enumerate(best):
            out_ids[i, :len(x)] = torch.tensor(x, dtype=torch.long)
            if len(x) < width and eos_ids:
                out_ids[i, len(x)] = eos_ids[0]
        return out_ids.to(dev)
