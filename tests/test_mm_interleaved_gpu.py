"""GPU tests of the reference-facing top-level surface ``MMInterleaved`` (mm_interleaved/models/mm_interleaved.py):
``forward(text_ids, image_tensors, num_image_per_seq, attention_mask)`` (:408-518), ``generate(mode=..., **batch)``
(:745-763) with ``generate_texts`` (:598-664), ``generate_scores`` (:666-743) and ``generate_images`` (:520-596),
driven with the batch keys the reference collators produce (collator.py:358-371), on a tiny configuration in fp32.

The visual tokenizer's own arithmetic is pinned in test_visual_tokenizer_gpu.py / test_oracle_tokenizer*.py; here its
output is taken as given and everything downstream (embed splice, visibility mask, feature packing, decoder, text head
with biases, target construction, loss, greedy decode, option scoring) is compared with the CPU oracle restatement.
Tolerance: logits |err| <= 1e-3 |ref| + 1e-4 (north-star 1e-3 rel); tokens exact; loss 1e-4 rel."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"

from oracle.glue import (cross_attention_mask_ref, gt_text_ids_ref, pack_mmfs_features_ref,  # noqa: E402
                         prepare_mm_embeds_ref, text_head_ref)
from oracle.llama import llama_model_ref  # noqa: E402
from tests.golden.make_golden import LLAMA_TINY, seeded_state_dict  # noqa: E402

ST = dict(bos_token_id=1, eos_token_id=2, pad_token_id=0, soi_token_id=62, image_token_id=63)
N_TOK = 3


def _build(with_image_decoder=False):
    import mm_interleaved_b200 as m
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(0)
    vt_cfg = dict(clip_config=m.visual_tokenizer.CLIPVisionConfigLite(hidden_size=512, intermediate_size=512, num_hidden_layers=4,
                                                                     num_attention_heads=4, image_size=56, patch_size=14),
                  perceiver_config=dict(num_queries=N_TOK, hidden_size=192, encoder_hidden_size=512, cross_attention_frequency=2,
                                        num_hidden_layers=2, num_attention_heads=3, intermediate_size=384,
                                        qk_normalization=True), grid_size=4)
    img_cfg = None
    if with_image_decoder:
        from mm_interleaved_b200 import unet_sd
        unet = unet_sd.UNet2DConditionModel(block_out_channels=(64, 128), layers_per_block=1, attention_head_dim=(2, 4),
                                            cross_attention_dim=96)
        net = m.MMFSNet(LLAMA_TINY["image_embed_dim"], (64, 128), 1, downsample_factor=2, spatial_shapes=[16, 8, 4, 2])
        img_cfg = dict(perceiver_config=dict(num_queries=7, hidden_size=96, encoder_hidden_size=LLAMA_TINY["hidden_size"],
                                             num_hidden_layers=2, num_attention_heads=4, intermediate_size=192,
                                             cross_attention_frequency=1, qk_normalization=True),
                       seq_len=7, embed_dim=96, unet=unet, mmfs_module=net, image_size=128, sd_base_seed=3)
    model = m.MMInterleaved(llm_config=dict(LLAMA_TINY, vocab_size=62), txt_vocab_size=64, seq_len=32, special_token_dict=ST,
                            visual_tokenizer_config=vt_cfg, image_decoder_config=img_cfg,
                            image_embed_dim=LLAMA_TINY["image_embed_dim"], cross_attention_frequency=2,
                            spatial_shapes=LLAMA_TINY["spatial_shapes"])
    sd = model.state_dict()
    llm = seeded_state_dict({k: v for k, v in sd.items() if k.split(".")[0] in ("mm_decoder", "text_decoder", "soi_token",
                                                                               "context_feat_proj")}, seed=2024)
    llm["text_decoder.head.weight"][60:] = 0          # special ids are never the arg-max (greedy tests)
    llm["text_decoder.head_new.weight"].zero_()
    llm["text_decoder.head.bias"][60:] = -1.0
    llm["text_decoder.head.weight"][:3] = 0           # ... nor pad / bos / eos (the oracle loop has no stopping rule)
    llm["text_decoder.head.bias"][:3] = -1.0
    llm["text_decoder.head_new.bias"].zero_()
    sd.update(llm)
    with torch.no_grad():
        sd["visual_tokenizer.proj.weight"] = torch.randn_like(sd["visual_tokenizer.proj.weight"]) * 0.05   # 1e-3 init
    model.load_state_dict(sd)
    return model.to(DEV).eval(), sd


def _batch():
    g = torch.Generator().manual_seed(11)
    L = 20
    ids = torch.randint(3, 60, (2, L), generator=g)
    ids[:, 0] = 1
    for r, c in ((0, 2), (0, 10), (1, 5)):
        ids[r, c] = ST["soi_token_id"]
        ids[r, c + 1:c + 1 + N_TOK] = ST["image_token_id"]
    images = torch.rand((3, 3, 56, 56), generator=g)
    return ids, images, torch.tensor([2, 1]), torch.ones((2, L), dtype=torch.long)


def _oracle_hidden(model, sd, ids, images, nimg, mask=None):
    with torch.no_grad():
        vis = model.visual_tokenizer(images.to(DEV))
    vis = {"vis_embed": vis["vis_embed"].cpu(), "multiscale_features": [f.cpu() for f in vis["multiscale_features"]]}
    dec = {k[len("mm_decoder."):]: v for k, v in sd.items() if k.startswith("mm_decoder.")}
    ocfg = dict(eps=LLAMA_TINY["rms_norm_eps"], n_heads=LLAMA_TINY["num_attention_heads"],
                n_layers=LLAMA_TINY["num_hidden_layers"], spatial_shapes=[(s, s) for s in LLAMA_TINY["spatial_shapes"]])
    feats = pack_mmfs_features_ref(vis["multiscale_features"], LLAMA_TINY["spatial_shapes"], nimg)
    cross = cross_attention_mask_ref(ids, nimg, ST["bos_token_id"], ST["soi_token_id"])

    def run(cur, n_extra=0):
        emb = torch.nn.functional.embedding(cur, dec["embed_tokens.weight"])
        emb = prepare_mm_embeds_ref(emb, cur, vis["vis_embed"], sd["soi_token"], ST["image_token_id"], ST["soi_token_id"])
        c = torch.cat([cross] + [cross[:, -1:]] * n_extra, dim=1)
        am = torch.ones_like(cur) if mask is None else torch.cat([mask, torch.ones((cur.shape[0], n_extra), dtype=mask.dtype)], 1)
        return llama_model_ref(dec, emb, am, None, feats, c, ocfg)[0]
    return run


def test_forward_signature_logits_and_text_loss():
    model, sd = _build()
    ids, images, nimg, mask = _batch()
    with torch.no_grad():
        out = model(text_ids=ids.to(DEV), image_tensors=images.to(DEV), num_image_per_seq=nimg.to(DEV),
                    attention_mask=mask.to(DEV), meta={"dataset_name": "synthetic"})
    assert {"loss", "loss_txt", "text_logits", "multiscale_features"} <= set(out.keys())
    hid = _oracle_hidden(model, sd, ids, images, nimg)(ids)
    want = text_head_ref(sd, hid, 62)
    got = out["text_logits"].float().cpu()
    assert bool(((got - want).abs() <= 1e-3 * want.abs() + 1e-4).all()), float((got - want).abs().max())
    gt = gt_text_ids_ref(ids, mask, ST)
    assert int((gt != -100).sum()) > 10
    loss = torch.nn.functional.cross_entropy(want[:, :-1].transpose(1, 2), gt)
    assert abs(float(out["loss_txt"]) - float(loss)) <= 1e-4 * abs(float(loss)) + 1e-5
    assert abs(float(out["loss"]) - float(loss)) <= 1e-4 * abs(float(loss)) + 1e-5       # loss_txt_weight = 1


def test_generate_texts_mode_matches_oracle_greedy_loop():
    model, sd = _build()
    ids, images, nimg, mask = _batch()
    batch = dict(text_ids=ids.to(DEV), image_tensors=images.to(DEV), num_image_per_seq=nimg.to(DEV),
                 attention_mask=mask.to(DEV), meta=None)
    n_new = 5
    out = model.generate(mode="generate_texts", **batch, num_beams=1, max_length=n_new, min_length=0)
    got = out["text_ids"].cpu()
    assert got.shape == (2, n_new)
    run = _oracle_hidden(model, sd, ids, images, nimg)
    cur, want = ids.clone(), []
    for step in range(n_new):
        nxt = text_head_ref(sd, run(cur, step)[:, -1], 62).argmax(-1)
        want.append(nxt)
        cur = torch.cat([cur, nxt[:, None]], 1)
    assert torch.equal(got, torch.stack(want, 1)), (got, torch.stack(want, 1))
    # the reference's defaults (5 beams, min_length 8, eos = [eos, soi]) run through the same entry point
    out5 = model.generate(mode="generate_vqa", **batch, max_length=4, min_length=2)
    assert out5["text_ids"].shape[0] == 2 and out5["text_ids"].shape[1] <= 4
    with pytest.raises(NotImplementedError):
        model.generate(mode="no_such_mode", **batch)


def test_generate_scores_matches_oracle_log_likelihoods():
    model, sd = _build()
    g = torch.Generator().manual_seed(5)
    ctx = []
    for L in (9, 9):
        t = torch.randint(3, 60, (L,), generator=g)
        t[0] = 1
        t[2] = ST["soi_token_id"]
        t[3:3 + N_TOK] = ST["image_token_id"]
        ctx.append(t)
    images = torch.rand((2, 3, 56, 56), generator=g)
    opts = [torch.randint(3, 60, (5, 4), generator=g), torch.randint(3, 60, (3, 4), generator=g)]
    opt_masks = [torch.ones((5, 4), dtype=torch.long), torch.ones((3, 4), dtype=torch.long)]
    opt_masks[0][1, 2:] = 0
    out = model.generate(mode="generate_scores", text_ids=[t.to(DEV) for t in ctx], image_tensors=images.to(DEV),
                         num_image_per_seq=torch.ones((2, 1), dtype=torch.long, device=DEV),
                         attention_mask=[torch.ones_like(t).to(DEV) for t in ctx],
                         options_ids=[opts[0].to(DEV), opts[0].to(DEV)],
                         options_attn_masks=[opt_masks[0].to(DEV), opt_masks[0].to(DEV)])
    got = out["scores"].cpu()
    assert got.shape == (2, 1, 5)
    for i in range(2):
        full = torch.cat([ctx[i][None].expand(5, -1), opts[0]], 1)
        hid = _oracle_hidden(model, sd, full, images[[i]].expand(5, -1, -1, -1), torch.ones(5, dtype=torch.long))(full)
        logits = text_head_ref(sd, hid, 62)[:, len(ctx[i]) - 1:-1]
        logp = torch.log_softmax(logits, -1).gather(-1, opts[0][..., None]).squeeze(-1)
        want = (logp * opt_masks[0]).sum(-1)
        assert bool(((got[i, 0] - want).abs() <= 1e-3 * want.abs() + 1e-3).all()), (got[i, 0], want)


def test_generate_images_mode_runs_through_the_reference_surface():
    model, _ = _build(with_image_decoder=True)
    ids, images, nimg, mask = _batch()
    batch = dict(text_ids=ids.to(DEV), image_tensors=images.to(DEV), num_image_per_seq=nimg.to(DEV),
                 attention_mask=mask.to(DEV), meta=None)
    out = model.generate(mode="generate_images", **batch, num_inference_steps=3, guidance_scale=3.0)
    assert out["image"].shape == (3, 4, 16, 16) and bool(torch.isfinite(out["image"]).all())
    again = model.generate(mode="generate_images", **batch, num_inference_steps=3, guidance_scale=3.0)
    scale = float(out["image"].abs().max())
    assert float((out["image"] - again["image"]).abs().max()) <= 1e-5 * scale   # seeded generator (sd.py:166-169)
    sel = model.generate(mode="generate_images", **batch, num_inference_steps=2, target_image_idxs=torch.tensor([1], device=DEV))
    assert sel["image"].shape[0] == 1
    # a VAE hand-off: decoded, rescaled to [0, 1] (sd.py:212-215)
    model.image_decoder.decoder.vae_decode = lambda z: torch.tanh(z.mean(1, keepdim=True).repeat(1, 3, 1, 1))
    img = model.generate(mode="generate_images", **batch, num_inference_steps=2)["image"]
    assert img.shape == (3, 3, 16, 16) and float(img.min()) >= 0.0 and float(img.max()) <= 1.0
