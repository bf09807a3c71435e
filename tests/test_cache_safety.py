"""Activation-derived caches must never outlive their source tensor (round-1 defect: caches keyed on
``data_ptr/_version`` returned the PREVIOUS forward's RMSNorm(vision) / value_proj when the caching allocator reused the
address).  CPU: the cache object itself.  GPU: two consecutive forwards with DIFFERENT images -- including a
``del`` + re-allocation that provably lands on the same address -- each against the oracle
(modeling_llama_mmfs.py:352-353 and ops/modules/mmfs.py:165-172 recompute per call; sd_mmfs.py:121)."""
import gc

import pytest
import torch


def test_source_cache_identity_version_and_lifetime():
    from mm_interleaved_b200._cache import SourceCache
    c = SourceCache()
    a = torch.zeros(4)
    assert c.get(a) is None
    c.put(a, "A", extra=(1,))
    assert c.get(a, (1,)) == "A"
    assert c.get(a, (2,)) is None                       # a weight version changed
    a.add_(1)                                           # in-place write bumps _version
    assert c.get(a, (1,)) is None
    c.put(a, "A2", (1,))
    b = a.clone()                                       # equal contents / shape, different object
    assert c.get(b, (1,)) is None
    v = a.view(4)                                       # same storage, different Python object: a miss (safe side)
    assert c.get(v, (1,)) is None
    del a, v
    gc.collect()
    d = torch.zeros(4)                                  # may or may not reuse the address: must miss either way
    assert c.get(d, (1,)) is None
    # multi-source form (MMFSNet's list of feature maps)
    xs = [torch.zeros(2), torch.zeros(3)]
    c.put(xs, "X")
    assert c.get(xs) == "X" and c.get(list(xs)) == "X"
    assert c.get([xs[0], torch.zeros(3)]) is None and c.get(xs[:1]) is None


def test_clear_activation_caches_walks_modules():
    import mm_interleaved_b200 as m
    from mm_interleaved_b200._cache import SourceCache, clear_activation_caches
    from tests.golden.make_golden import MMFSNET_TINY
    net = m.MMFSNet(**MMFSNET_TINY)
    caches = [v for mod in net.modules() for v in vars(mod).values() if isinstance(v, SourceCache)]
    assert len(caches) >= 3
    t = torch.zeros(1)
    for c in caches:
        c.put(t, 1)
    clear_activation_caches(net)
    assert all(c.get(t) is None for c in caches)


# ------------------------------------------------------------------------------------------------ GPU
def _fresh_same_address(make, old_ptr, tries=8):
    """Allocate through the caching allocator until the new tensor lands on ``old_ptr`` (it does on the first try when
    the previous tensor of that size was just freed); returns the tensor and whether the address matched."""
    t = None
    for _ in range(tries):
        t = make()
        if t.data_ptr() == old_ptr:
            return t, True
        del t
    return make(), False


@pytest.mark.gpu
def test_llama_mmfs_two_forwards_with_different_images_match_oracle():
    import mm_interleaved_b200 as m
    from oracle.llama import llama_model_ref
    from tests.golden.make_golden import LLAMA_TINY, llama_inputs, seeded_state_dict
    torch.backends.cuda.matmul.allow_tf32 = False
    model = m.LlamaModel(m.LlamaMMFSConfig(**LLAMA_TINY))
    sd = seeded_state_dict(model.state_dict(), seed=4242)
    model.load_state_dict(sd)
    model = model.cuda().eval()
    cfg = dict(eps=LLAMA_TINY["rms_norm_eps"], n_heads=LLAMA_TINY["num_attention_heads"],
               n_layers=LLAMA_TINY["num_hidden_layers"], spatial_shapes=[(s, s) for s in LLAMA_TINY["spatial_shapes"]])
    embeds, vision_a, attn_mask, position_ids, cross = llama_inputs(LLAMA_TINY, 2, 12, 2, seed=99)
    vision_b = torch.randn(vision_a.shape, generator=torch.Generator().manual_seed(1234)) * 1.5 + 0.25
    valid = attn_mask.bool()

    def run(vision_dev):
        with torch.no_grad():
            return model(inputs_embeds=embeds.cuda(), attention_mask=attn_mask.cuda(), position_ids=position_ids.cuda(),
                         vision_hidden_states=vision_dev, cross_attention_mask=cross.cuda(), use_cache=False
                         ).last_hidden_state.cpu()

    def check(out, vision):
        ref, _ = llama_model_ref(sd, embeds, attn_mask, position_ids, vision, cross, cfg)
        err = (out - ref).abs()[valid]
        assert bool((err <= 1e-3 * ref[valid].abs() + 2e-5).all()), float(err.max())
        return ref

    va = vision_a.cuda()
    ptr = va.data_ptr()
    ref_a = check(run(va), vision_a)
    del va
    vb, same = _fresh_same_address(lambda: vision_b.cuda(), ptr)
    if not same:   # allocator-dependent; the 4-call test below covers the natural pattern as well
        import warnings
        warnings.warn("allocator did not reuse the freed block; address-reuse variant not exercised in this run")
    ref_b = check(run(vb), vision_b)
    assert float((ref_a - ref_b).abs()[valid].max()) > 1e-2          # the two image sets really give different outputs
    # same object again (what every decode step of one generate call does): cache hit, identical result
    assert torch.equal(run(vb), run(vb))
    # in-place refill of the same tensor bumps its version: recomputed
    vb.copy_(vision_a.cuda())
    check(run(vb), vision_a)


@pytest.mark.gpu
def test_interleaved_forward_twice_with_different_images_matches_oracle():
    """``InterleavedForward.forward`` builds a fresh packed feature tensor per call and frees it on return -- the exact
    pattern that hit the stale cache.  Logits of call 2 (other images) must match the oracle for call 2."""
    from oracle.glue import cross_attention_mask_ref, pack_mmfs_features_ref, prepare_mm_embeds_ref, text_head_ref
    from oracle.llama import llama_model_ref
    from tests.test_generate_gpu import _setup
    cfg, dev, sd, ids, nimg, vis, vis_d = _setup()
    g = torch.Generator().manual_seed(77)
    vis2 = {"vis_embed": torch.randn(vis["vis_embed"].shape, generator=g) * 0.5,
            "multiscale_features": [torch.randn(f.shape, generator=g) * 2 for f in vis["multiscale_features"]]}
    dec = {k[len("mm_decoder."):]: v for k, v in sd.items() if k.startswith("mm_decoder.")}
    ocfg = dict(eps=cfg.rms_norm_eps, n_heads=cfg.num_attention_heads, n_layers=cfg.num_hidden_layers,
                spatial_shapes=[(s, s) for s in cfg.spatial_shapes])

    def oracle_logits(v):
        feats = pack_mmfs_features_ref(v["multiscale_features"], cfg.spatial_shapes, nimg)
        cross = cross_attention_mask_ref(ids, nimg, 1, 63)
        emb = torch.nn.functional.embedding(ids, dec["embed_tokens.weight"])
        emb = prepare_mm_embeds_ref(emb, ids, v["vis_embed"], sd["soi_token"], 62, 63)
        hid, _ = llama_model_ref(dec, emb, torch.ones_like(ids), None, feats, cross, ocfg)
        return text_head_ref(sd, hid, 62)

    outs = []
    for v in (vis, vis2, vis, vis2):                       # round 1 went wrong from the 2nd-4th call on
        vd = {"vis_embed": v["vis_embed"].cuda(), "multiscale_features": [f.cuda() for f in v["multiscale_features"]]}
        with torch.no_grad():
            outs.append(dev(ids.cuda(), vd, nimg.cuda(), 2).float().cpu())
        del vd
    want = [oracle_logits(vis), oracle_logits(vis2)]
    assert float((want[0] - want[1]).abs().max()) > 1e-2
    for i, got in enumerate(outs):
        ref = want[i % 2]
        err = (got - ref).abs()
        assert bool((err <= 1e-3 * ref.abs() + 5e-5).all()), (i, float(err.max()))


@pytest.mark.gpu
def test_mmfsnet_two_calls_with_different_features_match_oracle():
    import mm_interleaved_b200 as m
    from oracle.sd_mmfs import mmfsnet_ref
    from tests.golden.make_golden import MMFSNET_TINY, mmfsnet_inputs, mmfsnet_state_dict
    torch.backends.cuda.matmul.allow_tf32 = False
    net = m.MMFSNet(**MMFSNET_TINY)
    sd = mmfsnet_state_dict(net.state_dict())
    net.load_state_dict(sd, strict=True)
    net = net.cuda().eval()
    sample, res, feats_a, mask = mmfsnet_inputs()
    g = torch.Generator().manual_seed(5)
    feats_b = [torch.randn(f.shape, generator=g) * 1.7 for f in feats_a]

    def run(fd):
        with torch.no_grad():
            s, r = net(sample.cuda(), [x.cuda() for x in res], fd, mask.cuda())
        return [s.cpu()] + [x.cpu() for x in r]

    def check(got, feats):
        s, r = mmfsnet_ref(sd, sample, res, feats, mask, downsample_factor=8, n_down=4)
        for a, b in zip(got, [s] + list(r)):
            err = (a - b).abs()
            assert bool((err <= 1e-3 * b.abs() + 1e-5 * b.abs().max()).all()), float(err.max())

    fa = [f.cuda() for f in feats_a]
    ptrs = [f.data_ptr() for f in fa]
    check(run(fa), feats_a)
    del fa
    fb = [f.cuda() for f in feats_b]
    if [f.data_ptr() for f in fb] != ptrs:
        import warnings
        warnings.warn("allocator did not reuse the freed blocks; address-reuse variant not exercised in this run")
    check(run(fb), feats_b)
    check(run(fb), feats_b)                                   # cached path (same objects) still right
