"""CPU tests of the batch-preparation glue (vectorised, sync-free) against the loop restatement of the
reference helpers (mm_interleaved.py:121-252).  Integer / index work: bit-exact."""
import torch

from oracle.glue import cross_attention_mask_ref, pack_mmfs_features_ref, prepare_mm_embeds_ref


def synthetic_batch(seed=0):
    g = torch.Generator().manual_seed(seed)
    BOS, IMG, SOI = 1, 32000, 32001
    L = 96
    nimg = torch.tensor([3, 0, 1, 2])
    ids = torch.randint(3, 31999, (4, L), generator=g)
    ids[:, 0] = BOS
    layout = {0: [2, 30, 60], 2: [50], 3: [1, 40]}
    for b, starts in layout.items():
        for s in starts:
            ids[b, s] = SOI
            ids[b, s + 1:s + 5] = IMG              # 4 image tokens per image in this miniature
    ids[0, 45] = BOS                                # packed-document boundary: earlier images become invisible
    ids[3, 39] = BOS
    return ids, nimg, BOS, IMG, SOI, g


def test_cross_attention_mask_matches_reference_loops():
    from mm_interleaved_b200.mm_interleaved import cross_attention_mask_from_ids
    ids, nimg, BOS, IMG, SOI, _ = synthetic_batch()
    want = cross_attention_mask_ref(ids, nimg, BOS, SOI)
    got = cross_attention_mask_from_ids(ids, int(nimg.max()), BOS, SOI, nimg)
    assert torch.equal(got, want)
    assert want[0, 44, 0] == 1 and want[0, 46, 0] == 0 and want[0, 61:, 2].all()      # the gating really bites
    # a larger static image budget only appends never-visible slots
    got5 = cross_attention_mask_from_ids(ids, 5, BOS, SOI, nimg)
    assert torch.equal(got5[..., :3], want) and got5[..., 3:].sum() == 0


def test_embed_splice_matches_reference_scatter():
    from mm_interleaved_b200.mm_interleaved import splice_image_embeds
    ids, nimg, BOS, IMG, SOI, g = synthetic_batch(1)
    C = 16
    emb = torch.randn((4, ids.shape[1], C), generator=g)
    img = torch.randn((int(nimg.sum()), 4, C), generator=g)
    soi = torch.randn((1, C), generator=g)
    want = prepare_mm_embeds_ref(emb, ids, img, soi, IMG, SOI)
    got = splice_image_embeds(emb, ids, img, soi, IMG, SOI)
    assert torch.equal(got, want)


def test_feature_packing_matches_reference_loops():
    from mm_interleaved_b200.mm_interleaved import pack_mmfs_features
    g = torch.Generator().manual_seed(2)
    nimg = torch.tensor([3, 0, 1, 2])
    feats = [torch.randn((6, 8, s, s), generator=g) for s in (16, 8, 4, 2)]
    want = pack_mmfs_features_ref(feats, [8, 4, 2], nimg)
    got = pack_mmfs_features(feats, [8, 4, 2], nimg, 3)
    assert torch.equal(got, want)
