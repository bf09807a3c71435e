"""GPU tests of the visual tokenizer (CLIP ViT + ViT-Adapter + Q-Former).

* ViT-Adapter blocks (SpatialPriorModule, InteractionBlockWithCls incl. extra extractors, classic MSDeformAttn,
  ConvFFN / DWConv): against the committed outputs of the reference's adapter_modules.py
  (tests/golden/adapter_tiny.npz) -- fp32, |err| <= 1e-3*|ref| + 1e-5*max|ref|.
* CLIP encoder layer and Q-Former: against the transformers 5.x classes of the same name as a stand-in for the
  pinned transformers 4.31 the reference uses ("parity unpinned") -- fp32, 2e-4 abs.
* whole tokenizer: shape / finiteness / bf16-vs-fp32 consistency at a reduced configuration.
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"

from tests.golden.make_golden import ADAPTER_TINY, adapter_inputs, adapter_state_dict  # noqa: E402


@pytest.fixture(autouse=True)
def _no_tf32():
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield
    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old


class _Fake(torch.nn.Module):
    def forward(self, x):
        return torch.tanh(x) * 1.5


def _close(got, want, name):
    want = torch.from_numpy(want)
    err = (got.float().cpu() - want).abs()
    assert (err <= 1e-3 * want.abs() + 1e-5 * want.abs().max()).all(), (name, err.max().item())


def test_adapter_blocks_match_reference_golden():
    from mm_interleaved_b200 import visual_tokenizer as vt
    c = ADAPTER_TINY
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "adapter_tiny.npz"))
    spm = vt.SpatialPriorModule(inplanes=c["inplanes"], embed_dim=c["dim"])
    blk = vt.InteractionBlockWithCls(dim=c["dim"], num_heads=c["heads"], n_points=c["n_points"], init_values=0.0,
                                     with_cffn=True, cffn_ratio=0.25, deform_ratio=0.5, extra_extractor=True)
    sd_spm, sd_blk = adapter_state_dict(spm.state_dict(), 501), adapter_state_dict(blk.state_dict(), 502)
    chk = float(sum(v.double().sum() for v in list(sd_spm.values()) + list(sd_blk.values())))
    assert abs(chk - float(z["checksum"])) < 1e-5
    spm.load_state_dict(sd_spm); blk.load_state_dict(sd_blk)
    spm, blk = spm.to(DEV).eval(), blk.to(DEV).eval()
    img, x, cls = (t.to(DEV) for t in adapter_inputs())
    with torch.no_grad():
        c1, c2, c3, c4 = spm(img)
        d1, d2 = vt.adapter_deform_inputs(img.shape[2], img.shape[3], img.device)
        xo, co, clso = blk(x, torch.cat([c2, c3, c4], 1), cls, [_Fake()], d1, d2, c["H"], c["H"])
    for got, key in ((c1, "c1"), (c2, "c2"), (c3, "c3"), (c4, "c4"), (xo, "x"), (co, "c"), (clso, "cls")):
        _close(got, z[key], key)


def test_clip_encoder_layer_matches_transformers_standin():
    from transformers import CLIPVisionConfig
    from transformers.models.clip.modeling_clip import CLIPEncoderLayer as HFLayer
    from mm_interleaved_b200 import visual_tokenizer as vt
    cfg = vt.CLIPVisionConfigLite(hidden_size=256, intermediate_size=512, num_hidden_layers=1, num_attention_heads=4)
    hf_cfg = CLIPVisionConfig(hidden_size=256, intermediate_size=512, num_hidden_layers=1, num_attention_heads=4,
                              hidden_act="quick_gelu", layer_norm_eps=1e-5)
    hf_cfg._attn_implementation = "eager"
    hf = HFLayer(hf_cfg).to(DEV).eval()
    mine = vt.CLIPEncoderLayer(cfg).to(DEV).eval()
    mine.load_state_dict(hf.state_dict(), strict=True)
    x = torch.randn((2, 17, 256), device=DEV, generator=torch.Generator(device=DEV).manual_seed(0))
    with torch.no_grad():
        want = hf(x, None)
        want = want[0] if isinstance(want, tuple) else want
        got = mine(x)
    assert (got - want).abs().max() < 2e-4


def test_qformer_matches_transformers_standin():
    from transformers import Blip2QFormerConfig, Blip2QFormerModel
    from mm_interleaved_b200 import visual_tokenizer as vt
    kw = dict(hidden_size=192, encoder_hidden_size=256, num_hidden_layers=4, num_attention_heads=3, cross_attention_frequency=2,
              intermediate_size=384)
    hf_cfg = Blip2QFormerConfig(**kw, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    hf_cfg._attn_implementation = "eager"
    hf = Blip2QFormerModel(hf_cfg).to(DEV).eval()
    mine = vt.PerceiverResampler(num_queries=8, qk_normalization=False, **kw).to(DEV).eval()
    mine.blip2qformer.load_state_dict(hf.state_dict(), strict=True)
    g = torch.Generator(device=DEV).manual_seed(1)
    q = torch.randn((2, 8, 192), device=DEV, generator=g)
    enc = torch.randn((2, 17, 256), device=DEV, generator=g)
    with torch.no_grad():
        want = hf(query_embeds=q, encoder_hidden_states=enc, return_dict=True).last_hidden_state
        got = mine(encoder_hidden_states=enc, query_embeds=q)[0]
    assert (got - want).abs().max() < 2e-4


def test_tokenizer_end_to_end_reduced_config():
    from mm_interleaved_b200 import visual_tokenizer as vt
    torch.manual_seed(0)
    clip = vt.CLIPVisionConfigLite(hidden_size=256, intermediate_size=512, num_hidden_layers=4, num_attention_heads=4,
                                   image_size=56, patch_size=14)
    tok = vt.VisualTokenizer(clip_config=clip, perceiver_config=dict(num_queries=8, hidden_size=192, encoder_hidden_size=256,
                             cross_attention_frequency=2, num_hidden_layers=2, num_attention_heads=3,
                             intermediate_size=384, qk_normalization=True), llm_hidden_size=320, grid_size=4)
    with torch.no_grad():
        for m in tok.modules():                        # make the zero-initialised branches observable
            if isinstance(m, vt.Injector):
                m.gamma.fill_(0.5)
    tok = tok.to(DEV).eval()
    img = torch.rand((3, 3, 56, 56), device=DEV)
    with torch.no_grad():
        out = tok(img)
        out16 = tok.to(torch.bfloat16)(img.to(torch.bfloat16))
    assert out["vis_embed"].shape == (3, 8, 320) and out["image_embeds"].shape == (3, 16, 256)
    assert [tuple(f.shape[1:]) for f in out["multiscale_features"]] == [(256, 16, 16), (256, 8, 8), (256, 4, 4), (256, 2, 2)]
    for k in ("vis_embed", "image_embeds"):
        assert torch.isfinite(out[k]).all()
        ref = out[k].float()
        assert (out16[k].float() - ref).abs().max() <= 0.1 * ref.abs().max() + 1e-3
