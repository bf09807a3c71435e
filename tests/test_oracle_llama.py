"""CPU tests: pin oracle/llama.py against the committed outputs of the reference LlamaModel (tiny config,
seeded weights, left-padded batch, 2 images, prefill + one cached decode step)."""
import os

import numpy as np
import torch

from oracle import error_metrics
from oracle.llama import llama_model_ref
from tests.golden.make_golden import LLAMA_TINY, llama_inputs, seeded_state_dict

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "llama_tiny.npz")


def tiny_state_dict():
    import mm_interleaved_b200  # noqa: F401  (module structure only; no CUDA needed to build it)
    from mm_interleaved_b200.llama_mmfs import LlamaMMFSConfig, LlamaModel
    model = LlamaModel(LlamaMMFSConfig(**LLAMA_TINY))
    sd = seeded_state_dict(model.state_dict(), seed=4242)
    z = np.load(GOLDEN)
    assert abs(float(sum(v.double().sum() for v in sd.values())) - float(z["weight_checksum"])) < 1e-6, \
        "seeded weights do not reproduce (parameter names or RNG drifted)"
    return model, sd, z


def oracle_cfg():
    return dict(eps=LLAMA_TINY["rms_norm_eps"], n_heads=LLAMA_TINY["num_attention_heads"],
                n_layers=LLAMA_TINY["num_hidden_layers"], spatial_shapes=[(s, s) for s in LLAMA_TINY["spatial_shapes"]])


def test_state_dict_names_match_reference_and_oracle_matches_golden():
    _, sd, z = tiny_state_dict()
    embeds, vision, attn_mask, position_ids, cross = llama_inputs(LLAMA_TINY, 2, 12, 2, seed=99)
    out, kvs = llama_model_ref(sd, embeds, attn_mask, position_ids, vision, cross, oracle_cfg())
    valid = attn_mask.bool()
    m = error_metrics(out[valid], torch.from_numpy(z["prefill_fp32"])[valid])
    assert m["max_abs"] < 5e-5, m
    # cached decode step
    g = torch.Generator().manual_seed(7)
    step = torch.randn((2, 1, LLAMA_TINY["hidden_size"]), generator=g)
    attn2 = torch.cat([attn_mask, torch.ones((2, 1), dtype=torch.long)], 1)
    out2, _ = llama_model_ref(sd, step, attn2, position_ids[:, -1:] + 1, vision, torch.cat([cross, cross[:, -1:]], 1),
                              oracle_cfg(), past=kvs)
    assert error_metrics(out2, torch.from_numpy(z["decode_fp32"]))["max_abs"] < 5e-5
