"""Greedy decode timing of the Llama-13B MMFS decoder (random weights, bf16): prefill on 4 x 2048-token 4-image
sequences, then N new tokens with (a) the pre-allocated in-place KV cache and (b) the reference-style cat-per-token cache."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from benchmarks import workloads  # noqa: E402

wl = workloads.make("interleaved_cfg3", rank=0, world=1, local_batch=int(os.environ.get("LOCAL_BATCH", 4)))
wl.setup()
ids, img = wl.dev[0], wl.dev[1]
with torch.no_grad():
    wl.tok_in.copy_(img)
    wl.tok_graph.replay()
    vis = wl.tok_out
    n_new = int(os.environ.get("N_NEW", 16))
    for static in (True, False):
        for rep in range(2):
            torch.cuda.synchronize(); t0 = time.time()
            out = wl.model.generate_texts(ids, vis, wl.nimg, wl.N_IMG, max_new_tokens=1, eos_token_id=None, static_cache=static)
            torch.cuda.synchronize(); t1 = time.time()
            out = wl.model.generate_texts(ids, vis, wl.nimg, wl.N_IMG, max_new_tokens=1 + n_new, eos_token_id=None, static_cache=static)
            torch.cuda.synchronize(); t2 = time.time()
        per_tok = ((t2 - t1) - (t1 - t0)) / n_new
        print(f"static_cache={static}: prefill+1 {1e3 * (t1 - t0):.1f} ms, {1e3 * per_tok:.2f} ms per decoded token (batch {ids.shape[0]}), "
              f"{ids.shape[0] / per_tok:.1f} tokens/s")
