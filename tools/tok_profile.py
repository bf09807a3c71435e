"""Kernel-time breakdown of the visual tokenizer (16 images, bf16) with torch.profiler."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mm_interleaved_b200.visual_tokenizer import Injector, VisualTokenizer  # noqa: E402

torch.manual_seed(0)
tok = VisualTokenizer()
for m in tok.modules():
    if isinstance(m, Injector):
        m.gamma.data.fill_(0.5)
tok = tok.to("cuda", torch.bfloat16).eval()
x = torch.rand((16, 3, 224, 224), device="cuda", dtype=torch.bfloat16)
with torch.no_grad():
    for _ in range(3):
        tok(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); tok(x); e1.record(); torch.cuda.synchronize()
    print("eager ms", e0.elapsed_time(e1))
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        tok(x)
        torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=70))
