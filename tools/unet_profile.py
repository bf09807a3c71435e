"""Kernel-time breakdown of one SD-2.1 UNet evaluation at the sd_cfg4 batch (16 = 8 images x CFG 2, bf16, channels-last),
plus per-layer timing of the implicit-GEMM convolution against cuDNN on the same shapes."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mm_interleaved_b200 import ops, unet_sd  # noqa: E402

torch.manual_seed(0)
B = int(os.environ.get("UNET_B", 16))
unet = unet_sd.UNet2DConditionModel().to("cuda", torch.bfloat16).eval().to(memory_format=torch.channels_last)
x = torch.randn((B, 4, 64, 64), device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
ctx = torch.randn((B, 77, 1024), device="cuda", dtype=torch.bfloat16) * 0.1
t = torch.tensor(500, device="cuda")


def timed(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


with torch.no_grad():
    print("unet eval ms (ours)", timed(lambda: unet(x, t, ctx)))
    unet_sd.USE_CONV_KERNEL = False
    print("unet eval ms (cuDNN convs, torch group_norm)", timed(lambda: unet(x, t, ctx)))
    unet_sd.USE_CONV_KERNEL = True
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        unet(x, t, ctx)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=80))
    rows = sorted(prof.key_averages(), key=lambda e: -e.count)
    print(f"kernels per evaluation: {sum(e.count for e in rows)}; by launch count:")
    for e in rows[:24]:
        print(f"  n={e.count:4d}  total {e.device_time_total / 1e3:7.2f} ms  avg {e.device_time_total / max(e.count, 1):7.1f} us  {e.key[:90]}")
    print("layer: B Cin Cout H k s | ours us  TF/s | cudnn us")
    for (Cin, Cout, H, k, s) in [(320, 320, 64, 3, 1), (320, 320, 64, 3, 2), (640, 640, 32, 3, 1), (320, 640, 32, 3, 1),
                                 (1280, 1280, 16, 3, 1), (1280, 1280, 8, 3, 1), (2560, 1280, 8, 3, 1), (2560, 1280, 16, 3, 1),
                                 (1920, 640, 32, 3, 1), (960, 320, 64, 3, 1), (640, 320, 64, 3, 1), (640, 320, 64, 1, 1)]:
        conv = torch.nn.Conv2d(Cin, Cout, k, stride=s, padding=k // 2).to("cuda", torch.bfloat16).to(memory_format=torch.channels_last)
        xi = torch.randn((B, Cin, H, H), device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
        w = conv.weight.detach().permute(0, 2, 3, 1).contiguous()
        ours = timed(lambda: ops.conv2d(xi, w, conv.bias, s, k // 2), 10) * 1e3
        lib = timed(lambda: conv(xi), 10) * 1e3
        Ho = H // s
        fl = 2.0 * B * Ho * Ho * Cout * Cin * k * k
        print(f"{B} {Cin:5d} {Cout:5d} {H:3d} {k} {s} | {ours:8.1f} {fl / ours / 1e6:7.1f} | {lib:8.1f} {fl / lib / 1e6:7.1f}")
