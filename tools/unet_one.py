"""One SD-2.1 UNet evaluation with the MMFS hook at the sd_cfg4 batch (16 rows = 8 images x CFG 2, bf16, channels-last):
the ncu target for conv_igemm_kernel / groupnorm_nhwc / attn_fwd_kernel (tensor-pipe utilisation) captures."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mm_interleaved_b200 as m  # noqa: E402
from mm_interleaved_b200 import unet_sd  # noqa: E402

torch.manual_seed(0)
B = int(os.environ.get("UNET_B", 16))
dt = torch.bfloat16
unet = unet_sd.UNet2DConditionModel().to("cuda", dt).eval().to(memory_format=torch.channels_last)
net = m.MMFSNet(1024, (320, 640, 1280, 1280), 2).to("cuda", dt).eval()
x = torch.randn((B, 4, 64, 64), device="cuda", dtype=dt).contiguous(memory_format=torch.channels_last)
ctx = torch.randn((B, 77, 1024), device="cuda", dtype=dt) * 0.1
feats = [torch.randn((B, 1, 1024, s, s), device="cuda", dtype=dt) for s in (64, 32, 16, 8)]
mask = torch.ones((B, 1), device="cuda")
t = torch.tensor(500, device="cuda")
with torch.no_grad():
    for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
        out = unet(x, t, ctx, mmfs_features=feats, mmfs_mask=mask, mmfs_module=net)
torch.cuda.synchronize()
print("ok", float(out.float().abs().mean()))
