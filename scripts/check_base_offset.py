"""Which smem-descriptor base_offset convention do row-shifted views need?  (GPU box)"""
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
from autoware_vision_pilot_b200 import _lib as L  # noqa: E402
from tests.gpu_util import conv_gemm, pad_img  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
g = torch.Generator().manual_seed(0)
H, W, Cin, Cout = 16, 32, 64, 64
x = torch.randn(H, W, Cin, generator=g).half().cuda()
w = (torch.randn(9, Cout, Cin, generator=g) / 24).half().cuda()
b = torch.randn(Cout, generator=g).cuda()
wf = w.float().view(3, 3, Cout, Cin).permute(2, 3, 0, 1).contiguous()
ref = F.conv2d(x.float().permute(2, 0, 1).unsqueeze(0), wf, b, padding=1)[0].permute(1, 2, 0)
for set_bo in (0, 1):
    _, _, out = conv_gemm(pad_img(x), w, b, taps=9, in_pad=1, out_pad=1, algo=L.ALGO_LINEAR, set_bo=set_bo)
    err = (out[1:-1, 1:-1].float() - ref).abs().max().item()
    print(f"base_offset {'= dx' if set_bo else '= 0'}: max abs err {err:.4g}  ->", "OK" if err < 5e-3 else "WRONG")
