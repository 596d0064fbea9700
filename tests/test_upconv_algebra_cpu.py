"""CPU check of the algebra behind the fused ConvTranspose2d(k2,s2) [+ Conv1x1 skip] -> Conv3x3 GEMM:
the fp64 composition used as the yardstick by tests/test_upconv_gpu.py, evaluated in the kernel's form (four output
phases = 2x2 convolutions of the LOW-resolution tensor, 3x3 convolution of the skip tensor, nine border-class bias
rows), must reproduce torch's conv_transpose2d + conv2d of the reference's layer definitions
(Models/model_components/scene_neck.py:30-37) to fp64 round-off."""
import pytest
import torch
import torch.nn.functional as F

from tests.test_upconv_gpu import _compose64, _params


@pytest.mark.parametrize("H,W,Cin,Cmid,Cout,C2", [(5, 6, 6, 5, 7, 3), (1, 1, 4, 4, 4, 0), (2, 7, 8, 3, 5, 2), (4, 1, 3, 6, 2, 0)])
def test_composed_form_equals_two_layers(H, W, Cin, Cmid, Cout, C2):
    dd = torch.float64
    wt, bt, w3, b3, ws, bs = _params(Cin, Cmid, Cout, C2, seed=H * 10 + W)
    wf, w2f, b9 = _compose64(wt, bt, w3, b3, ws, bs)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(1, Cin, H, W, generator=g, dtype=dd)
    s = torch.randn(1, C2, 2 * H, 2 * W, generator=g, dtype=dd) if C2 else None
    up = F.conv_transpose2d(x, wt.to(dd), bt.to(dd), stride=2)
    if C2:
        up = up + F.conv2d(s, ws.to(dd).reshape(Cmid, C2, 1, 1), bs.to(dd))
    ref = F.conv2d(up, w3.to(dd), b3.to(dd), padding=1)[0]

    xp = F.pad(x, (1, 1, 1, 1))
    out = torch.zeros(Cout, 2 * H, 2 * W, dtype=dd)
    for a in range(2):
        for b in range(2):
            k = wf[(a * 2 + b) * 4:(a * 2 + b) * 4 + 4].reshape(2, 2, Cout, Cin).permute(2, 3, 0, 1).contiguous()
            y = F.conv2d(xp, k)[0]            # y[i, j] = sum_{ty,tx} k[ty,tx] . x[i+ty-1, j+tx-1]
            out[:, a::2, b::2] = y[:, a:a + H, b:b + W]
    if C2:
        out += F.conv2d(s, w2f.reshape(3, 3, Cout, C2).permute(2, 3, 0, 1).contiguous(), padding=1)[0]
    cy = torch.ones(2 * H, dtype=torch.long); cy[0] = 0; cy[-1] = 2
    cx = torch.ones(2 * W, dtype=torch.long); cx[0] = 0; cx[-1] = 2
    out += b9[cy[:, None] * 3 + cx[None, :]].permute(2, 0, 1)
    assert (out - ref).abs().max().item() < 1e-11
