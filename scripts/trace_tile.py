"""Timeline of CTA 0 of the tile convolution kernel (clock64 stamps through the dbg_trace hook).
python scripts/trace_tile.py H W Cin Cout taps phases [bn]"""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
from autoware_vision_pilot_b200 import _lib as L  # noqa: E402


def main():
    H, W, Cin, Cout, taps, phases = map(int, sys.argv[1:7])
    bn = int(sys.argv[7]) if len(sys.argv) > 7 else 0
    lib = L.lib()
    x = torch.randn(H, W, Cin, device="cuda").half()
    w = (torch.randn(taps * phases, Cout, Cin, device="cuda") * 0.02).half()
    b = torch.randn(Cout, device="cuda")
    Ho, Wo = (2 * H, 2 * W) if phases == 4 else (H, W)
    o = torch.zeros(Ho, Wo, Cout, device="cuda", dtype=torch.half)
    tr = torch.zeros(256, device="cuda", dtype=torch.int64)
    a = L.ConvArgs()
    a.dtype = L.VPB_F16
    a.H, a.W, a.Cin, a.ldi, a.Cout, a.taps, a.phases = H, W, Cin, Cin, Cout, taps, phases
    a.inp, a.w, a.bias, a.out, a.ldo, a.bn = x.data_ptr(), w.data_ptr(), b.data_ptr(), o.data_ptr(), Cout, bn
    for i in range(3):
        a.dbg_trace = tr.data_ptr() if i == 2 else None
        L.check(lib.vpb_conv_gemm(C.byref(a), None), "conv")
        torch.cuda.synchronize()
    t = tr.cpu().numpy()
    t0 = t[255]
    print("tile | load issue k0..k3 | full seen k0..k3 | last mma issued | acc free | epi: bias staged, acc full, done   (cycles since kernel start)")
    for i in range(8):
        r = t[i * 16:(i + 1) * 16]
        if r[0] == 0:
            break
        f = lambda v: "%6d" % (v - t0) if v else "     -"
        print(i, "|", " ".join(f(v) for v in r[0:4]), "|", " ".join(f(v) for v in r[4:8]), "|", f(r[8]), "|", f(r[9]), "|", f(r[10]), f(r[11]), f(r[12]))


if __name__ == "__main__":
    main()
