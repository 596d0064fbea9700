"""Timeline of CTA 0 of convt_ws_kernel (clock64 stamps through the dbg_trace hook): python scripts/trace_ws.py H W Cin Cout"""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
from autoware_vision_pilot_b200 import _lib as L  # noqa: E402


def main():
    H, W, Cin, Cout = map(int, sys.argv[1:5])
    lib = L.lib()
    x = torch.randn(H, W, Cin, device="cuda").half()
    w = (torch.randn(4, Cout, Cin, device="cuda") * 0.02).half()
    b = torch.randn(Cout, device="cuda")
    o = torch.zeros(2 * H, 2 * W, Cout, device="cuda", dtype=torch.half)
    tr = torch.zeros(256, device="cuda", dtype=torch.int64)
    a = L.ConvArgs()
    a.dtype = L.VPB_F16
    a.H, a.W, a.Cin, a.ldi, a.Cout, a.taps, a.phases = H, W, Cin, Cin, Cout, 1, 4
    a.inp, a.w, a.bias, a.out, a.ldo = x.data_ptr(), w.data_ptr(), b.data_ptr(), o.data_ptr(), Cout
    for i in range(3):
        a.dbg_trace = tr.data_ptr() if i == 2 else None
        L.check(lib.vpb_conv_gemm(C.byref(a), None), "conv")
        torch.cuda.synchronize()
    t = tr.cpu().numpy()
    t0 = t[255]
    f = lambda v: "%6d" % (v - t0) if v else "     -"
    print("weights resident at", f(t[254]), " all stores complete at", f(t[253]), "(cycles since kernel body start)")
    print("tile | A load issue first,last | first A seen by MMA | last MMA issued | acc free | epi: acc full, slabs free, tmem->smem done, store issued")
    for i in range(8):
        r = t[i * 16:(i + 1) * 16]
        if r[0] == 0 and r[10] == 0:
            break
        print(i, "|", f(r[0]), f(r[1]), "|", f(r[4]), "|", f(r[8]), "|", f(r[9]), "|", f(r[10]), f(r[11]), f(r[12]), f(r[13]))


if __name__ == "__main__":
    main()
