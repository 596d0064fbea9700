"""Timeline of CTA 0 of upconv_pair_kernel (clock64 stamps through the dbg_trace hook):
python scripts/trace_upconv.py H W Cin Cout C2 [bn [gb]]"""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
from autoware_vision_pilot_b200 import _lib as L  # noqa: E402


def main():
    H, W, Cin, Cout, C2 = map(int, sys.argv[1:6])
    bn = int(sys.argv[6]) if len(sys.argv) > 6 else 0
    gb = int(sys.argv[7]) if len(sys.argv) > 7 else 0     # 3 = staged TMA-store epilogue
    lib = L.lib()
    x = torch.randn(H, W, Cin, device="cuda").half()
    w = (torch.randn(16, Cout, Cin, device="cuda") * 0.02).half()
    b = torch.randn(9, Cout, device="cuda")
    o = torch.zeros(2 * H, 2 * W, Cout, device="cuda", dtype=torch.half)
    tr = torch.zeros(256, device="cuda", dtype=torch.int64)
    a = L.ConvArgs()
    a.dtype = L.VPB_F16
    a.H, a.W, a.Cin, a.ldi, a.Cout, a.taps, a.phases = H, W, Cin, Cin, Cout, 4, 4
    a.act, a.bn, a.dbg_gb = L.ACT_GELU, bn, gb
    a.inp, a.w, a.bias, a.out, a.ldo = x.data_ptr(), w.data_ptr(), b.data_ptr(), o.data_ptr(), Cout
    if C2:
        s = torch.randn(2 * H, 2 * W, C2, device="cuda").half()
        w2 = (torch.randn(9, Cout, C2, device="cuda") * 0.02).half()
        a.in2, a.w2, a.Cin2, a.ld2, a.taps2 = s.data_ptr(), w2.data_ptr(), C2, C2, 9
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for i in range(3):
        a.dbg_trace = tr.data_ptr() if i == 2 else None
        ev[0].record()
        L.check(lib.vpb_conv_gemm(C.byref(a), None), "conv")
        ev[1].record()
        torch.cuda.synchronize()
        if i == 1:
            print("kernel time %.1f us" % (ev[0].elapsed_time(ev[1]) * 1e3))
    t = tr.cpu().numpy()
    t0 = t[255]
    f = lambda v: "%6d" % (v - t0) if v else "     -"
    print("all stores complete at", f(t[253]), "(cycles since kernel body start)")
    print("tile | load issue: first, last main tap, last skip tap | MMA: acc free, first issue, last issue | epi: acc full, slab free, tmem->smem done")
    for i in range(8):
        r = t[i * 16:(i + 1) * 16]
        if r[0] == 0 and r[10] == 0:
            break
        print(i, "|", f(r[0]), f(r[1]), f(r[2]), "|", f(r[9]), f(r[4]), f(r[8]), "|", f(r[10]), f(r[11]), f(r[12]))


if __name__ == "__main__":
    main()
