"""Run one conv layer shape a few times (for ncu).  python scripts/run_layer.py name H W Cin Cout taps phases mode"""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
from autoware_vision_pilot_b200 import _lib as L  # noqa: E402


def main():
    name = sys.argv[1]
    H, W, Cin, Cout, taps, phases = map(int, sys.argv[2:8])
    mode = int(sys.argv[8]) if len(sys.argv) > 8 else 0
    iters = int(sys.argv[9]) if len(sys.argv) > 9 else 3
    lin = taps == 9 and "tile" not in sys.argv
    lib = L.lib()
    if lin:
        x = torch.zeros(H + 2, W + 2, Cin, device="cuda").half()
        x[1:-1, 1:-1] = torch.randn(H, W, Cin, device="cuda").half()
    else:
        x = torch.randn(H, W, Cin, device="cuda").half()
    w = (torch.randn(taps * phases, Cout, Cin, device="cuda") * 0.02).half()
    b = torch.randn(Cout, device="cuda")
    a = L.ConvArgs()
    a.dtype = L.VPB_F16
    a.H, a.W, a.Cin, a.ldi, a.Cout, a.taps, a.phases = H, W, Cin, Cin, Cout, taps, phases
    a.act = L.ACT_GELU if taps == 9 else L.ACT_NONE
    a.inp, a.w, a.bias = x.data_ptr(), w.data_ptr(), b.data_ptr()
    if lin:
        a.in_pad, a.algo = 1, L.ALGO_LINEAR
    Ho, Wo = (2 * H, 2 * W) if phases == 4 else (H, W)
    if Cout <= 16:
        of = torch.empty(Cout, H, W, device="cuda")
        oc = torch.empty(H, W, device="cuda", dtype=torch.uint8)
        a.mode, a.final_kind, a.out_f32, a.out_cls = L.EPI_FINAL, L.FINAL_ARGMAX, of.data_ptr(), oc.data_ptr()
    else:
        ldo = (Cout + 7) // 8 * 8
        pad = 1 if lin else 0
        o = torch.zeros(Ho + 2 * pad, Wo + 2 * pad, ldo, device="cuda", dtype=torch.half)
        a.mode, a.out, a.ldo, a.out_pad = mode, o.data_ptr(), ldo, pad
        if mode == L.EPI_ADD:
            a.res, a.ldr = o.data_ptr(), ldo
    for _ in range(iters):
        L.check(lib.vpb_conv_gemm(C.byref(a), None), name)
    torch.cuda.synchronize()
    print("ran", name)


if __name__ == "__main__":
    main()
