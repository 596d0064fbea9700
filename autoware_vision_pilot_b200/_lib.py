"""ctypes binding of libvp_b200.so (the C-ABI declared in include/vp_b200*.h).

The library is the product; this module only loads it.  There is deliberately no
fallback: if the shared object is missing the import of any compute entry point
raises, so a GPU box can never silently run something else.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libvp_b200.so")

VPB_F16, VPB_BF16 = 0, 1
ACT_NONE, ACT_GELU, ACT_SILU, ACT_SIGMOID = 0, 1, 2, 3
EPI_STORE, EPI_ADD, EPI_MULADD, EPI_FINAL = 0, 1, 2, 3
FINAL_NONE, FINAL_ARGMAX, FINAL_THRESH, FINAL_EGOLANES = 0, 1, 2, 3
ALGO_TILE, ALGO_LINEAR = 0, 1


class ConvArgs(C.Structure):
    """Mirror of vpb_conv_args (include/vp_b200_ops.h)."""

    _fields_ = [
        ("dtype", C.c_int),
        ("H", C.c_int), ("W", C.c_int), ("Cin", C.c_int), ("ldi", C.c_int),
        ("Cout", C.c_int), ("taps", C.c_int), ("phases", C.c_int),
        ("act", C.c_int), ("mode", C.c_int), ("final_kind", C.c_int),
        ("inp", C.c_void_p), ("w", C.c_void_p), ("bias", C.c_void_p),
        ("out", C.c_void_p), ("ldo", C.c_int),
        ("res", C.c_void_p), ("ldr", C.c_int),
        ("out_f32", C.c_void_p), ("out_cls", C.c_void_p),
        ("bn", C.c_int),
        ("in_pad", C.c_int), ("out_pad", C.c_int), ("res_pad", C.c_int),
        ("algo", C.c_int), ("dbg_ms", C.c_int), ("dbg_gb", C.c_int), ("dbg_base_offset", C.c_int),
        ("in2", C.c_void_p), ("w2", C.c_void_p),
        ("Cin2", C.c_int), ("ld2", C.c_int), ("in2_pad", C.c_int), ("dbg_pair", C.c_int),
    ]


_lib = None


def lib() -> C.CDLL:
    """Load (once) and return the shared library; raise loudly if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing — run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU or PyTorch fallback for the B200 path)")
        _lib = C.CDLL(LIB_PATH)
        _lib.vpb_last_error.restype = C.c_char_p
        _lib.vpb_conv_gemm.argtypes = [C.POINTER(ConvArgs), C.c_void_p]
        _lib.vpb_conv_gemm.restype = C.c_int
    return _lib


def last_error() -> str:
    return lib().vpb_last_error().decode("utf-8", "replace")


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"{what} failed (rc={rc}): {last_error()}")
