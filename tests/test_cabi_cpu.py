"""CPU-only checks of the shared library: it loads, exports every symbol the headers declare, and
its host-only pieces (integer resize tables, argument validation, .vpw writer) agree with the
oracle.  No kernel is launched here."""
import ctypes as C
import os
import re
import struct

import numpy as np
import pytest

from autoware_vision_pilot_b200 import _lib as L
from autoware_vision_pilot_b200 import weights as W
from oracle import resize

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    names = set()
    for h in ("vp_b200.h", "vp_b200_ops.h", "vp_b200_multicam.h", "vp_b200_autospeed.h"):
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b(vpb?_[a-z0-9_]+)\s*\(", src))
    return sorted(names)


def test_library_exports_every_declared_symbol():
    lib = L.lib()
    syms = _declared_symbols()
    assert len(syms) >= 20
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing


@pytest.mark.parametrize("mode,in_size,out_size", [(1, 1920, 640), (1, 1080, 320), (1, 700, 320), (1, 401, 640),
                                                   (2, 1920, 640), (2, 1080, 320), (2, 660, 320), (2, 517, 640)])
def test_resize_tables_match_oracle(mode, in_size, out_size):
    """The C++ host code that builds the kernel's integer coefficient tables reproduces the
    Pillow / OpenCV restatements exactly (which are themselves pinned against the libraries)."""
    lib = L.lib()
    lib.vpb_resize_tables_host.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                           C.c_int, C.POINTER(C.c_int)]
    bounds = (C.c_int * out_size)()
    cap = out_size * 64
    coeffs = (C.c_int * cap)()
    ks = C.c_int()
    L.check(lib.vpb_resize_tables_host(mode, in_size, out_size, bounds, coeffs, cap, C.byref(ks)), "tables")
    k = ks.value
    got = np.frombuffer(coeffs, dtype=np.int32)[: out_size * k].reshape(out_size, k)
    if mode == 1:
        b, cs = resize.pil_coeffs(in_size, out_size)
        for o in range(out_size):
            assert bounds[o] == b[o]
            n = len(cs[o])
            assert np.array_equal(got[o, :n], cs[o])
            assert not got[o, n:].any()
    else:
        idx, w0, w1 = resize.cv_linear_coeffs(in_size, out_size)
        assert np.array_equal(np.frombuffer(bounds, dtype=np.int32), idx)
        assert np.array_equal(got[:, 0], w0) and np.array_equal(got[:, 1], w1)


def test_conv_rejects_bad_arguments_without_a_gpu():
    lib = L.lib()
    a = L.ConvArgs()
    a.H, a.W, a.Cin, a.ldi, a.Cout, a.taps, a.phases = 8, 8, 12, 12, 8, 9, 1   # Cin not multiple of 8
    assert lib.vpb_conv_gemm(C.byref(a), None) == -1
    assert "multiples of 8" in L.last_error()
    a.Cin = a.ldi = 16
    a.taps = 4
    assert lib.vpb_conv_gemm(C.byref(a), None) == -1
    # upconv (taps = 4, phases = 4): the [9][Cout] bias, Cout % 16 and the skip tap count are checked before any device work
    a.phases, a.Cout = 4, 32
    assert lib.vpb_conv_gemm(C.byref(a), None) == -1 and "upconv" in L.last_error()      # no bias
    dummy = (C.c_float * (9 * 32))()
    a.bias = C.addressof(dummy)
    a.Cout = 24
    assert lib.vpb_conv_gemm(C.byref(a), None) == -1 and "upconv" in L.last_error()      # Cout not a multiple of 16
    a.Cout, a.in2, a.w2, a.Cin2, a.ld2, a.taps2 = 32, C.addressof(dummy), C.addressof(dummy), 8, 8, 1
    assert lib.vpb_conv_gemm(C.byref(a), None) == -1 and "upconv" in L.last_error()      # skip taps must be 9
    lib.vpb_upconv_compose.argtypes = [C.c_void_p] * 6 + [C.c_int] * 4 + [C.c_void_p] * 4
    assert lib.vpb_upconv_compose(None, None, None, None, None, None, 8, 8, 8, 0, None, None, None, None) == -1


def test_vpw_writer_layout(tmp_path):
    sd = {"a.weight": np.arange(24, dtype=np.float32).reshape(2, 3, 2, 2),
          "a.num_batches_tracked": np.array(7, dtype=np.int64)}
    p = W.write_vpw(sd, str(tmp_path / "t.vpw"))
    raw = open(p, "rb").read()
    assert raw[:4] == b"VPW1" and struct.unpack("<I", raw[4:8])[0] == 2
    nl = struct.unpack("<I", raw[8:12])[0]
    assert raw[12:12 + nl] == b"a.weight"
    off = 12 + nl
    dt, nd = struct.unpack("<II", raw[off:off + 8])
    assert (dt, nd) == (0, 4)
    dims = struct.unpack("<4I", raw[off + 8:off + 24])
    assert dims == (2, 3, 2, 2)
    nbytes = struct.unpack("<Q", raw[off + 24:off + 32])[0]
    assert nbytes == 96
    assert np.array_equal(np.frombuffer(raw[off + 32:off + 32 + 96], dtype=np.float32), np.arange(24))


def test_infer_helpers_keep_reference_error_behaviour():
    from autoware_vision_pilot_b200.inference import (DomainSegNetworkInfer, EgoLanesNetworkInfer,
                                                      Scene3DNetworkInfer, SceneSegNetworkInfer)
    for K in (SceneSegNetworkInfer, Scene3DNetworkInfer, DomainSegNetworkInfer):
        with pytest.raises(ValueError):
            K(checkpoint_path="")          # scene_seg_infer.py:32-33
    # EgoLanes accepts an empty path ("vanilla" randomly initialised model, ego_lanes_infer.py:34-44): no ValueError;
    # without a GPU the engine itself then refuses (no CPU fallback)
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            EgoLanesNetworkInfer(checkpoint_path="")


def test_vanilla_ego_lanes_state_dict_has_the_reference_layout():
    """Names / shapes of the randomly initialised EgoLanes checkpoint == the oracle's state_dict spec (which is
    strict-loaded into the unmodified reference module in tests/test_oracle_vs_reference.py)."""
    from oracle import synth
    a = [(n, tuple(s)) for n, s in W.ego_lanes_spec()]
    b = [(n, tuple(s)) for n, s, _ in synth.state_dict_spec("ego_lanes")]
    assert a == b
    sd = W.vanilla_ego_lanes_state_dict()
    assert sd["BEVBackbone.encoder.0.1.running_var"].min() == 1.0 and not sd["BEVBackbone.encoder.0.1.bias"].any()
    w = sd["EgopathNeck.decode_layer_0.weight"]
    assert abs(w).max() <= 1.0 / np.sqrt(1456 * 9) and w.std() > 0


def test_write_vpw_is_atomic_under_concurrent_writers(tmp_path):
    """Several ranks converting the same checkpoint at start-up must never publish a partial file."""
    import multiprocessing as mp
    sd = {"a.weight": np.arange(1 << 16, dtype=np.float32)}
    path = str(tmp_path / "shared.vpw")
    ctx = mp.get_context("fork")
    procs = [ctx.Process(target=W.write_vpw, args=(sd, path)) for _ in range(6)]
    for p in procs:
        p.start()
    for p in procs:
        p.join()
        assert p.exitcode == 0
    raw = open(path, "rb").read()
    assert raw[:4] == b"VPW1" and len(raw) == 4 + 4 + 4 + 8 + 8 + 4 + 8 + 4 * (1 << 16)
    assert not [f for f in os.listdir(tmp_path) if f.endswith(".tmp")]
    assert W.cache_path_for("/x/y.pth") == W.cache_path_for("/x/y.pth")          # stable across calls / processes


def test_engine_create_fails_loudly_without_gpu(tmp_path):
    """No CPU fallback: on a box without a B200 the engine must refuse, not degrade."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from autoware_vision_pilot_b200 import engine as E
    with pytest.raises(RuntimeError):
        E.Engine([E.SCENE_SEG], [str(tmp_path / "missing.vpw")])


def test_ctypes_mirrors_match_the_c_struct_layouts(tmp_path):
    """The headers are the contract; the ctypes Structures in _lib.py / engine.py are hand-written mirrors.
    A C program compiled against include/*.h prints sizeof / offsetof of every struct and field, which must
    equal what ctypes computes — catches a field added on one side only."""
    import subprocess
    from autoware_vision_pilot_b200 import engine as E
    from autoware_vision_pilot_b200 import multicam as M
    mirrors = {
        "vp_tap_view": (E._TapView, {}),
        "vp_multicam_view": (M._View, {}),
        "vpb_conv_args": (L.ConvArgs, {"inp": "in"}),
        "vpb_lateral_state": (L.LateralState, {}),
        "vpb_lateral_out": (L.LateralOut, {}),
        "vp_engine_config": (E._Config, {}),
        "vp_output": (E._Output, {}),
        "vp_engine_stats": (E._Stats, {}),
    }
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "vp_b200.h"', '#include "vp_b200_ops.h"', '#include "vp_b200_multicam.h"',
             'int main(void) {']
    for cname, (cls, rename) in mirrors.items():
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {rename.get(fname, fname)}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = dict(l.split() for l in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for cname, (cls, _) in mirrors.items():
        assert int(out[cname]) == C.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(out[f"{cname}.{fname}"]) == getattr(cls, fname).offset, f"{cname}.{fname}"


def test_pillow_bilinear_tables_reproduce_pillow(tmp_path):
    """AutoSpeed letterbox (auto_speed_infer.py:38): the C++ coefficient tables for Pillow's BILINEAR (antialias) filter,
    run through the kernel's integer arithmetic in numpy, reproduce Image.resize(BILINEAR) bit for bit."""
    from PIL import Image
    lib = L.lib()
    lib.vpb_resize_tables_host.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                           C.c_int, C.POINTER(C.c_int)]

    def tables(in_size, out_size):
        bounds = (C.c_int * out_size)()
        coeffs = (C.c_int * (out_size * 64))()
        ks = C.c_int()
        L.check(lib.vpb_resize_tables_host(3, in_size, out_size, bounds, coeffs, out_size * 64, C.byref(ks)), "tables")
        return (np.frombuffer(bounds, dtype=np.int32).copy(),
                np.frombuffer(coeffs, dtype=np.int32)[: out_size * ks.value].reshape(out_size, ks.value).copy())

    rng = np.random.default_rng(4)
    for (h, w, oh, ow) in ((270, 480, 128, 227), (100, 130, 256, 333)):
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        xb, xk = tables(w, ow)
        yb, yk = tables(h, oh)

        def axis_pass(a, b, k, n_in):            # a [rows, n_in, 3] -> [rows, len(b), 3]
            out = np.zeros((a.shape[0], len(b), 3), np.int64)
            for o in range(len(b)):
                n = min(k.shape[1], n_in - b[o])
                out[:, o] = (a[:, b[o]:b[o] + n].astype(np.int64) * k[o, :n, None]).sum(1)
            return np.clip((out + (1 << 21)) >> 22, 0, 255).astype(np.uint8)

        hor = axis_pass(img, xb, xk, w)
        ver = axis_pass(hor.transpose(1, 0, 2), yb, yk, h).transpose(1, 0, 2)
        assert np.array_equal(ver, np.asarray(Image.fromarray(img).resize((ow, oh), Image.BILINEAR)))
