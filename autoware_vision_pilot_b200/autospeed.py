"""Python face of the AutoSpeed C-ABI (include/vp_b200_autospeed.h) — a thin ctypes wrapper, no compute."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _lib as L

NUM_ANCHORS, NUM_OUT = 10752, 8
_bound = False


def _bind():
    global _bound
    lib = L.lib()
    if _bound:
        return lib
    lib.vp_autospeed_create.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]
    lib.vp_autospeed_destroy.argtypes = [C.c_void_p]
    lib.vp_autospeed_destroy.restype = None
    lib.vp_autospeed_set_thresholds.argtypes = [C.c_void_p, C.c_float, C.c_float]
    lib.vp_autospeed_infer.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.vp_autospeed_infer_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
    lib.vp_autospeed_sync.argtypes = [C.c_void_p, C.c_int]
    lib.vp_autospeed_detections.argtypes = [C.c_void_p, C.POINTER(C.POINTER(C.c_float)), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.vp_autospeed_raw.argtypes = [C.c_void_p, C.POINTER(C.POINTER(C.c_float)), C.POINTER(C.c_void_p), C.POINTER(C.c_int),
                                     C.POINTER(C.c_int)]
    lib.vp_autospeed_stats.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_double)]
    lib.vp_autospeed_read_tap.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_long, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                          C.POINTER(C.c_int)]
    lib.vp_autospeed_read_tap.restype = C.c_long
    _bound = True
    return lib


class AutoSpeedEngine:
    def __init__(self, weights_vpw: str, *, gpu_id: int = 0, dtype: str = "fp16", stream: Optional[int] = None):
        self._lib = _bind()
        self._h = C.c_void_p()
        L.check(self._lib.vp_autospeed_create(weights_vpw.encode(), gpu_id, L.VPB_BF16 if dtype == "bf16" else L.VPB_F16,
                                              stream, C.byref(self._h)), "vp_autospeed_create")

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.vp_autospeed_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def set_thresholds(self, conf: float = 0.6, iou: float = 0.45) -> None:
        L.check(self._lib.vp_autospeed_set_thresholds(self._h, conf, iou), "vp_autospeed_set_thresholds")

    def infer(self, frame: np.ndarray, fetch_raw: bool = False) -> np.ndarray:
        """frame uint8 [h, w, 3] RGB (any size) -> detections float32 [n, 6] = x1, y1, x2, y2, score, class."""
        if not isinstance(frame, np.ndarray) or frame.dtype != np.uint8 or frame.ndim != 3 or frame.shape[2] != 3:
            raise ValueError("frame must be uint8 [h, w, 3]")
        if frame.strides[2] != 1 or frame.strides[1] != 3:
            frame = np.ascontiguousarray(frame)
        h, w, _ = frame.shape
        L.check(self._lib.vp_autospeed_infer(self._h, frame.ctypes.data, h, w, frame.strides[0], int(fetch_raw)),
                "vp_autospeed_infer")
        return self.detections()

    def infer_device(self, dev_ptr: int, h: int, w: int, stride: int) -> None:
        L.check(self._lib.vp_autospeed_infer_device(self._h, dev_ptr, h, w, stride), "vp_autospeed_infer_device")

    def sync(self, fetch: int = 1) -> None:
        L.check(self._lib.vp_autospeed_sync(self._h, fetch), "vp_autospeed_sync")

    def detections(self) -> np.ndarray:
        det, n, nc = C.POINTER(C.c_float)(), C.c_int(), C.c_int()
        L.check(self._lib.vp_autospeed_detections(self._h, C.byref(det), C.byref(n), C.byref(nc)), "vp_autospeed_detections")
        self.n_candidates = nc.value
        if n.value == 0:
            return np.zeros((0, 6), np.float32)
        return np.ctypeslib.as_array(det, shape=(n.value, 6)).copy()

    def raw(self) -> np.ndarray:
        """[8, 10752] float32 (host copy made by infer(fetch_raw=True) / sync(2))."""
        rh, ch, na = C.POINTER(C.c_float)(), C.c_int(), C.c_int()
        L.check(self._lib.vp_autospeed_raw(self._h, C.byref(rh), None, C.byref(ch), C.byref(na)), "vp_autospeed_raw")
        return np.ctypeslib.as_array(rh, shape=(ch.value, na.value)).copy()

    def stats(self) -> dict:
        n, f = C.c_int(), C.c_double()
        L.check(self._lib.vp_autospeed_stats(self._h, C.byref(n), C.byref(f)), "vp_autospeed_stats")
        return {"n_launches": n.value, "flops": f.value}

    def read_tap(self, name: str) -> np.ndarray:
        c, h, w = C.c_int(), C.c_int(), C.c_int()
        n = self._lib.vp_autospeed_read_tap(self._h, name.encode(), None, 0, C.byref(c), C.byref(h), C.byref(w))
        if n < 0:
            raise RuntimeError(L.last_error())
        buf = np.empty((c.value, h.value, w.value), dtype=np.float32)
        if self._lib.vp_autospeed_read_tap(self._h, name.encode(), buf.ctypes.data, buf.size, None, None, None) < 0:
            raise RuntimeError(L.last_error())
        return buf
