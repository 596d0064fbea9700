"""Host-side mirror of the production lateral post-process that follows EgoLanes:
LaneFilter::update (production_release/src/lane_filtering/lane_filter.cpp:232-323) followed by
LaneTracker::update (src/lane_tracking/lane_tracking.cpp:36-300), executed by ONE device kernel
(csrc/lateral.cu) on the EgoLanes masks while they are still in HBM.  State (previous fits, BEV lane
width history) lives on the device between frames, like the members of the two reference classes.

torch is used only to own the two small device buffers.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib as L


class LateralPostProcess:
    """`update(masks, steering)` == LaneFilter.update -> LaneTracker.update -> (if the BEV lines are valid)
    PathFinder.update of the reference's lateral thread (production_release/main.cpp:540-577)."""

    def __init__(self, image_size=(1920, 1080), smoothing_factor: float = 0.5,
                 homography: Optional[Sequence[float]] = None, device: str = "cuda:0"):
        self._lib = L.lib()
        self._lib.vpb_lateral_init.argtypes = [C.c_void_p, C.c_void_p]
        self._lib.vpb_lateral_update.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                                                 C.POINTER(C.c_double), C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
        self.image_size = tuple(image_size)
        self.smoothing = float(smoothing_factor)
        self._hom = (C.c_double * 9)(*homography) if homography is not None else None
        self._state = torch.zeros(C.sizeof(L.LateralState), dtype=torch.uint8, device=device)
        self._out = torch.zeros(C.sizeof(L.LateralOut), dtype=torch.uint8, device=device)
        self.reset()

    def reset(self) -> None:
        """LaneFilter::reset() + a fresh LaneTracker."""
        L.check(self._lib.vpb_lateral_init(self._state.data_ptr(), None), "vpb_lateral_init")

    def update_device(self, masks_ptr: int, height: int = 80, width: int = 160, stream: int = 0,
                      autosteer_steering_rad: float = 0.0) -> None:
        """Enqueue one frame: masks_ptr = device float [3][height][width] (ego_left, ego_right, other)."""
        L.check(self._lib.vpb_lateral_update(masks_ptr, height, width, self.image_size[0], self.image_size[1],
                                             self.smoothing, self._hom, float(autosteer_steering_rad),
                                             self._state.data_ptr(),
                                             self._out.data_ptr(), stream or None), "vpb_lateral_update")

    def result(self) -> dict:
        """Copy the vpb_lateral_out record to the host (synchronises) and return it as a dict."""
        raw = self._out.cpu().numpy().tobytes()
        o = L.LateralOut.from_buffer_copy(raw)
        d = {}
        for name, ctype in L.LateralOut._fields_:
            v = getattr(o, name)
            d[name] = np.array(v[:]) if hasattr(v, "__len__") else v
        return d

    def update(self, masks: torch.Tensor, autosteer_steering_rad: float = 0.0) -> dict:
        m = masks.contiguous()
        assert m.dtype == torch.float32 and m.dim() == 3 and m.shape[0] == 3 and m.is_cuda
        self.update_device(m.data_ptr(), m.shape[1], m.shape[2], autosteer_steering_rad=autosteer_steering_rad)
        return self.result()
