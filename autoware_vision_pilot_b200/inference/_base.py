"""Shared implementation of the four `*NetworkInfer` drop-in classes (boundary #1).

Mirrors Models/inference/{scene_seg,scene_3d,domain_seg,ego_lanes}_infer.py: same class names,
constructor argument, `inference(image)` signature, return dtype/shape and exceptions.  The body is
a call into libvp_b200.so — there is no PyTorch forward and no CPU fallback.

Extension beyond the reference (opt-in, default behaviour unchanged): `resize_mode` lets the
helper take the camera frame at its native size and do the caller-side resize
(Models/visualizations/SceneSeg/image_visualization.py:108-109) inside the fused GPU pre-process.
"""
from __future__ import annotations

import os
import tempfile

import numpy as np

from .. import engine as E
from .. import weights as W


def _resolve_checkpoint(path: str) -> str:
    if path.endswith(".vpw"):
        return path
    out = os.path.splitext(path)[0] + ".vpw"
    try:
        return W.convert_checkpoint(path, out)
    except (PermissionError, OSError):
        return W.convert_checkpoint(path, W.cache_path_for(path))


def _as_hwc_uint8(image) -> np.ndarray:
    """PIL.Image or HWC uint8 ndarray (what transforms.ToTensor accepts, ego_lanes_infer.py:53)."""
    if isinstance(image, np.ndarray):
        a = image
    else:
        a = np.asarray(image.convert("RGB") if getattr(image, "mode", "RGB") != "RGB" else image)
    if a.dtype != np.uint8 or a.ndim != 3 or a.shape[2] != 3:
        raise ValueError("image must be an RGB PIL image or an HWC uint8 array")
    return np.ascontiguousarray(a)


class NetworkInferBase:
    KIND = None
    REQUIRE_CHECKPOINT = True
    CHECK_SIZE = True

    def __init__(self, checkpoint_path: str = "", *, resize_mode: str = "none", precision: str = "fp16",
                 gpu_id: int = 0):
        if not checkpoint_path and self.REQUIRE_CHECKPOINT:
            # scene_seg_infer.py:32-33 (message kept, typo included)
            raise ValueError("No path to checkpiont file provided in class initialization")
        self.device = f"cuda:{gpu_id}"
        print(f"Using {self.device} for inference")
        self._resize_mode = resize_mode
        vpw = _resolve_checkpoint(checkpoint_path) if checkpoint_path else self._vanilla_checkpoint()
        self._engine = E.Engine([self.KIND], [vpw], gpu_id=gpu_id,
                                dtype=precision, resize_mode=E.RESIZE_BY_NAME[resize_mode],
                                convention=E.CONV_RGB, fetch_raw=True)

    def _vanilla_checkpoint(self) -> str:
        raise ValueError("No path to checkpiont file provided in class initialization")

    def _run(self, image) -> None:
        a = _as_hwc_uint8(image)
        h, w, _ = a.shape
        if self.CHECK_SIZE and self._resize_mode == "none" and (w != 640 or h != 320):
            # scene_seg_infer.py:40-42
            raise ValueError("Incorrect input size - input image must have height of 320px and width of 640px")
        self._engine.infer(a)
