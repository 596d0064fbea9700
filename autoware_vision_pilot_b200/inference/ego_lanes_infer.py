"""Drop-in for Models/inference/ego_lanes_infer.py (EgoLanesNetworkInfer)."""
import numpy as np

from .. import engine as E
from ._base import NetworkInferBase


class EgoLanesNetworkInfer(NetworkInferBase):
    """The reference accepts an empty checkpoint path and then runs randomly initialised weights
    "for training" (ego_lanes_infer.py:34-44).  An inference engine has nothing meaningful to do
    in that case, so an empty path is rejected like the other three helpers."""
    KIND = E.EGO_LANES

    def _run(self, image):
        from ._base import _as_hwc_uint8
        a = _as_hwc_uint8(image)
        if self._resize_mode == "none" and a.shape[:2] != (320, 640):
            # the reference has no size check (ego_lanes_infer.py:51-62) but its network only
            # works at 320x640 (reshape([10,20]), auto_steer_context.py:44)
            raise ValueError("Incorrect input size - input image must have height of 320px and width of 640px")
        self._engine.infer(a)

    def inference(self, image):
        """-> float32 [3,80,160] raw logits (ego_lanes_infer.py:60)."""
        self._run(image)
        return self._engine.raw(0).copy()
