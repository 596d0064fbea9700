"""Drop-in for Models/inference/ego_lanes_infer.py (EgoLanesNetworkInfer)."""
import os
import tempfile

import numpy as np

from .. import engine as E
from .. import weights as W
from ._base import NetworkInferBase


class EgoLanesNetworkInfer(NetworkInferBase):
    """Same boundary as the reference helper (ego_lanes_infer.py:9-62):

    * `checkpoint_path=""` is ALLOWED and gives the "vanilla" (randomly initialised) network
      (ego_lanes_infer.py:34-44) — here the EgoLanesNetwork state_dict layout filled with PyTorch's default
      initialisers' distributions (weights.vanilla_ego_lanes_state_dict);
    * `inference(image)` takes a PIL image or an HWC uint8 array and performs NO explicit size check
      (ego_lanes_infer.py:51-62).  The reference network itself only runs at 320x640 (its context block does
      `reshape([10, 20])`, auto_steer_context.py:44, and torch raises a RuntimeError for any other size);
      the same RuntimeError is raised here, unless the helper was built with a `resize_mode`.
    """
    KIND = E.EGO_LANES
    REQUIRE_CHECKPOINT = False
    CHECK_SIZE = False

    def _vanilla_checkpoint(self) -> str:
        print("Loading vanilla AutoSteer model for training")          # ego_lanes_infer.py:44
        d = tempfile.mkdtemp(prefix="vpb_vanilla_")
        return W.write_vpw(W.vanilla_ego_lanes_state_dict(), os.path.join(d, "ego_lanes_vanilla.vpw"))

    def _run(self, image):
        from ._base import _as_hwc_uint8
        a = _as_hwc_uint8(image)
        if self._resize_mode == "none" and a.shape[:2] != (320, 640):
            raise RuntimeError(f"shape '[10, 20]' is invalid for the context of a {a.shape[0]}x{a.shape[1]} input "
                               "(EgoLanesNetwork only runs at 320x640, auto_steer_context.py:44)")
        self._engine.infer(a)

    def inference(self, image):
        """-> float32 [3,80,160] raw logits (ego_lanes_infer.py:60)."""
        self._run(image)
        return self._engine.raw(0).copy()
