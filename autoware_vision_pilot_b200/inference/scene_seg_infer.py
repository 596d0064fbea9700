"""Drop-in for Models/inference/scene_seg_infer.py (SceneSegNetworkInfer)."""
import numpy as np

from .. import engine as E
from ._base import NetworkInferBase


class SceneSegNetworkInfer(NetworkInferBase):
    KIND = E.SCENE_SEG

    def inference(self, image):
        """-> int64 [320,640] class indices {0 bg, 1 fg, 2 road} (scene_seg_infer.py:52-57); the
        argmax itself is fused into the last convolution's epilogue."""
        self._run(image)
        return self._engine.cls(0).astype(np.int64)
