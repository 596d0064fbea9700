"""Drop-in for Models/inference/scene_3d_infer.py (Scene3DNetworkInfer)."""
import numpy as np

from .. import engine as E
from ._base import NetworkInferBase


class Scene3DNetworkInfer(NetworkInferBase):
    KIND = E.SCENE_3D

    def inference(self, image):
        """-> float32 [320,640,1] raw relative depth (scene_3d_infer.py:54-58)."""
        self._run(image)
        return np.ascontiguousarray(self._engine.raw(0).transpose(1, 2, 0))
