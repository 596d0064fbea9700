"""Drop-in for Models/inference/domain_seg_infer.py (DomainSegNetworkInfer)."""
import numpy as np

from .. import engine as E
from ._base import NetworkInferBase


class DomainSegNetworkInfer(NetworkInferBase):
    KIND = E.DOMAIN_SEG

    def inference(self, image):
        """-> float32 [320,640,1] in {0,1} (domain_seg_infer.py:54-60); the `> 0` test is fused into
        the last convolution's epilogue."""
        self._run(image)
        return self._engine.cls(0).astype(np.float32)[..., None]
