"""ORACLE support (build-container only): import the UNMODIFIED reference modules.

/root/reference exists only in the build container, never on the GPU box, so nothing that runs
under `-m gpu`, smoke() or bench.py may import this file.  It is used by scripts/make_golden.py
(to generate tests/golden/*) and by tests/test_oracle_vs_reference.py (skipped when the reference
is absent) to pin oracle/net.py against the real thing.

The single patch: Models/model_components/backbone.py:9 requests ImageNet weights, which needs a
network download; `torchvision.models.efficientnet_b0` is wrapped to force `weights=None` before
the reference modules are imported.  The architecture is unchanged and every parameter is then
overwritten from the synthetic state_dict (scene_seg_infer.py:30-31 does the same from a file).
"""
from __future__ import annotations

import os
import sys

REFERENCE_ROOT = os.environ.get("VP_REFERENCE_ROOT", "/root/reference")
MODELS_DIR = os.path.join(REFERENCE_ROOT, "Models")


def available() -> bool:
    return os.path.isdir(os.path.join(MODELS_DIR, "model_components"))


_patched = False


def _patch():
    global _patched
    if _patched:
        return
    import torchvision

    orig = torchvision.models.efficientnet_b0

    def efficientnet_b0_no_download(*args, **kwargs):
        kwargs["weights"] = None
        if args:
            args = ()
        return orig(**kwargs)

    torchvision.models.efficientnet_b0 = efficientnet_b0_no_download
    if MODELS_DIR not in sys.path:
        sys.path.insert(0, MODELS_DIR)
    _patched = True


def build_network(model: str, state_dict):
    """Instantiate the reference nn.Module exactly like the *_infer.py helpers do and load the
    given state_dict (strict)."""
    _patch()
    from model_components.scene_seg_network import SceneSegNetwork
    if model == "scene_seg":
        m = SceneSegNetwork()                                   # scene_seg_infer.py:27
    elif model == "scene_3d":
        from model_components.scene_3d_network import Scene3DNetwork
        m = Scene3DNetwork(SceneSegNetwork())                   # scene_3d_infer.py:28-29
    elif model == "domain_seg":
        from model_components.domain_seg_network import DomainSegNetwork
        m = DomainSegNetwork(SceneSegNetwork())                 # domain_seg_infer.py:28-29
    elif model == "ego_lanes":
        from model_components.ego_lanes_network import EgoLanesNetwork
        m = EgoLanesNetwork()                                   # ego_lanes_infer.py:34
    else:
        raise ValueError(model)
    m.load_state_dict(state_dict, strict=True)
    return m.eval()


def infer_class(model: str):
    """The reference's Models/inference helper class (boundary #1)."""
    _patch()
    if model == "scene_seg":
        from inference.scene_seg_infer import SceneSegNetworkInfer as K
    elif model == "scene_3d":
        from inference.scene_3d_infer import Scene3DNetworkInfer as K
    elif model == "domain_seg":
        from inference.domain_seg_infer import DomainSegNetworkInfer as K
    else:
        from inference.ego_lanes_infer import EgoLanesNetworkInfer as K
    return K
