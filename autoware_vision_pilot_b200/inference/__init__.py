"""Drop-in replacements for the reference's Models/inference helper classes."""
from .domain_seg_infer import DomainSegNetworkInfer  # noqa: F401
from .ego_lanes_infer import EgoLanesNetworkInfer  # noqa: F401
from .scene_3d_infer import Scene3DNetworkInfer  # noqa: F401
from .scene_seg_infer import SceneSegNetworkInfer  # noqa: F401
from .auto_speed_infer import AutoSpeedNetworkInfer  # noqa: F401
