"""ORACLE (test infrastructure, never shipped, never timed as the product).

CPU fp32 restatement of the four hot-path networks of autoware_vision_pilot, written as plain
functions over a `state_dict` so that it travels to the GPU box (where /root/reference does not
exist).  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / --impl reference
legs may import this module.

Every function cites the reference lines it restates (paths relative to the reference repo).
The EfficientNet-B0 trunk is third-party code (torchvision, unpinned `torchvision>=0.22.0` in
Models/requirements.txt; 0.26.0 installed here) reached through
Models/model_components/backbone.py:9 — its published architecture is restated from the module
dump in SURVEY.md Appendix A and validated against torchvision in tests/test_oracle_vs_reference.py.

Parity pin: the reference ships no golden vectors, no weights and no tests for this path
 — "parity unpinned" by the reference itself.  The pin used instead is the
reference's own modules imported from /root/reference in the build container
(oracle/ref_import.py), run on the committed synthetic weights; their outputs are committed under
tests/golden/ by scripts/make_golden.py and this restatement is checked against them.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]

# (expand_ratio, kernel, stride, cin, cout, repeats) for encoder[1..7]  — SURVEY Appendix A
MBCONV_STAGES = [
    (1, 3, 1, 32, 16, 1),
    (6, 3, 2, 16, 24, 2),
    (6, 5, 2, 24, 40, 2),
    (6, 3, 2, 40, 80, 3),
    (6, 5, 1, 80, 112, 3),
    (6, 5, 2, 112, 192, 4),
    (6, 3, 1, 192, 320, 1),
]
BN_EPS = 1e-5

# state_dict prefixes per network (scene_seg_network.py:12-21,
# scene_3d_network.py:13-22, domain_seg_upstream.py:11-19, ego_lanes_network.py:15-27)
PREFIX = {
    "scene_seg": dict(enc="Backbone.encoder.", ctx="SceneContext.", neck="SceneNeck.",
                      head="SceneSegHead."),
    "scene_3d": dict(enc="PreTrainedBackbone.pretrainedBackBone.encoder.", ctx="DepthContext.",
                     neck="DepthNeck.", head="SuperDepthHead."),
    "domain_seg": dict(enc="DomainSegUpstream.pretrainedBackBone.encoder.",
                       ctx="DomainSegUpstream.pretrainedContext.",
                       neck="DomainSegUpstream.pretrainedNeck.", head="DomainSegHead."),
    "ego_lanes": dict(enc="BEVBackbone.encoder.", ctx="AutoSteerContext.", neck="EgopathNeck.",
                      head="EgoLanesHead."),
}
MODELS = tuple(PREFIX.keys())

# A hook lets the calibration pass (oracle/synth.py) rescale a weight the first time it is used.
Hook = Optional[Callable[[str, torch.Tensor, torch.Tensor], torch.Tensor]]


class Ctx:
    """Forward-pass context: the state_dict, an optional weight hook and a tap recorder."""

    def __init__(self, sd: SD, hook: Hook = None, taps: Optional[dict] = None):
        self.sd, self.hook, self.taps = sd, hook, taps

    def w(self, name: str, x: torch.Tensor) -> torch.Tensor:
        w = self.sd[name]
        if self.hook is not None:
            w = self.hook(name, w, x)
        return w

    def tap(self, name: str, t: torch.Tensor) -> None:
        if self.taps is not None:
            self.taps[name] = t


def _bn(c: Ctx, p: str, x: torch.Tensor) -> torch.Tensor:
    sd = c.sd
    return F.batch_norm(x, sd[p + "running_mean"], sd[p + "running_var"], sd[p + "weight"],
                        sd[p + "bias"], False, 0.0, BN_EPS)


def _conv_bn(c: Ctx, p: str, x, stride=1, groups=1, act=True):
    """torchvision Conv2dNormActivation: bias-free conv, BatchNorm(eps 1e-5), optional SiLU."""
    w = c.w(p + "0.weight", x)
    k = w.shape[-1]
    y = F.conv2d(x, w, None, stride=stride, padding=(k - 1) // 2, groups=groups)
    y = _bn(c, p + "1.", y)
    return F.silu(y) if act else y


def _se(c: Ctx, p: str, x):
    """torchvision SqueezeExcitation: GAP -> 1x1 -> SiLU -> 1x1 -> sigmoid -> scale."""
    s = F.adaptive_avg_pool2d(x, 1)
    s = F.silu(F.conv2d(s, c.w(p + "fc1.weight", s), c.sd[p + "fc1.bias"]))
    s = torch.sigmoid(F.conv2d(s, c.w(p + "fc2.weight", s), c.sd[p + "fc2.bias"]))
    return x * s


def backbone(c: Ctx, p: str, image: torch.Tensor) -> List[torch.Tensor]:
    """Backbone.forward (backbone.py:11-22): EfficientNet-B0 features, taps l0,l2,l3,l4,l8."""
    x = _conv_bn(c, p + "0.", image, stride=2)                       # encoder[0]
    outs = [x]
    for si, (exp, k, stride, cin, cout, reps) in enumerate(MBCONV_STAGES, start=1):
        for r in range(reps):
            bp = f"{p}{si}.{r}.block."
            s = stride if r == 0 else 1
            inp = x
            i = 0
            if exp != 1:
                x = _conv_bn(c, f"{bp}{i}.", x)                       # 1x1 expand
                i += 1
            x = _conv_bn(c, f"{bp}{i}.", x, stride=s, groups=x.shape[1])  # depthwise
            x = _se(c, f"{bp}{i + 1}.", x)
            x = _conv_bn(c, f"{bp}{i + 2}.", x, act=False)           # 1x1 project
            if s == 1 and inp.shape[1] == x.shape[1]:
                x = x + inp                                          # StochasticDepth == id (eval)
        outs.append(x)
    x = _conv_bn(c, p + "8.", x)                                      # encoder[8] 320->1280
    outs.append(x)
    feats = [outs[0], outs[2], outs[3], outs[4], outs[8]]
    for i, f in enumerate(feats):
        c.tap(f"f{i}", f)
    return feats


def feature_fusion(feats: List[torch.Tensor]) -> torch.Tensor:
    """BackboneFeatureFusion.forward (backbone_feature_fusion.py:13-38)."""
    outs = []
    for f, n in zip(feats[:4], (4, 3, 2, 1)):
        for _ in range(n):
            f = F.max_pool2d(f, 2, 2)
        outs.append(f)
    outs.append(feats[4])
    return torch.cat(outs, 1)


def context(c: Ctx, p: str, features: torch.Tensor) -> torch.Tensor:
    """SceneContext / DepthContext / AutoSteerContext.forward
    (scene_context.py:25-57, depth_context.py:25-57, auto_steer_context.py:28-60).
    Dropout is identity in eval()."""
    sd = c.sd
    v = torch.mean(features, dim=[2, 3])
    c0 = F.gelu(F.linear(v, c.w(p + "context_layer_0.weight", v), sd[p + "context_layer_0.bias"]))
    c1 = F.gelu(F.linear(c0, c.w(p + "context_layer_1.weight", c0), sd[p + "context_layer_1.bias"]))
    c2 = torch.sigmoid(F.linear(c1, c.w(p + "context_layer_2.weight", c1), sd[p + "context_layer_2.bias"]))
    x = c2.reshape([10, 20]).unsqueeze(0).unsqueeze(0)
    c.tap("ctx_map", x)
    for i in (3, 4, 5, 6):
        x = F.gelu(F.conv2d(x, c.w(f"{p}context_layer_{i}.weight", x), sd[f"{p}context_layer_{i}.bias"],
                            padding=1))
    out = x * features + features
    c.tap("context", out)
    return out


def _conv3(c: Ctx, name: str, x, act=True):
    y = F.conv2d(x, c.w(name + ".weight", x), c.sd[name + ".bias"], padding=1)
    return F.gelu(y) if act else y


def _up_skip(c: Ctx, p: str, i: int, x, skip):
    """ConvTranspose2d(k2,s2) + Conv1x1(skip) summed before any activation
    (scene_neck.py:30-32)."""
    y = F.conv_transpose2d(x, c.w(f"{p}upsample_layer_{i}.weight", x), c.sd[f"{p}upsample_layer_{i}.bias"],
                           stride=2)
    if skip is not None:
        y = y + F.conv2d(skip, c.w(f"{p}skip_link_layer_{i}.weight", skip),
                         c.sd[f"{p}skip_link_layer_{i}.bias"])
    return y


def neck(c: Ctx, p: str, ctx: torch.Tensor, feats: List[torch.Tensor]) -> torch.Tensor:
    """SceneNeck / Scene3DNeck / EgoPathNeck.forward (scene_neck.py:26-60)."""
    d = _up_skip(c, p, 0, ctx, feats[3])
    d = _conv3(c, p + "decode_layer_0", d)
    d = _conv3(c, p + "decode_layer_1", d)
    d = _up_skip(c, p, 1, d, feats[2])
    d = _conv3(c, p + "decode_layer_2", d)
    d = _conv3(c, p + "decode_layer_3", d)
    d = _up_skip(c, p, 2, d, feats[1])
    d = _conv3(c, p + "decode_layer_4", d)
    d = _conv3(c, p + "decode_layer_5", d)
    c.tap("neck", d)
    return d


def seg_head(c: Ctx, p: str, nk: torch.Tensor, feats: List[torch.Tensor]) -> torch.Tensor:
    """SceneSegHead / Scene3DHead / DomainSegHead.forward
    (scene_seg_head.py:21-44, scene_3d_head.py:21-47, domain_seg_head.py:21-44)."""
    d = _up_skip(c, p, 3, nk, feats[0])
    d = _conv3(c, p + "decode_layer_6", d)
    d = _conv3(c, p + "decode_layer_7", d)
    d = _up_skip(c, p, 4, d, None)
    d = _conv3(c, p + "decode_layer_8", d)
    d = _conv3(c, p + "decode_layer_9", d)
    c.tap("d9", d)
    return _conv3(c, p + "decode_layer_10", d, act=False)


def ego_lanes_head(c: Ctx, p: str, nk: torch.Tensor) -> torch.Tensor:
    """EgoLanesHead.forward (ego_lanes_head.py:17-26)."""
    d = _conv3(c, p + "decode_layer_6", nk)
    d = _conv3(c, p + "decode_layer_7", d)
    return _conv3(c, p + "decode_layer_8", d, act=False)


@torch.no_grad()
def forward(model: str, sd: SD, image: torch.Tensor, hook: Hook = None,
            taps: Optional[dict] = None) -> torch.Tensor:
    """image: fp32 [1,3,320,640] normalised tensor -> raw network output.
    scene_seg_network.py:24-29, scene_3d_network.py:25-31, domain_seg_network.py:17-20,
    ego_lanes_network.py:30-37."""
    pf = PREFIX[model]
    c = Ctx(sd, hook, taps)
    feats = backbone(c, pf["enc"], image)
    if model == "ego_lanes":
        fused = feature_fusion(feats)
        c.tap("fused", fused)
        ctx = context(c, pf["ctx"], fused)
        nk = neck(c, pf["neck"], ctx, feats)
        out = ego_lanes_head(c, pf["head"], nk)
    else:
        ctx = context(c, pf["ctx"], feats[4])
        nk = neck(c, pf["neck"], ctx, feats)
        out = seg_head(c, pf["head"], nk, feats)
    c.tap("out", out)
    return out


# ---------------------------------------------------------------------------------------------
# Pre- and post-processing of the Python boundary (Models/inference/*_infer.py)
# ---------------------------------------------------------------------------------------------
MEAN = (0.485, 0.456, 0.406)
STD = (0.229, 0.224, 0.225)


def to_tensor_normalize(img_u8_hwc) -> torch.Tensor:
    """transforms.ToTensor + Normalize (scene_seg_infer.py:15-20,44-45): uint8 HWC -> fp32
    [1,3,H,W]; x/255 then (x-mean)/std, in that order, in fp32."""
    x = torch.from_numpy(img_u8_hwc).permute(2, 0, 1).contiguous().to(torch.float32).div(255)
    mean = torch.tensor(MEAN, dtype=torch.float32).view(3, 1, 1)
    std = torch.tensor(STD, dtype=torch.float32).view(3, 1, 1)
    return ((x - mean) / std).unsqueeze(0)


def postprocess(model: str, out: torch.Tensor):
    """scene_seg_infer.py:52-57 (argmax, int64 [320,640]); scene_3d_infer.py:54-58 (fp32
    [320,640,1]); domain_seg_infer.py:54-60 (0/1 fp32 [320,640,1]); ego_lanes_infer.py:60
    (raw [3,80,160])."""
    p = out.squeeze(0).detach()
    if model == "ego_lanes":
        return p.numpy()
    p = p.permute(1, 2, 0)
    if model == "scene_seg":
        return torch.max(p, dim=2)[1].numpy()
    o = p.numpy().copy()
    if model == "domain_seg":
        o[o <= 0] = 0.0
        o[o > 0] = 1.0
    return o


def preprocess_cpp_generic(bgr_u8_hwc, resize_fn) -> torch.Tensor:
    """TensorRTBackend::preprocess (VisionPilot/middleware_recipes/common/backends/
    tensorrt_backend.cpp:160-177): cv::resize (INTER_LINEAR) on BGR, convertTo(1/255), subtract
    Scalar(0.406,0.456,0.485), divide Scalar(0.225,0.224,0.229) — BGR-ordered stats, NO channel
    swap (tensor channel 0 = blue) — then split to CHW."""
    import numpy as np
    small = resize_fn(bgr_u8_hwc, 640, 320)
    x = small.astype(np.float32) * np.float32(1.0 / 255.0)
    mean = np.array([0.406, 0.456, 0.485], dtype=np.float32)
    std = np.array([0.225, 0.224, 0.229], dtype=np.float32)
    x = (x - mean) / std
    return torch.from_numpy(np.ascontiguousarray(x.transpose(2, 0, 1))).unsqueeze(0)


def preprocess_cpp_egolanes(bgr_u8_hwc, resize_fn) -> torch.Tensor:
    """EgoLanesTensorRTEngine::preprocessEgoLanes (VisionPilot/production_release/src/inference/
    tensorrt_engine.cpp:190-220): INTER_LINEAR resize, BGR->RGB, convertTo(1/255),
    (x - MEAN[c]) / STD[c] with RGB stats, CHW."""
    import numpy as np
    small = resize_fn(bgr_u8_hwc, 640, 320)[..., ::-1]
    x = small.astype(np.float32) * np.float32(1.0 / 255.0)
    x = (x - np.array(MEAN, dtype=np.float32)) / np.array(STD, dtype=np.float32)
    return torch.from_numpy(np.ascontiguousarray(x.transpose(2, 0, 1))).unsqueeze(0)


def ego_lanes_masks(raw, threshold: float = 0.0):
    """EgoLanesTensorRTEngine::postProcess (tensorrt_engine.cpp:264-305): three float masks
    (v > threshold ? 1 : 0) — and the class-id rule of createEgoLanesMaskKernel
    (common/visualizers/cuda_visualization_kernels.cu:45-75): other > right > left, else 255."""
    import numpy as np
    r = np.asarray(raw)
    masks = (r > threshold).astype(np.float32)
    ids = np.full(r.shape[1:], 255, dtype=np.uint8)
    ids[r[0] > 0] = 0
    ids[r[1] > 0] = 1
    ids[r[2] > 0] = 2
    return masks, ids


def seg_mask_255(raw):
    """createMaskKernel / the CPU fallback of RunModelNode::onImage
    (cuda_visualization_kernels.cu:13-42, ROS2/models/src/run_model_node.cpp:148-172): argmax with
    strict '>' from -1e9 (first max wins), class 1 -> 255 else 0; single channel: v > 0 -> 255."""
    import numpy as np
    r = np.asarray(raw)
    if r.shape[0] > 1:
        return np.where(np.argmax(r, axis=0) == 1, 255, 0).astype(np.uint8)
    return np.where(r[0] > 0, 255, 0).astype(np.uint8)
