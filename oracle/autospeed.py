"""ORACLE (test infrastructure, never shipped, never timed as the product).

CPU fp32 restatement of the AutoSpeed detector as plain functions over a `state_dict`, so that
it travels to the GPU box (where /root/reference does not exist).  Only `tests/`, `bench.py`'s cpu_baseline leg and the
golden-generating scripts may import this module.

Restates (paths relative to the reference repo):
  network   Models/model_components/auto_speed/auto_speed_network.py:34-50 (YOLO = backbone + neck + head),
            auto_speed_backbone.py:9-48, auto_speed_neck.py:7-24, auto_speed_head.py:25-68 (DFL decode, anchors),
            building blocks Models/model_components/common_layers.py (Conv :5-17, Residual :20-27, C3K :158-173,
            C3K2 :176-191, CTX :194-239, SPPF :242-254, Attention :77-104, PSABlock :107-118, C2PSA :257-269, DFL :141-155)
  helper    Models/inference/auto_speed_infer.py:16-108 (letterbox to 1024x512 with Pillow BILINEAR + gray 114 padding,
            ToTensor, second sigmoid + confidence 0.6 filter, cx/cy/w/h -> x1/y1/x2/y2, class-agnostic NMS at IoU 0.45
            (torchvision.ops.nms: third-party, restated from its published greedy algorithm), un-letterbox + clamp)
Variant 'n' (width [3,16,32,64,128,256], depth 1, csp [False, True]) with num_classes = 4 — the configuration of
auto_steer/auto_speed inference scripts (auto_speed_network.py:55-60).

Parity pin: tests/test_oracle_autospeed_vs_reference.py (build container) strict-loads the synthetic state_dict into the
UNMODIFIED reference module and compares raw predictions, and the reference helper's post-process on them.
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import numpy as np
import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]
WIDTH = (3, 16, 32, 64, 128, 256)
NC, DFL_CH = 4, 16
IMG_W, IMG_H = 1024, 512
BN_EPS = 1e-3                       # common_layers.py:10
STRIDES = (8.0, 16.0, 32.0)


# Calibration hook (synthetic weights only): called with (layer prefix, pre-activation output) the first time a layer
# runs; returns the factor the layer's affine parameters were rescaled by (see calibrate()).
_CAL = None


def _cal(name, y, sd, keys):
    if _CAL is None:
        return y
    f = _CAL(name, y)
    for k in keys:
        sd[k] = sd[k] * f
    return y * f


# ------------------------------------------------------------------ building blocks
def conv(sd: SD, p: str, x, k=1, s=1, pad=0, g=1, act=True):
    """Conv (common_layers.py:5-17): Conv2d(bias=False) + BatchNorm2d(eps 1e-3) + SiLU / Identity."""
    y = F.conv2d(x, sd[p + ".conv.weight"], None, s, pad, 1, g)
    y = F.batch_norm(y, sd[p + ".norm.running_mean"], sd[p + ".norm.running_var"], sd[p + ".norm.weight"],
                     sd[p + ".norm.bias"], False, 0.0, BN_EPS)
    y = _cal(p, y, sd, (p + ".norm.weight", p + ".norm.bias"))
    return F.silu(y) if act else y


def plain(sd: SD, p: str, x, pad=0, cal_bias=True):
    """nn.Conv2d with bias (CTX layers, the heads' last 1x1 convs).  cal_bias=False: the calibration pass rescales the
    weight only (class logits keep their bias)."""
    if cal_bias or _CAL is None:
        y = F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], 1, pad)
        return _cal(p, y, sd, (p + ".weight", p + ".bias"))
    y = _cal(p, F.conv2d(x, sd[p + ".weight"], None, 1, pad), sd, (p + ".weight",))
    return y + sd[p + ".bias"].view(1, -1, 1, 1)


def residual(sd, p, x):                                     # common_layers.py:20-27 (e folded into the weight shapes)
    return x + conv(sd, p + ".conv2", conv(sd, p + ".conv1", x, 3, 1, 1), 3, 1, 1)


def c3k(sd, p, x):                                          # :158-173
    y = residual(sd, p + ".res_m.1", residual(sd, p + ".res_m.0", conv(sd, p + ".conv1", x)))
    return conv(sd, p + ".conv3", torch.cat((y, conv(sd, p + ".conv2", x)), 1))


def c3k2(sd, p, x, csp: bool):                              # :176-191 (n = 1)
    y = list(conv(sd, p + ".conv1", x).chunk(2, 1))
    y.append(c3k(sd, p + ".res_m.0", y[-1]) if csp else residual(sd, p + ".res_m.0", y[-1]))
    return conv(sd, p + ".conv2", torch.cat(y, 1))


def ctx(sd, p, x, h, w):                                    # :194-239
    b = x.shape[0]
    y = x.mean(dim=(2, 3), keepdim=True)
    e = F.conv1d(y.squeeze(-1), sd[p + ".exp0.weight"], sd[p + ".exp0.bias"], 1, 1)
    e = _cal(p + ".exp0", e, sd, (p + ".exp0.weight", p + ".exp0.bias"))
    c1 = F.silu(F.silu(e).view(b, 1, h, w))
    c2 = F.silu(plain(sd, p + ".ctx0", c1, 1))
    c4 = F.silu(plain(sd, p + ".ctx1", c2, 1))
    c4 = c4 * x + x
    return plain(sd, p + ".ctx2", F.silu(c4), 1)


def sppf(sd, p, x):                                         # :242-254
    x = conv(sd, p + ".cv1", x)
    y1 = F.max_pool2d(x, 5, 1, 2)
    y2 = F.max_pool2d(y1, 5, 1, 2)
    return conv(sd, p + ".cv2", torch.cat((x, y1, y2, F.max_pool2d(y2, 5, 1, 2)), 1))


def attention(sd, p, x, num_head):                          # :77-104
    b, c, h, w = x.shape
    dim_head = c // num_head
    dim_key = dim_head // 2
    qkv = conv(sd, p + ".qkv", x, act=False).view(b, num_head, dim_key * 2 + dim_head, h * w)
    q, k, v = qkv.split([dim_key, dim_key, dim_head], dim=2)
    attn = ((q.transpose(-2, -1) @ k) * dim_key ** -0.5).softmax(dim=-1)
    o = (v @ attn.transpose(-2, -1)).view(b, c, h, w) + conv(sd, p + ".conv1", v.reshape(b, c, h, w), 3, 1, 1, g=c, act=False)
    return conv(sd, p + ".conv2", o, act=False)


def psablock(sd, p, x, num_head):                           # :107-118
    x = x + attention(sd, p + ".conv1", x, num_head)
    return x + conv(sd, p + ".conv2.1", conv(sd, p + ".conv2.0", x), act=False)


def c2psa(sd, p, x):                                        # :257-269
    c_ = x.shape[1] // 2
    a, y = conv(sd, p + ".cv1", x).split((c_, c_), 1)
    return conv(sd, p + ".cv2", torch.cat((a, psablock(sd, p + ".middle_block", y, c_ // 64)), 1))


# ------------------------------------------------------------------ network
def backbone(sd, x, taps=None):                             # auto_speed_backbone.py:41-48
    p1 = conv(sd, "net.p1", x, 3, 2, 1)
    p2 = ctx(sd, "net.p2.1", conv(sd, "net.p2.0", p1, 3, 2, 1), 128, 256)
    p3 = ctx(sd, "net.p3.1", conv(sd, "net.p3.0", p2, 3, 2, 1), 64, 128)
    p4 = ctx(sd, "net.p4.1", conv(sd, "net.p4.0", p3, 3, 2, 1), 32, 64)
    q = ctx(sd, "net.p5.1", conv(sd, "net.p5.0", p4, 3, 2, 1), 16, 32)
    s = sppf(sd, "net.p5.2", q)
    p5 = c2psa(sd, "net.p5.3", s)
    if taps is not None:
        taps.update(p1=p1, p2=p2, p3=p3, p4=p4, p5_ctx=q, p5_sppf=s, p5=p5)
    return p3, p4, p5


def neck(sd, feats, taps=None):                             # auto_speed_neck.py:17-24
    p3, p4, p5 = feats
    up = lambda t: F.interpolate(t, scale_factor=2.0, mode="nearest")
    p4 = c3k2(sd, "fpn.h1", torch.cat((up(p5), p4), 1), False)
    p3 = c3k2(sd, "fpn.h2", torch.cat((up(p4), p3), 1), False)
    p4 = c3k2(sd, "fpn.h4", torch.cat((conv(sd, "fpn.h3", p3, 3, 2, 1), p4), 1), False)
    p5 = c3k2(sd, "fpn.h6", torch.cat((conv(sd, "fpn.h5", p4, 3, 2, 1), p5), 1), True)
    if taps is not None:
        taps.update(n3=p3, n4=p4, n5=p5)
    return p3, p4, p5


def make_anchors(shapes, strides=STRIDES, offset=0.5):      # auto_speed_head.py:8-21
    a, s = [], []
    for (h, w), st in zip(shapes, strides):
        sx = torch.arange(w, dtype=torch.float32) + offset
        sy = torch.arange(h, dtype=torch.float32) + offset
        yy, xx = torch.meshgrid(sy, sx, indexing="ij")
        a.append(torch.stack((xx, yy), -1).view(-1, 2))
        s.append(torch.full((h * w, 1), st))
    return torch.cat(a).t(), torch.cat(s).t()


def head_raw(sd, feats) -> List[torch.Tensor]:
    """Per level [1, 64 + nc, h, w]: box branch (DFL logits) | class logits (auto_speed_head.py:47-49)."""
    outs = []
    for i, x in enumerate(feats):
        b = conv(sd, f"head.box.{i}.1", conv(sd, f"head.box.{i}.0", x, 3, 1, 1), 3, 1, 1)
        b = plain(sd, f"head.box.{i}.2", b)
        c = conv(sd, f"head.cls.{i}.0", x, 3, 1, 1, g=x.shape[1])
        c = conv(sd, f"head.cls.{i}.1", c)
        c = conv(sd, f"head.cls.{i}.2", c, 3, 1, 1, g=c.shape[1])
        c = conv(sd, f"head.cls.{i}.3", c)
        c = plain(sd, f"head.cls.{i}.4", c, cal_bias=False)
        outs.append(torch.cat((b, c), 1))
    return outs


def head_decode(levels: List[torch.Tensor]) -> torch.Tensor:
    """auto_speed_head.py:53-63: DFL expectation, anchors -/+ distances, (cx, cy, w, h) * stride, sigmoid(cls)."""
    anchors, strides = make_anchors([t.shape[-2:] for t in levels])
    x = torch.cat([t.view(t.shape[0], 4 * DFL_CH + NC, -1) for t in levels], 2)
    box, cls = x.split((4 * DFL_CH, NC), 1)
    b, _, a = box.shape
    d = box.view(b, 4, DFL_CH, a).transpose(2, 1).softmax(1)
    d = (d * torch.arange(DFL_CH, dtype=torch.float32).view(1, DFL_CH, 1, 1)).sum(1)      # DFL.conv, weights 0..15
    lt, rb = d.chunk(2, 1)
    lt = anchors.unsqueeze(0) - lt
    rb = anchors.unsqueeze(0) + rb
    box = torch.cat(((lt + rb) / 2, rb - lt), 1)
    return torch.cat((box * strides, cls.sigmoid()), 1)


def forward(sd: SD, x: torch.Tensor, taps=None) -> torch.Tensor:
    """x [1,3,512,1024] fp32 in [0,1] -> predictions [1, 4 + nc, 10752]."""
    with torch.no_grad():
        levels = head_raw(sd, neck(sd, backbone(sd, x, taps), taps))
        if taps is not None:
            for i, t in enumerate(levels):
                taps[f"head{i}"] = t
        return head_decode(levels)


# ------------------------------------------------------------------ helper: pre / post (auto_speed_infer.py)
def letterbox_geometry(orig_w: int, orig_h: int) -> Tuple[float, int, int, int, int]:
    scale = min(IMG_W / orig_w, IMG_H / orig_h)
    new_w, new_h = int(orig_w * scale), int(orig_h * scale)
    return scale, new_w, new_h, (IMG_W - new_w) // 2, (IMG_H - new_h) // 2


def letterbox(frame_rgb: np.ndarray):
    """auto_speed_infer.py:24-45 with Pillow itself (the resize is third-party code: Image.BILINEAR with antialias)."""
    from PIL import Image
    img = Image.fromarray(frame_rgb)
    scale, new_w, new_h, pad_x, pad_y = letterbox_geometry(*img.size)
    padded = Image.new("RGB", (IMG_W, IMG_H), (114, 114, 114))
    padded.paste(img.resize((new_w, new_h), Image.BILINEAR), (pad_x, pad_y))
    return np.asarray(padded), scale, pad_x, pad_y


def to_tensor(img_u8: np.ndarray) -> torch.Tensor:
    """transforms.ToTensor (auto_speed_infer.py:50): HWC uint8 -> 1x3xHxW fp32 / 255 (the reference then casts to
    half for its fp16 checkpoint; the oracle stays fp32)."""
    return torch.from_numpy(np.ascontiguousarray(img_u8)).permute(2, 0, 1).contiguous().to(torch.float32).div(255).unsqueeze(0)


def nms(boxes: np.ndarray, scores: np.ndarray, iou_thres: float) -> np.ndarray:
    """torchvision.ops.nms (third-party): greedy, descending score (stable for ties), suppress IoU > threshold."""
    order = np.argsort(-scores, kind="stable")
    x1, y1, x2, y2 = boxes.T
    area = (x2 - x1) * (y2 - y1)
    keep, dead = [], np.zeros(len(boxes), bool)
    for idx in order:
        if dead[idx]:
            continue
        keep.append(idx)
        w = np.maximum(0.0, np.minimum(x2[idx], x2) - np.maximum(x1[idx], x1))
        h = np.maximum(0.0, np.minimum(y2[idx], y2) - np.maximum(y1[idx], y1))
        inter = w * h
        iou = inter / (area[idx] + area - inter)
        dead |= iou > iou_thres
    return np.asarray(keep, dtype=np.int64)


def post_process(raw: torch.Tensor, conf_thres=0.6, iou_thres=0.45) -> np.ndarray:
    """auto_speed_infer.py:71-91: NOTE the second sigmoid on the already-sigmoided class scores (:78)."""
    pred = raw[0].t().numpy().astype(np.float32)
    boxes, probs = pred[:, :4], pred[:, 4:]
    sg = (1.0 / (1.0 + np.exp(-probs.astype(np.float32)))).astype(np.float32)
    scores, cls = sg.max(1), sg.argmax(1)
    m = scores > conf_thres
    if not m.any():
        return np.zeros((0, 6), np.float32)
    b = boxes[m]
    xyxy = np.stack((b[:, 0] - b[:, 2] / 2, b[:, 1] - b[:, 3] / 2, b[:, 0] + b[:, 2] / 2, b[:, 1] + b[:, 3] / 2), 1)
    comb = np.concatenate((xyxy, scores[m][:, None], cls[m][:, None].astype(np.float32)), 1).astype(np.float32)
    return comb[nms(comb[:, :4], comb[:, 4], iou_thres)]


def unletterbox(pred: np.ndarray, scale: float, pad_x: int, pad_y: int, orig_w: int, orig_h: int) -> np.ndarray:
    """auto_speed_infer.py:100-106."""
    p = pred.copy()
    p[:, [0, 2]] = np.clip((p[:, [0, 2]] - pad_x) / scale, 0, orig_w)
    p[:, [1, 3]] = np.clip((p[:, [1, 3]] - pad_y) / scale, 0, orig_h)
    return p


def inference(sd: SD, frame_rgb: np.ndarray) -> np.ndarray:
    """AutoSpeedNetworkInfer.inference (auto_speed_infer.py:88-108) -> [[x1,y1,x2,y2,score,class], ...]."""
    img, scale, pad_x, pad_y = letterbox(frame_rgb)
    pred = post_process(forward(sd, to_tensor(img)))
    return unletterbox(pred, scale, pad_x, pad_y, frame_rgb.shape[1], frame_rgb.shape[0]) if len(pred) else pred


# ------------------------------------------------------------------ synthetic checkpoint
def state_dict_spec() -> List[Tuple[str, tuple, str]]:
    """(name, shape, kind) in the reference module's state_dict order; kind in w | b | bn_w | bn_b | bn_m | bn_v | bn_n | dfl."""
    out: List[Tuple[str, tuple, str]] = []

    def cv(p, cin, cout, k=1, g=1):
        out.append((p + ".conv.weight", (cout, cin // g, k, k), "w"))
        for n, kind in (("weight", "bn_w"), ("bias", "bn_b"), ("running_mean", "bn_m"), ("running_var", "bn_v")):
            out.append((f"{p}.norm.{n}", (cout,), kind))
        out.append((p + ".norm.num_batches_tracked", (), "bn_n"))

    def wb(p, shape):
        out.append((p + ".weight", shape, "w"))
        out.append((p + ".bias", (shape[0],), "b"))

    def ctxm(p, cin, cout, h, w):
        wb(p + ".exp0", (h * w, cin, 3))
        wb(p + ".ctx0", (cin // 2, 1, 3, 3))
        wb(p + ".ctx1", (cin, cin // 2, 3, 3))
        wb(p + ".ctx2", (cout, cin, 3, 3))

    def res(p, ch, e):
        cv(p + ".conv1", ch, int(ch * e), 3)
        cv(p + ".conv2", int(ch * e), ch, 3)

    def c3k2m(p, cin, cout, csp):
        c = cout // 2
        cv(p + ".conv1", cin, 2 * c)
        cv(p + ".conv2", 3 * c, cout)
        if not csp:
            res(p + ".res_m.0", c, 0.5)
        else:
            q = p + ".res_m.0"
            cv(q + ".conv1", c, c // 2)
            cv(q + ".conv2", c, c // 2)
            cv(q + ".conv3", 2 * (c // 2), c)
            res(q + ".res_m.0", c // 2, 1.0)
            res(q + ".res_m.1", c // 2, 1.0)

    w = WIDTH
    cv("net.p1", w[0], w[1], 3)
    cv("net.p2.0", w[1], w[2], 3); ctxm("net.p2.1", w[2], w[3], 128, 256)
    cv("net.p3.0", w[3], w[3], 3); ctxm("net.p3.1", w[3], w[4], 64, 128)
    cv("net.p4.0", w[4], w[4], 3); ctxm("net.p4.1", w[4], w[4], 32, 64)
    cv("net.p5.0", w[4], w[5], 3); ctxm("net.p5.1", w[5], w[5], 16, 32)
    cv("net.p5.2.cv1", w[5], w[5] // 2); cv("net.p5.2.cv2", w[5] * 2, w[5])
    c_ = w[5] // 2
    cv("net.p5.3.cv1", w[5], 2 * c_); cv("net.p5.3.cv2", 2 * c_, w[5])
    a = "net.p5.3.middle_block"
    nh = c_ // 64
    cv(a + ".conv1.qkv", c_, c_ + (c_ // nh // 2) * nh * 2)
    cv(a + ".conv1.conv1", c_, c_, 3, g=c_)
    cv(a + ".conv1.conv2", c_, c_)
    cv(a + ".conv2.0", c_, 2 * c_); cv(a + ".conv2.1", 2 * c_, c_)
    c3k2m("fpn.h1", w[4] + w[5], w[4], False)
    c3k2m("fpn.h2", w[4] + w[4], w[3], False)
    cv("fpn.h3", w[3], w[3], 3)
    c3k2m("fpn.h4", w[3] + w[4], w[4], False)
    cv("fpn.h5", w[4], w[4], 3)
    c3k2m("fpn.h6", w[4] + w[5], w[5], True)
    out.append(("head.dfl.conv.weight", (1, DFL_CH, 1, 1), "dfl"))
    filt = (w[3], w[4], w[5])
    box_c, cls_c = max(64, filt[0] // 4), max(80, filt[0], NC)
    for i, f in enumerate(filt):
        cv(f"head.box.{i}.0", f, box_c, 3); cv(f"head.box.{i}.1", box_c, box_c, 3)
        wb(f"head.box.{i}.2", (4 * DFL_CH, box_c, 1, 1))
    for i, f in enumerate(filt):
        cv(f"head.cls.{i}.0", f, f, 3, g=f); cv(f"head.cls.{i}.1", f, cls_c)
        cv(f"head.cls.{i}.2", cls_c, cls_c, 3, g=cls_c); cv(f"head.cls.{i}.3", cls_c, cls_c)
        wb(f"head.cls.{i}.4", (NC, cls_c, 1, 1))
    return out


CALIB_PATH = __import__("os").path.join(__import__("os").path.dirname(__import__("os").path.abspath(__file__)), "..",
                                         "tests", "golden", "autospeed_calib.json")


def calibrate(seed: int = 1240) -> Dict[str, float]:
    """LSUV-style pass in execution order on the letterboxed synthetic frame 0: every layer's affine parameters are
    rescaled so that its pre-activation output has unit standard deviation (class logits: std 1 around a -3 bias, so
    that ~2 % of the 10 752 anchors pass the helper's 0.6 confidence filter and the NMS has real work).  The factors
    are committed (tests/golden/autospeed_calib.json); synth_state_dict applies them, so every machine builds
    bit-identical weights."""
    global _CAL
    from oracle import synth
    sd = draw_state_dict(seed)
    factors: Dict[str, float] = {}

    def hook(name, y):
        f = float(1.0 / max(float(y.std()), 1e-12))
        factors[name] = f
        return f

    _CAL = hook
    try:
        forward(sd, to_tensor(letterbox(synth.synth_frame(0))[0]))
    finally:
        _CAL = None
    return factors


def synth_state_dict(seed: int = 1240, calib: Dict[str, float] | None = None) -> SD:
    """The seeded draw with the committed calibration factors applied."""
    import json
    if calib is None:
        with open(CALIB_PATH) as f:
            calib = json.load(f)
    sd = draw_state_dict(seed)
    for name, f in calib.items():
        if name + ".norm.weight" in sd:
            keys = (name + ".norm.weight", name + ".norm.bias")
        elif name.startswith("head.cls"):
            keys = (name + ".weight",)                    # class logits: std 1 around the drawn bias (-3)
        else:
            keys = (name + ".weight", name + ".bias")
        for k in keys:
            sd[k] = sd[k] * np.float32(f)
    return sd


def draw_state_dict(seed: int = 1240) -> SD:
    """Seeded synthetic AutoSpeed checkpoint before calibration (the real weights are not reachable offline):
    weights ~ N(0, 1/fan_in), BatchNorm statistics != (0, 1) so that folding is exercised."""
    sd: SD = {}
    for idx, (name, shape, kind) in enumerate(state_dict_spec()):
        rng = np.random.default_rng([seed, idx])
        if kind == "w":
            fan_in = int(np.prod(shape[1:]))
            if name.endswith("exp0.weight"):
                fan_in = shape[1]                                    # only the centre tap of the Conv1d sees data
            t = (rng.standard_normal(shape) / math.sqrt(fan_in)).astype(np.float32)
        elif kind == "b":
            if name.startswith("head.cls") and name.endswith(".4.bias"):
                t = (-3.0 + 0.2 * rng.standard_normal(shape)).astype(np.float32)
            elif name.startswith("head.box") and name.endswith(".2.bias"):
                t = (1.0 + 0.5 * rng.standard_normal(shape)).astype(np.float32)
            else:
                t = (0.1 * rng.standard_normal(shape)).astype(np.float32)
        elif kind in ("bn_b", "bn_m"):
            t = (0.1 * rng.standard_normal(shape)).astype(np.float32)
        elif kind in ("bn_w", "bn_v"):
            t = rng.uniform(0.7, 1.3, shape).astype(np.float32)
        elif kind == "bn_n":
            t = np.zeros((), dtype=np.int64)
        else:                                                        # DFL projection 0..15 (common_layers.py:148-149)
            t = np.arange(DFL_CH, dtype=np.float32).reshape(shape)
        sd[name] = torch.from_numpy(np.asarray(t))
    return sd
