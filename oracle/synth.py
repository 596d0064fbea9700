"""ORACLE support (test infrastructure): deterministic synthetic frames and `state_dict`s.

The reference ships no weights (Google-Drive links, Models/model_library/SceneSeg/README.md:18-21)
and no fixtures, and plain default init collapses activations to ~1e-13, so the
parity harness generates its own:

* weights: every tensor is drawn from `numpy.random.default_rng([seed, tensor_index])` (portable,
  independent of torch's RNG and of module construction order), scaled 1/sqrt(fan_in) and then by
  a per-layer calibration factor that makes each layer's pre-activation output ~unit variance on a
  fixed calibration image (LSUV-style).  The factors are data-dependent, so they are computed once
  (scripts/make_golden.py) and committed in tests/golden/calib_scales.json — the generator itself
  is then pure and reproduces bit-identical tensors on any machine;
* BatchNorm running stats are drawn != (0,1) so that BN folding is exercised;
* Scene3D / DomainSeg re-use SceneSeg's frozen modules exactly as scene_3d_infer.py:28-29 and
  domain_seg_infer.py:28-29 construct them (shared encoder; DomainSeg also context + neck).

Frames follow SURVEY.md §8d: low-frequency noise (27x48 grid, bicubic-upsampled) * 0.8 + i.i.d.
uniform noise * 0.2, plus a pure i.i.d. frame as the adversarial resize case.
"""
from __future__ import annotations

import json
import os
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import net

SEEDS = {"scene_seg": 1234, "scene_3d": 1235, "domain_seg": 1236, "ego_lanes": 1237}
GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
CALIB_PATH = os.path.join(GOLDEN_DIR, "calib_scales.json")


# ---------------------------------------------------------------------------------------------
# state_dict layout — names, shapes and tensor kinds, in checkpoint order
# ---------------------------------------------------------------------------------------------
def _bn_spec(p: str, c: int):
    return [(p + "weight", (c,), "bn_w"), (p + "bias", (c,), "bn_b"),
            (p + "running_mean", (c,), "bn_m"), (p + "running_var", (c,), "bn_v"),
            (p + "num_batches_tracked", (), "bn_n")]


def encoder_spec(p: str) -> List[Tuple[str, tuple, str]]:
    s = [(p + "0.0.weight", (32, 3, 3, 3), "w")] + _bn_spec(p + "0.1.", 32)
    for si, (exp, k, stride, cin, cout, reps) in enumerate(net.MBCONV_STAGES, start=1):
        for r in range(reps):
            ci = cin if r == 0 else cout
            ce = ci * exp
            sq = max(1, ci // 4)
            bp = f"{p}{si}.{r}.block."
            i = 0
            if exp != 1:
                s += [(f"{bp}{i}.0.weight", (ce, ci, 1, 1), "w")] + _bn_spec(f"{bp}{i}.1.", ce)
                i += 1
            s += [(f"{bp}{i}.0.weight", (ce, 1, k, k), "w")] + _bn_spec(f"{bp}{i}.1.", ce)
            s += [(f"{bp}{i + 1}.fc1.weight", (sq, ce, 1, 1), "w"), (f"{bp}{i + 1}.fc1.bias", (sq,), "b"),
                  (f"{bp}{i + 1}.fc2.weight", (ce, sq, 1, 1), "w"), (f"{bp}{i + 1}.fc2.bias", (ce,), "b")]
            s += [(f"{bp}{i + 2}.0.weight", (cout, ce, 1, 1), "w")] + _bn_spec(f"{bp}{i + 2}.1.", cout)
    s += [(p + "8.0.weight", (1280, 320, 1, 1), "w")] + _bn_spec(p + "8.1.", 1280)
    return s


def _wb(name: str, shape: tuple):
    return [(name + ".weight", shape, "w"), (name + ".bias", (shape[1] if name.split(".")[-1].startswith("upsample") else shape[0],), "b")]


def context_spec(p: str, c: int):
    s = _wb(p + "context_layer_0", (800, c)) + _wb(p + "context_layer_1", (800, 800))
    s += _wb(p + "context_layer_2", (200, 800))
    s += _wb(p + "context_layer_3", (128, 1, 3, 3)) + _wb(p + "context_layer_4", (256, 128, 3, 3))
    s += _wb(p + "context_layer_5", (512, 256, 3, 3)) + _wb(p + "context_layer_6", (c, 512, 3, 3))
    return s


def neck_spec(p: str, c: int):
    s = _wb(p + "upsample_layer_0", (c, c, 2, 2)) + _wb(p + "skip_link_layer_0", (c, 80, 1, 1))
    s += _wb(p + "decode_layer_0", (768, c, 3, 3)) + _wb(p + "decode_layer_1", (768, 768, 3, 3))
    s += _wb(p + "upsample_layer_1", (768, 768, 2, 2)) + _wb(p + "skip_link_layer_1", (768, 40, 1, 1))
    s += _wb(p + "decode_layer_2", (512, 768, 3, 3)) + _wb(p + "decode_layer_3", (512, 512, 3, 3))
    s += _wb(p + "upsample_layer_2", (512, 512, 2, 2)) + _wb(p + "skip_link_layer_2", (512, 24, 1, 1))
    s += _wb(p + "decode_layer_4", (512, 512, 3, 3)) + _wb(p + "decode_layer_5", (256, 512, 3, 3))
    return s


def seg_head_spec(p: str, c9: int, cout: int):
    s = _wb(p + "upsample_layer_3", (256, 256, 2, 2)) + _wb(p + "skip_link_layer_3", (256, 32, 1, 1))
    s += _wb(p + "decode_layer_6", (256, 256, 3, 3)) + _wb(p + "decode_layer_7", (128, 256, 3, 3))
    s += _wb(p + "upsample_layer_4", (128, 128, 2, 2))
    s += _wb(p + "decode_layer_8", (128, 128, 3, 3)) + _wb(p + "decode_layer_9", (c9, 128, 3, 3))
    s += _wb(p + "decode_layer_10", (cout, c9, 3, 3))
    return s


def ego_head_spec(p: str):
    return (_wb(p + "decode_layer_6", (256, 256, 3, 3)) + _wb(p + "decode_layer_7", (128, 256, 3, 3))
            + _wb(p + "decode_layer_8", (3, 128, 3, 3)))


def state_dict_spec(model: str) -> List[Tuple[str, tuple, str]]:
    pf = net.PREFIX[model]
    c = 1456 if model == "ego_lanes" else 1280
    s = encoder_spec(pf["enc"]) + context_spec(pf["ctx"], c) + neck_spec(pf["neck"], c)
    if model == "scene_seg":
        s += seg_head_spec(pf["head"], 64, 3)
    elif model == "scene_3d":
        s += seg_head_spec(pf["head"], 128, 1)
    elif model == "domain_seg":
        s += seg_head_spec(pf["head"], 64, 1)
    else:
        s += ego_head_spec(pf["head"])
    return s


# ---------------------------------------------------------------------------------------------
# weights
# ---------------------------------------------------------------------------------------------
def _draw(seed: int, idx: int, shape: tuple, kind: str) -> np.ndarray:
    rng = np.random.default_rng([seed, idx])
    if kind == "w":
        fan_in = int(np.prod(shape[1:])) if len(shape) > 1 else 1
        if len(shape) == 4 and shape[2] == 2:          # ConvTranspose2d [Cin, Cout, 2, 2]: fan_in = Cin
            fan_in = shape[0]
        return (rng.standard_normal(shape) / np.sqrt(fan_in)).astype(np.float32)
    if kind in ("b", "bn_b", "bn_m"):
        return (0.1 * rng.standard_normal(shape)).astype(np.float32)
    if kind in ("bn_w", "bn_v"):
        return rng.uniform(0.5, 1.5, shape).astype(np.float32)
    if kind == "bn_n":
        return np.zeros((), dtype=np.int64)
    raise ValueError(kind)


def load_calib() -> Dict[str, Dict[str, float]]:
    with open(CALIB_PATH) as f:
        return json.load(f)


def _strip(model: str, name: str) -> Tuple[str, str]:
    """(part, local name) of a checkpoint key, e.g. ('enc', '0.0.weight')."""
    for part, p in sorted(net.PREFIX[model].items(), key=lambda kv: -len(kv[1])):
        if name.startswith(p):
            return part, name[len(p):]
    raise KeyError(name)


# which parts a downstream network takes frozen from SceneSeg
SHARED_PARTS = {"scene_seg": (), "scene_3d": ("enc",), "domain_seg": ("enc", "ctx", "neck"),
                "ego_lanes": ()}


def synth_state_dict(model: str, calib: Optional[Dict[str, Dict[str, float]]] = None,
                     share: bool = True) -> Dict[str, torch.Tensor]:
    """Seeded synthetic checkpoint for `model`.  With share=True (default) the parts listed in
    SHARED_PARTS are byte-identical to SceneSeg's (the reference's construction); share=False
    gives the deliberately non-identical pair used to test the no-sharing fallback."""
    if calib is None:
        calib = load_calib()
    sd: Dict[str, torch.Tensor] = {}
    ss_pf = net.PREFIX["scene_seg"]
    for idx, (name, shape, kind) in enumerate(state_dict_spec(model)):
        part, local = _strip(model, name)
        if share and part in SHARED_PARTS[model]:
            src_model, src_name = "scene_seg", ss_pf[part] + local
            src_idx = _SPEC_INDEX("scene_seg")[src_name]
            arr = _draw(SEEDS[src_model], src_idx, shape, kind)
            scale = calib.get(src_model, {}).get(src_name, 1.0) if kind == "w" else 1.0
        else:
            arr = _draw(SEEDS[model], idx, shape, kind)
            scale = calib.get(model, {}).get(name, 1.0) if kind == "w" else 1.0
        if kind == "w" and scale != 1.0:
            arr = (arr * np.float32(scale)).astype(np.float32)
        sd[name] = torch.from_numpy(np.ascontiguousarray(arr))
    return sd


_spec_index_cache: Dict[str, Dict[str, int]] = {}


def _SPEC_INDEX(model: str) -> Dict[str, int]:
    if model not in _spec_index_cache:
        _spec_index_cache[model] = {n: i for i, (n, _, _) in enumerate(state_dict_spec(model))}
    return _spec_index_cache[model]


def calibrate(model: str, image: torch.Tensor, calib_so_far: Dict[str, Dict[str, float]],
              target_std: float = 1.0) -> Dict[str, float]:
    """One forward pass in execution order; every conv / linear weight is rescaled so that its
    bias-free output has std `target_std` on `image`.  Shared (frozen) parts keep SceneSeg's
    factors.  Returns {checkpoint key: factor}."""
    import torch.nn.functional as F

    sd = synth_state_dict(model, calib_so_far)
    shared_pfx = tuple(net.PREFIX[model][p] for p in SHARED_PARTS[model])
    scales: Dict[str, float] = {}

    def hook(name: str, w: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
        if name.startswith(shared_pfx) and shared_pfx:
            return w
        if name in scales:
            return w
        if w.dim() == 2:
            y = F.linear(x, w)
        elif w.dim() == 4 and w.shape[2] == 2 and "upsample" in name:
            y = F.conv_transpose2d(x, w, stride=2)
        elif w.dim() == 4 and w.shape[1] == 1 and x.shape[1] == w.shape[0] and x.shape[1] > 1:
            y = F.conv2d(x, w, padding=(w.shape[-1] - 1) // 2, groups=w.shape[0])
        else:
            y = F.conv2d(x, w, padding=(w.shape[-1] - 1) // 2)
        s = float(target_std / max(float(y.std()), 1e-12))
        scales[name] = s
        w.mul_(s)                     # in place: the sd now holds the calibrated tensor
        return w

    net.forward(model, sd, image, hook=hook)
    return scales


# ---------------------------------------------------------------------------------------------
# frames
# ---------------------------------------------------------------------------------------------
def _cubic_matrix(n_out: int, n_in: int) -> np.ndarray:
    """Catmull-Rom (a=-0.5) interpolation matrix [n_out, n_in], align-centres, clamped borders."""
    m = np.zeros((n_out, n_in), dtype=np.float64)
    scale = n_in / n_out
    for o in range(n_out):
        c = (o + 0.5) * scale - 0.5
        i0 = int(np.floor(c))
        t = c - i0
        ws = [((-0.5 * t + 1.0) * t - 0.5) * t, ((1.5 * t - 2.5) * t) * t + 1.0,
              ((-1.5 * t + 2.0) * t + 0.5) * t, ((0.5 * t - 0.5) * t) * t]
        for k, wgt in enumerate(ws):
            m[o, min(max(i0 - 1 + k, 0), n_in - 1)] += wgt
    return m


def synth_frame(seed: int, h: int = 1080, w: int = 1920, kind: str = "natural") -> np.ndarray:
    """uint8 HWC RGB synthetic camera frame."""
    rng = np.random.default_rng(seed)
    if kind == "iid":
        return rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    gh, gw = 27, 48
    grid = rng.uniform(0.0, 255.0, (gh, gw, 3))
    my, mx = _cubic_matrix(h, gh), _cubic_matrix(w, gw)
    low = np.einsum("yg,gwc->ywc", my, np.einsum("xw,gwc->gxc", mx, grid))
    noise = rng.uniform(0.0, 255.0, (h, w, 3))
    img = 0.8 * low + 0.2 * noise
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def stream_seed(stream: int, frame: int) -> int:
    """Config 4: stream k uses seeds 1000*k + f."""
    return 1000 * stream + frame
