"""Wider parity set: all four networks on
  * 8 stills of the reference's own test video (VisionPilot/software_defined_vehicle/OpenADKit/Test/traffic-driving.mp4,
    committed under tests/golden/real/ by scripts/make_real_golden.py together with the outputs of the UNMODIFIED
    reference modules on them),
  * 8 synthetic 1080p frames + the i.i.d. adversarial frame through the fused Pillow-bicubic pre-process,
each against the fp32 CPU oracle, and the integer maps against the committed reference goldens under the margin rule.

Gates: synthetic frames — the 16-bit-operand gates of tests/test_engine_gpu.py (0.075 / 0.005 sigma).  Real stills —
EgoLanes in the 16-bit mode (measured 0.008 sigma); SceneSeg / Scene3D / DomainSeg in the split-fp16 fp32-grade mode at
0.02 / 0.002 sigma (measured 0.006): with the SYNTHETIC checkpoints (no trained weights are reachable offline) real
frames drive those three random networks into an amplifying regime (f3 -> f4 gain 3.4x, context |max| 2 273, output
sigma 44 instead of 1) in which ANY 16-bit inference — a CPU emulation of plain fp16 storage included
(profiles/r2_se_gate_precision.md) — is ~1 sigma away from fp32; that is a property of the random weights, not of
the engine, and the fp32-grade mode shows the arithmetic is right on exactly these frames."""
import os

import numpy as np
import pytest
import torch
from PIL import Image

from autoware_vision_pilot_b200 import engine as E
from autoware_vision_pilot_b200 import weights as W
from oracle import net, resize, synth

pytestmark = pytest.mark.gpu
REAL = os.path.join(synth.GOLDEN_DIR, "real")
REAL_FRAMES = (0, 4, 8, 12, 16, 20, 24, 28)
GMAX, GMEAN = 0.075, 0.005          # fp16 operands


@pytest.fixture(scope="module")
def ckpt(tmp_path_factory):
    d = tmp_path_factory.mktemp("ckpt_real")
    return {m: (synth.synth_state_dict(m), None) for m in net.MODELS}, d


def _vpw(ckpt, m):
    sds, d = ckpt
    p = str(d / f"{m}.vpw")
    if not os.path.exists(p):
        W.write_vpw(sds[m][0], p)
    return sds[m][0], p


def _check(model, ref, raw, cls, tag="", gmax=GMAX, gmean=GMEAN):
    sig = ref.std()
    err = np.abs(raw - ref)
    print(f"  {model} {tag}: sigma {sig:.3f}  max|d| {err.max() / sig:.4f} sigma  mean|d| {err.mean() / sig:.5f} sigma")
    assert np.isfinite(raw).all()
    assert err.max() <= gmax * sig and err.mean() <= gmean * sig, (model, tag, err.max() / sig, err.mean() / sig)
    tau = 2 * err.max()
    if model == "scene_seg":
        srt = np.sort(ref, axis=0)
        bad = cls != ref.argmax(0)
        assert not (bad & ((srt[-1] - srt[-2]) > tau)).any() and bad.mean() < 5e-3
    elif model == "domain_seg":
        bad = cls != (ref[0] > 0)
        assert not (bad & (np.abs(ref[0]) > tau)).any() and bad.mean() < 5e-3
    elif model == "ego_lanes":
        bad = cls != net.ego_lanes_masks(ref)[1]
        assert not (bad & (np.abs(ref).min(axis=0) > tau)).any()
    return err.max() / sig


@pytest.mark.parametrize("model", net.MODELS)
def test_real_video_stills(model, ckpt):
    sd, vpw = _vpw(ckpt, model)
    gold = np.load(os.path.join(REAL, f"{model}_real.npz"))
    precise = model != "ego_lanes"
    eng = E.Engine([E.KIND_BY_NAME[model]], [vpw], resize_mode=E.RESIZE_NONE, dtype="fp32" if precise else "fp16")
    gates = (0.02, 0.002) if precise else (GMAX, GMEAN)
    worst = 0.0
    for i in REAL_FRAMES:
        small = np.asarray(Image.open(os.path.join(REAL, f"frame_{i:02d}.png")).convert("RGB"))
        assert small.shape == (320, 640, 3)
        eng.infer(small)
        ref = net.forward(model, sd, net.to_tensor_normalize(small))[0].numpy()
        # the oracle on this still == what the unmodified reference module produced in the build container
        # (fp32 reassociation differs between the two hosts' CPU kernels: 5e-4 sigma)
        assert np.abs(ref[:, ::8, ::8] - gold[f"sample_{i}"]).max() <= 5e-4 * max(1.0, float(gold[f"std_{i}"]))
        raw, cls = eng.raw(0), eng.cls(0)
        worst = max(worst, _check(model, ref, raw, cls, f"real frame {i}", *gates))
        if model != "scene_3d":       # integer map vs the REFERENCE's own (margin from the oracle logits)
            gpost = gold[f"post_{i}"]
            margin = {"scene_seg": lambda r: np.sort(r, axis=0)[-1] - np.sort(r, axis=0)[-2],
                      "domain_seg": lambda r: np.abs(r[0]), "ego_lanes": lambda r: np.abs(r).min(axis=0)}[model](ref)
            bad = cls != gpost
            assert not (bad & (margin > 2 * np.abs(raw - ref).max() + 1e-3)).any()
    print(f"{model}: 8 real stills, worst max|d| = {worst:.4f} sigma")


@pytest.mark.parametrize("model", net.MODELS)
def test_eight_synthetic_frames_and_iid(model, ckpt):
    sd, vpw = _vpw(ckpt, model)
    eng = E.Engine([E.KIND_BY_NAME[model]], [vpw], resize_mode=E.RESIZE_PIL_BICUBIC)
    rng = np.random.default_rng(77)
    frames = [synth.synth_frame(synth.stream_seed(k % 4, 10 + k)) for k in range(8)]
    frames.append(rng.integers(0, 256, (1080, 1920, 3), dtype=np.uint8))          # the adversarial resize case
    for k, f in enumerate(frames):
        small = resize.pil_bicubic_resize(f, 640, 320)
        eng.infer(f)
        assert np.array_equal(eng.read_resized(), small)
        ref = net.forward(model, sd, net.to_tensor_normalize(small))[0].numpy()
        _check(model, ref, eng.raw(0), eng.cls(0), f"synthetic {k}" if k < 8 else "iid")


def test_fused_resize_on_the_real_1080p_frame(ckpt):
    sd, vpw = _vpw(ckpt, "scene_seg")
    full = np.asarray(Image.open(os.path.join(REAL, "frame_12_1080p.png")).convert("RGB"))
    assert full.shape == (1080, 1920, 3)
    small = np.asarray(Image.open(os.path.join(REAL, "frame_12.png")).convert("RGB"))
    assert np.array_equal(np.asarray(Image.fromarray(full).resize((640, 320))), small)   # Pillow here == build container
    eng = E.Engine([E.SCENE_SEG], [vpw], resize_mode=E.RESIZE_PIL_BICUBIC)
    eng.infer(full)
    assert np.array_equal(eng.read_resized(), small)                                       # fused kernel == Pillow
    a = eng.cls(0).copy()
    eng2 = E.Engine([E.SCENE_SEG], [vpw], resize_mode=E.RESIZE_NONE)
    eng2.infer(small)
    assert np.array_equal(eng2.cls(0), a)            # resize inside or outside the engine: same class map
