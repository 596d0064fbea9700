"""Build-container script: calibrate the synthetic weights and generate tests/golden/*.

Needs /root/reference (imports the UNMODIFIED reference modules through oracle/ref_import.py).
Outputs (all small, committed):
  tests/golden/calib_scales.json      per-layer LSUV factors for oracle/synth.py
  tests/golden/<model>_f<frame>.npz   reference outputs on synthetic frames: class map / masks,
                                      stride-4 logits sample, per-tap statistics, logits checksum
Usage:  python scripts/make_golden.py [--calib] [--golden]
"""
import argparse
import hashlib
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import net, ref_import, resize, synth  # noqa: E402

GOLDEN_FRAMES = {"scene_seg": (0, 1), "scene_3d": (0,), "domain_seg": (0,), "ego_lanes": (0,)}


def calib_image():
    """Calibration input: synthetic 1080p frame 0 -> Pillow-bicubic 640x320 -> ToTensor/Normalize."""
    small = resize.pil_bicubic_resize(synth.synth_frame(0), 640, 320)
    return net.to_tensor_normalize(small)


def do_calib():
    img = calib_image()
    calib = {}
    for m in net.MODELS:  # scene_seg first: the others take its frozen parts
        t = time.time()
        calib[m] = synth.calibrate(m, img, calib)
        print(f"calibrated {m}: {len(calib[m])} layers in {time.time() - t:.1f}s", flush=True)
    os.makedirs(synth.GOLDEN_DIR, exist_ok=True)
    with open(synth.CALIB_PATH, "w") as f:
        json.dump(calib, f, indent=0, sort_keys=True)
    print("wrote", synth.CALIB_PATH)


def tap_stats(t):
    t = t.float()
    return np.array([t.mean().item(), t.std().item(), t.abs().max().item()], dtype=np.float64)


def do_golden():
    assert ref_import.available(), "needs /root/reference"
    torch.set_num_threads(os.cpu_count())
    for m in net.MODELS:
        sd = synth.synth_state_dict(m)
        ref = ref_import.build_network(m, sd)
        for fidx in GOLDEN_FRAMES[m]:
            frame = synth.synth_frame(fidx)
            small = resize.pil_bicubic_resize(frame, 640, 320)
            x = net.to_tensor_normalize(small)
            with torch.no_grad():
                out_ref = ref(x)
            taps = {}
            out_orc = net.forward(m, sd, x, taps=taps)
            diff = (out_ref - out_orc).abs().max().item()
            print(f"{m} frame {fidx}: |reference - oracle| max = {diff:.3e}, out std {out_ref.std().item():.3f}")
            o = out_ref[0].numpy()
            rec = {
                "small_sha": np.frombuffer(hashlib.sha256(small.tobytes()).digest(), dtype=np.uint8),
                "out_sample": o[:, ::4, ::4].astype(np.float32),
                "out_stats": tap_stats(out_ref),
                "out_sha": np.frombuffer(hashlib.sha256(o.astype(np.float32).tobytes()).digest(), dtype=np.uint8),
                "post": np.asarray(net.postprocess(m, out_ref)),
            }
            if m == "scene_seg":
                rec["post"] = rec["post"].astype(np.uint8)
                srt = np.sort(o, axis=0)
                rec["margin_f16"] = (srt[-1] - srt[-2]).astype(np.float16)
            elif m == "domain_seg":
                rec["post"] = rec["post"].astype(np.uint8)
            elif m == "scene_3d":
                rec["post"] = rec["post"][::4, ::4].astype(np.float32)
            for k, v in taps.items():
                rec["stat_" + k] = tap_stats(v)
            path = os.path.join(synth.GOLDEN_DIR, f"{m}_f{fidx}.npz")
            np.savez_compressed(path, **rec)
            print("  wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--calib", action="store_true")
    ap.add_argument("--golden", action="store_true")
    a = ap.parse_args()
    if a.calib:
        do_calib()
    if a.golden:
        do_golden()
