"""Build-container script: AutoSpeed goldens from the UNMODIFIED reference module
(Models/model_components/auto_speed/auto_speed_network.py, Models/inference/auto_speed_infer.py) with the seeded synthetic
checkpoint of oracle/autospeed.py.  Writes tests/golden/autospeed_calib.json (with --calib) and
tests/golden/autospeed_f{0,1}.npz: letterbox image hash, strided raw predictions, per-tap statistics, the helper's
final detections.  Usage: python scripts/make_autospeed_golden.py [--calib]"""
import hashlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import autospeed as A  # noqa: E402
from oracle import ref_import, synth  # noqa: E402


def main():
    assert ref_import.available()
    if "--calib" in sys.argv:
        json.dump(A.calibrate(), open(A.CALIB_PATH, "w"), indent=0, sort_keys=True)
    sys.path.insert(0, ref_import.REFERENCE_ROOT)
    sys.path.insert(0, ref_import.MODELS_DIR)
    from Models.model_components.auto_speed.auto_speed_network import AutoSpeedNetwork
    from inference.auto_speed_infer import AutoSpeedNetworkInfer
    sd = A.synth_state_dict()
    m = AutoSpeedNetwork().build_model("n", 4).eval()
    m.load_state_dict(sd, strict=True)
    helper = AutoSpeedNetworkInfer.__new__(AutoSpeedNetworkInfer)
    helper.train_size = (A.IMG_W, A.IMG_H)
    from PIL import Image
    for fi in (0, 1):
        frame = synth.synth_frame(fi)
        img, scale, pad_x, pad_y = helper.resize_letterbox(Image.fromarray(frame))
        img = np.asarray(img)
        with torch.no_grad():
            raw = m(A.to_tensor(img))
        det = helper.post_process_predictions(raw).numpy()
        det = A.unletterbox(det, scale, pad_x, pad_y, frame.shape[1], frame.shape[0])
        taps = {}
        A.forward(sd, A.to_tensor(img), taps)
        rec = {"img_sha": np.frombuffer(hashlib.sha256(img.tobytes()).digest(), dtype=np.uint8),
               "raw_sample": raw[0, :, ::7].numpy().astype(np.float32), "detections": det.astype(np.float32),
               "geom": np.array([scale, pad_x, pad_y], dtype=np.float64)}
        for k, v in taps.items():
            rec["stat_" + k] = np.array([v.mean().item(), v.std().item(), v.abs().max().item()])
        path = os.path.join(synth.GOLDEN_DIR, f"autospeed_f{fi}.npz")
        np.savez_compressed(path, **rec)
        print("wrote", path, os.path.getsize(path) // 1024, "KiB,", len(det), "detections")


if __name__ == "__main__":
    main()
