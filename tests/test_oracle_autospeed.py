"""CPU: the AutoSpeed oracle (oracle/autospeed.py) against (a) the committed goldens produced by the UNMODIFIED
reference module + helper (scripts/make_autospeed_golden.py) — runs anywhere; (b) the reference itself when
/root/reference is present: strict state_dict load, raw predictions, the helper's post-process."""
import hashlib
import os
import sys

import numpy as np
import pytest
import torch

from oracle import autospeed as A
from oracle import ref_import, synth


@pytest.fixture(scope="module")
def sd():
    return A.synth_state_dict()


@pytest.mark.parametrize("fi", [0, 1])
def test_oracle_matches_committed_reference_goldens(sd, fi):
    g = np.load(os.path.join(synth.GOLDEN_DIR, f"autospeed_f{fi}.npz"))
    frame = synth.synth_frame(fi)
    img, scale, pad_x, pad_y = A.letterbox(frame)
    assert np.array_equal(np.frombuffer(hashlib.sha256(img.tobytes()).digest(), dtype=np.uint8), g["img_sha"])
    assert (scale, pad_x, pad_y) == tuple(g["geom"])
    taps = {}
    raw = A.forward(sd, A.to_tensor(img), taps)
    assert np.abs(raw[0, :, ::7].numpy() - g["raw_sample"]).max() <= 2e-3          # box coordinates reach ~1e3
    for k, v in taps.items():
        assert np.allclose([v.mean().item(), v.std().item()], g["stat_" + k][:2], rtol=1e-3, atol=1e-4), k
    det = A.inference(sd, frame)
    assert det.shape == g["detections"].shape and len(det) >= 10
    assert np.abs(det - g["detections"]).max() <= 5e-3


def test_nms_and_postprocess_edge_cases():
    assert A.post_process(torch.zeros(1, 8, 10)).shape == (0, 6)                    # nothing above the confidence filter
    boxes = np.array([[0, 0, 10, 10], [1, 1, 11, 11], [20, 20, 30, 30], [0, 0, 10, 10]], np.float32)
    keep = A.nms(boxes, np.array([0.9, 0.8, 0.7, 0.9], np.float32), 0.45)
    assert keep.tolist() == [0, 2]                                                  # duplicate + overlap suppressed, stable ties
    s, nw, nh, px, py = A.letterbox_geometry(1920, 1080)
    assert (nw, nh, px, py) == (910, 512, 57, 0)
    s, nw, nh, px, py = A.letterbox_geometry(640, 640)
    assert (nw, nh, px, py) == (512, 512, 256, 0)


@pytest.mark.skipif(not ref_import.available(), reason="needs /root/reference (build container)")
def test_oracle_equals_the_unmodified_reference(sd):
    sys.path.insert(0, ref_import.REFERENCE_ROOT)
    sys.path.insert(0, ref_import.MODELS_DIR)
    from Models.model_components.auto_speed.auto_speed_network import AutoSpeedNetwork
    from inference.auto_speed_infer import AutoSpeedNetworkInfer
    m = AutoSpeedNetwork().build_model("n", 4).eval()
    assert [(k, tuple(v.shape)) for k, v in m.state_dict().items()] == [(k, tuple(v.shape)) for k, v in sd.items()]
    m.load_state_dict(sd, strict=True)
    helper = AutoSpeedNetworkInfer.__new__(AutoSpeedNetworkInfer)
    helper.train_size = (A.IMG_W, A.IMG_H)
    from PIL import Image
    frame = synth.synth_frame(2)
    img_ref, scale, pad_x, pad_y = helper.resize_letterbox(Image.fromarray(frame))
    img, s2, px2, py2 = A.letterbox(frame)
    assert np.array_equal(np.asarray(img_ref), img) and (scale, pad_x, pad_y) == (s2, px2, py2)
    with torch.no_grad():
        raw_ref = m(A.to_tensor(img))
    raw = A.forward(sd, A.to_tensor(img))
    assert (raw_ref - raw).abs().max().item() <= 1e-3
    det_ref = helper.post_process_predictions(raw_ref).numpy()
    det = A.post_process(raw)
    assert det.shape == det_ref.shape and np.abs(det - det_ref).max() <= 1e-3
