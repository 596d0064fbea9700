"""The numpy integer restatements in oracle/resize.py must be bit-exact against the libraries the
reference's callers use (Pillow for Models/visualizations, OpenCV for the C++ backends)."""
import numpy as np
import pytest

from oracle import resize, synth


@pytest.mark.parametrize("kind", ["natural", "iid"])
def test_pil_bicubic_1080p(kind):
    from PIL import Image
    f = synth.synth_frame(7, kind=kind)
    got = resize.pil_bicubic_resize(f, 640, 320)
    exp = np.asarray(Image.fromarray(f).resize((640, 320)))   # default filter == BICUBIC
    assert np.array_equal(got, exp)


@pytest.mark.parametrize("kind", ["natural", "iid"])
def test_cv_linear_1080p_and_egolanes_crop(kind):
    import cv2
    f = synth.synth_frame(9, kind=kind)
    assert np.array_equal(resize.cv_linear_resize(f, 640, 320), cv2.resize(f, (640, 320)))
    crop = np.ascontiguousarray(f[420:])                       # main.cpp:497-502
    assert np.array_equal(resize.cv_linear_resize(crop, 640, 320),
                          cv2.resize(crop, (640, 320), interpolation=cv2.INTER_LINEAR))


@pytest.mark.parametrize("h,w", [(700, 401), (333, 517), (320, 640), (480, 640), (2160, 3840)])
def test_ragged_sizes(h, w):
    import cv2
    from PIL import Image
    rng = np.random.default_rng(h * 7 + w)
    f = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    assert np.array_equal(resize.pil_bicubic_resize(f, 640, 320),
                          np.asarray(Image.fromarray(f).resize((640, 320))))
    assert np.array_equal(resize.cv_linear_resize(f, 640, 320), cv2.resize(f, (640, 320)))


def test_identity_size_is_passthrough():
    f = synth.synth_frame(1, 320, 640)
    assert np.array_equal(resize.pil_bicubic_resize(f, 640, 320), f)


def test_frames_are_deterministic():
    a, b = synth.synth_frame(5, 90, 160), synth.synth_frame(5, 90, 160)
    assert np.array_equal(a, b) and a.dtype == np.uint8 and a.shape == (90, 160, 3)
    assert not np.array_equal(a, synth.synth_frame(6, 90, 160))
    assert synth.stream_seed(3, 17) == 3017
