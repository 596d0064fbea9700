"""Pin oracle/post.py against the libraries the reference calls (cv2 here; Eigen QR == SVD to ~1e-12
for these 3-column systems)."""
import numpy as np
import pytest

from oracle import post


@pytest.mark.parametrize("dh,dw", [(1080, 1920), (720, 1280), (333, 517)])
def test_resize_nearest_matches_cv2(dh, dw):
    import cv2
    rng = np.random.default_rng(1)
    m = rng.integers(0, 2, (320, 640), dtype=np.uint8) * 255
    assert np.array_equal(post.resize_nearest(m, dw, dh), cv2.resize(m, (dw, dh), interpolation=cv2.INTER_NEAREST))


@pytest.mark.parametrize("dh,dw", [(1080, 1920), (720, 1280), (333, 517)])
def test_resize_linear_f32_matches_cv2(dh, dw):
    import cv2
    rng = np.random.default_rng(2)
    d = rng.standard_normal((320, 640)).astype(np.float32)
    got = post.resize_linear_f32(d, dw, dh)
    exp = cv2.resize(d, (dw, dh), interpolation=cv2.INTER_LINEAR)
    # opencv-python's IPP-backed float path evaluates the sample coordinates at a different precision
    # than resize.cpp's `fx = (float)((dx+0.5)*scale_x - 0.5)`; the difference is bounded by one fp32 ulp
    # of the coordinate (6e-5 at x~600) times the local gradient
    assert np.abs(got - exp).max() <= 3e-4 * max(1.0, np.abs(exp).max())


@pytest.mark.parametrize("order,n", [(1, 12), (2, 40), (2, 200), (3, 60)])
def test_polyfit_matches_cv_solve_svd(order, n):
    import cv2
    rng = np.random.default_rng(order * 100 + n)
    ys = rng.uniform(40, 79, n)
    xs = 0.01 * ys ** 2 - 0.7 * ys + 90 + rng.normal(0, 0.8, n)
    A = np.stack([ys ** (order - k) for k in range(order + 1)], axis=1)
    ok, c = cv2.solve(A, xs.reshape(-1, 1), flags=cv2.DECOMP_SVD)
    assert ok
    got = post.polyfit(xs, ys, order)
    assert np.abs(got - c[:, 0]).max() <= 1e-9 * np.abs(c).max()
    assert np.isnan(post.polyfit(xs[:order], ys[:order], order)).all()


def test_estimator_update_properties():
    st = post.initial_state()
    m = post.pathfinder_measurement([0.001, 0.02, -1.8], [0.001, 0.02, 1.9], 0.05, st[12, 0])
    s1 = post.estimator_update(st, m)
    # fused CTE is the inverse-variance mean of slots 0..2
    w = 1.0 / s1[0:3, 1]
    assert abs(s1[3, 0] - (w * s1[0:3, 0]).sum() / w.sum()) < 1e-12
    assert abs(s1[3, 1] - 1.0 / w.sum()) < 1e-12
    # NaN measurement only inflates the variance (estimator.cpp:33-37)
    assert s1[0, 0] == st[0, 0] and s1[0, 1] == st[0, 1] * 1.25
    # both lanes missing -> default width measurement (path_finder.cpp:141-143)
    nan3 = [float("nan")] * 3
    assert post.pathfinder_measurement(nan3, nan3, 0.0, 3.5)[12, 0] == 4.0


@pytest.mark.parametrize("viz", ["scene", "domain", "egolanes"])
def test_visualize_mask_matches_cv2_pipeline(viz):
    """createColorMask -> cv::resize(INTER_NEAREST) -> cv::addWeighted as the reference composes them
    (masks_visualization_engine.cpp:11-38), against the oracle's single-expression restatement."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(11)
    vals = {"scene": [0, 255], "domain": [0, 255, 7], "egolanes": [0, 1, 2, 255]}[viz]
    mask = rng.choice(vals, size=(320, 640)).astype(np.uint8)
    frame = rng.integers(0, 256, (1080, 1920, 3), dtype=np.uint8)
    cm = post.color_mask(mask, viz)
    ref_cm = np.zeros((320, 640, 3), np.uint8)
    if viz == "scene":
        ref_cm[cv2.inRange(mask, 1, 255) > 0] = (0, 0, 255)
    elif viz == "domain":
        ref_cm[mask == 0] = (255, 93, 61); ref_cm[mask == 255] = (145, 28, 255)
    else:
        ref_cm[mask == 0] = (255, 0, 0); ref_cm[mask == 1] = (255, 0, 200); ref_cm[mask == 2] = (0, 153, 0)
    assert np.array_equal(cm, ref_cm)
    ref = cv2.addWeighted(cv2.resize(ref_cm, (1920, 1080), interpolation=cv2.INTER_NEAREST), 0.5, frame, 0.5, 0.0)
    assert np.array_equal(post.visualize_mask(mask, frame, viz), ref)
    a = np.arange(256, dtype=np.uint8).reshape(-1, 1).repeat(256, 1)
    b = a.T.copy()
    assert np.array_equal(post.add_weighted_half(a, b), cv2.addWeighted(a, 0.5, b, 0.5, 0.0))
