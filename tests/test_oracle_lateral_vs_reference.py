"""Pins oracle/lateral.py (and oracle/post.estimator_update) against the reference's OWN sources: the unmodified
lane_filter.cpp / lane_tracking.cpp / estimator.cpp compiled by oracle/build_ref.py into oracle/_ref (OpenCV's C++
API replaced by the minimal stand-in oracle/cvstub).  Runs where /root/reference (or a previously built
oracle/_ref/libref_lateral.so) is available; skipped otherwise."""
import ctypes as C

import numpy as np
import pytest

from oracle import build_ref
from oracle import lateral as LT
from oracle import post


class RefOut(C.Structure):
    _fields_ = [(n, C.c_double * 6) for n in ("left", "right", "center", "bev_left", "bev_right", "bev_center",
                                             "filt_left", "filt_right")] + \
               [(n, C.c_double) for n in ("lane_offset", "yaw_offset", "curvature", "bev_lane_offset", "bev_yaw_offset",
                                          "bev_curvature", "width_px")] + \
               [(n, C.c_int) for n in ("left_valid", "right_valid", "path_valid", "bev_valid", "filt_left_valid",
                                       "filt_right_valid", "n_left_windows", "n_right_windows", "n_bev_left", "n_bev_right")] + \
               [("bev_left_pts", C.c_float * 512), ("bev_right_pts", C.c_float * 512)]


@pytest.fixture(scope="module")
def ref():
    path = build_ref.build()
    if path is None:
        pytest.skip("reference sources not available and no prebuilt oracle/_ref")
    lib = C.CDLL(path)
    lib.ref_lateral_create.restype = C.c_void_p
    lib.ref_lateral_create.argtypes = [C.c_float]
    lib.ref_lateral_destroy.argtypes = [C.c_void_p]
    lib.ref_lateral_update.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(RefOut)]
    lib.ref_estimator_update.argtypes = [C.c_void_p, C.c_void_p]
    return lib


def _frames(seed0, n):
    rng = np.random.default_rng(seed0)
    out = []
    for k in range(n):
        kind = rng.integers(0, 10)
        if kind == 0:
            out.append(np.zeros((3, 80, 160), np.float32))
        elif kind == 1:
            out.append((rng.uniform(size=(3, 80, 160)) < rng.uniform(0.01, 0.2)).astype(np.float32))
        elif kind == 2:                                   # horizontal runs: rank-deficient fits
            m = np.zeros((3, 80, 160), np.float32)
            y = int(rng.integers(42, 79))
            m[0, y, 20:70] = 1
            m[1, y:y + 2, 90:150] = 1
            out.append(m)
        else:
            out.append(LT.synth_lane_masks(int(rng.integers(0, 1 << 30)), drop_left=rng.uniform() < 0.2,
                                           drop_right=rng.uniform() < 0.2, noise=float(rng.uniform(0, 0.03))))
    return out


@pytest.mark.parametrize("seed0", [1, 2, 3, 4, 5, 6])
def test_filter_and_tracker_match_the_compiled_reference(ref, seed0):
    """40-frame stateful sequences (temporal smoothing, width history, recovery branches, empty / noise-only /
    single-row frames): integer results identical, every coefficient and curve parameter within 1e-9.  Also
    confirms that the reference's RANSAC loop (unseeded mt19937) never changes its result: the restatement has
    no sampler and still agrees on every frame."""
    h = ref.ref_lateral_create(0.5)
    f, t = LT.LaneFilter(0.5), LT.LaneTracker()
    try:
        for k, m in enumerate(_frames(seed0, 40)):
            m = np.ascontiguousarray(m, dtype=np.float32)
            o = RefOut()
            ref.ref_lateral_update(h, m.ctypes.data, 80, 160, 1920, 1080, C.byref(o))
            fo = f.update(m)
            tr = t.update(fo.left, fo.right)
            tag = f"seed {seed0} frame {k}"
            assert bool(o.filt_left_valid) == (fo.left is not None), tag
            assert bool(o.filt_right_valid) == (fo.right is not None), tag
            if fo.left is not None:
                np.testing.assert_allclose(np.array(o.filt_left), fo.left, rtol=1e-9, atol=1e-9, err_msg=tag)
            if fo.right is not None:
                np.testing.assert_allclose(np.array(o.filt_right), fo.right, rtol=1e-9, atol=1e-9, err_msg=tag)
            assert bool(o.left_valid) == (tr.left is not None) and bool(o.right_valid) == (tr.right is not None), tag
            if tr.left is not None:
                np.testing.assert_allclose(np.array(o.left), tr.left, rtol=1e-9, atol=1e-9, err_msg=tag)
            if tr.right is not None:
                np.testing.assert_allclose(np.array(o.right), tr.right, rtol=1e-9, atol=1e-9, err_msg=tag)
            assert bool(o.path_valid) == tr.path_valid and bool(o.bev_valid) == tr.bev_valid, tag
            if tr.bev_valid:
                assert o.n_bev_left == len(tr.bev_left_pts) and o.n_bev_right == len(tr.bev_right_pts), tag
                np.testing.assert_array_equal(np.array(o.bev_left_pts[:2 * o.n_bev_left], np.float32).reshape(-1, 2),
                                              tr.bev_left_pts, err_msg=tag)
                np.testing.assert_array_equal(np.array(o.bev_right_pts[:2 * o.n_bev_right], np.float32).reshape(-1, 2),
                                              tr.bev_right_pts, err_msg=tag)
                np.testing.assert_allclose(np.array(o.center), tr.center, rtol=1e-9, atol=1e-9, err_msg=tag)
                for name, got in (("bev_left", tr.bev_left), ("bev_right", tr.bev_right), ("bev_center", tr.bev_center)):
                    np.testing.assert_allclose(np.array(getattr(o, name)), got, rtol=1e-7, atol=1e-7, err_msg=tag + name)
                for name, got in (("lane_offset", tr.lane_offset), ("yaw_offset", tr.yaw_offset), ("curvature", tr.curvature)):
                    np.testing.assert_allclose(getattr(o, name), got, rtol=1e-9, atol=1e-9, err_msg=tag + name)
                for name, got in (("bev_lane_offset", tr.bev_lane_offset), ("bev_yaw_offset", tr.bev_yaw_offset),
                                  ("bev_curvature", tr.bev_curvature)):
                    np.testing.assert_allclose(getattr(o, name), got, rtol=1e-7, atol=1e-7, err_msg=tag + name)
                np.testing.assert_allclose(o.width_px, t.width, rtol=1e-12, err_msg=tag)
    finally:
        ref.ref_lateral_destroy(h)


def test_sliding_window_count_matches_reference_debug_rects(ref):
    """LaneSegmentation.left/right_sliding_windows (one cv::Rect per visited window) vs the oracle's window list."""
    h = ref.ref_lateral_create(0.5)
    try:
        for seed in range(30, 50):
            m = LT.synth_lane_masks(seed, noise=0.02)
            o = RefOut()
            ref.ref_lateral_update(h, m.ctypes.data, 80, 160, 1920, 1080, C.byref(o))
            sl, sr = LT.find_starting_points(m)
            nl = len(LT.sliding_window_search(m, sl, True)[1]) if sl else 0
            nr = len(LT.sliding_window_search(m, sr, False)[1]) if sr else 0
            assert (o.n_left_windows, o.n_right_windows) == (nl, nr), seed
    finally:
        ref.ref_lateral_destroy(h)


def test_estimator_update_matches_reference(ref):
    rng = np.random.default_rng(9)
    for _ in range(50):
        st = np.stack([rng.normal(size=14), rng.uniform(0.01, 5.0, 14)], axis=1)
        ms = np.stack([rng.normal(size=14), rng.uniform(0.001, 1.0, 14)], axis=1)
        ms[rng.uniform(size=14) < 0.3, 0] = np.nan
        want = np.ascontiguousarray(st.copy())
        m = np.ascontiguousarray(ms)
        ref.ref_estimator_update(want.ctypes.data, m.ctypes.data)
        got = post.estimator_update(st, ms)
        np.testing.assert_allclose(got, want, rtol=1e-13, atol=0, equal_nan=True)


def test_fit_quad_poly_and_fitted_curve_match_reference(ref):
    """fitQuadPoly (Eigen colPivHouseholderQr, poly_fit.cpp:36-75) and FittedCurve (:26-34) vs oracle/post.py."""
    ref.ref_fit_quad.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    rng = np.random.default_rng(21)
    for n in (0, 2, 3, 5, 40, 130):
        ys = np.sort(rng.uniform(0, 40, n)).astype(np.float32)
        xs = (0.002 * ys ** 2 - 0.05 * ys + rng.normal(1.8, 0.05, n)).astype(np.float32)
        pts = np.ascontiguousarray(np.stack([xs, ys], 1), dtype=np.float32)
        coeff, cy = np.zeros(3), np.zeros(2)
        ref.ref_fit_quad(pts.ctypes.data, n, coeff.ctypes.data, cy.ctypes.data)
        got = post.polyfit(xs, ys, 2)
        if n <= 2:
            assert np.isnan(coeff).all() and np.isnan(got).all() and np.isnan(cy).all()
            continue
        np.testing.assert_allclose(got, coeff, rtol=1e-8, atol=1e-10)
        cte, yaw = post.fitted_curve(coeff)
        np.testing.assert_allclose([cte, yaw], cy, rtol=1e-13)


def test_pathfinder_update_matches_reference(ref):
    """PathFinder::update (path_finder.cpp:48-181) over a 30-frame sequence with missing lines.  The reference adds
    an unseeded U(-1e-5, 1e-5) to every state mean in its predict step, the restatement adds 0: means agree to
    1e-3, variances (deterministic) to 1e-12."""
    ref.ref_pathfinder_create.restype = C.c_void_p
    ref.ref_pathfinder_create.argtypes = [C.c_double]
    ref.ref_pathfinder_destroy.argtypes = [C.c_void_p]
    ref.ref_pathfinder_update.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_double, C.c_void_p,
                                          C.c_void_p]
    h = ref.ref_pathfinder_create(4.0)
    pf = LT.PathFinder(4.0)
    rng = np.random.default_rng(33)
    try:
        for k in range(30):
            ys = np.arange(250, 640, 13, dtype=np.float32)
            left = np.stack([288 + 0.02 * (640 - ys) + rng.normal(0, 0.5, len(ys)), ys], 1).astype(np.float32)
            right = np.stack([352 + 0.02 * (640 - ys) + rng.normal(0, 0.5, len(ys)), ys], 1).astype(np.float32)
            if k in (7, 8):
                left = left[:2]                    # <= 2 points: NaN fit, "left missing" branch
            if k == 15:
                right = right[:0]
            if k == 20:
                left, right = left[:1], right[:2]  # both missing -> default width
            steer = 0.01 * k
            lm, rm = LT.pixels_to_meters(left), LT.pixels_to_meters(right)
            lm, rm = np.ascontiguousarray(lm), np.ascontiguousarray(rm)
            out, state = np.zeros(9), np.zeros((14, 2))
            ref.ref_pathfinder_update(h, lm.ctypes.data, len(lm), rm.ctypes.data, len(rm), steer, out.ctypes.data,
                                      state.ctypes.data)
            got = pf.update(left, right, steer)
            np.testing.assert_allclose(pf.state[:, 1], state[:, 1], rtol=1e-12, err_msg=f"variances, frame {k}")
            np.testing.assert_allclose(pf.state[:, 0], state[:, 0], rtol=0, atol=1e-3, err_msg=f"means, frame {k}")
            np.testing.assert_allclose([got["cte"], got["yaw_error"], got["lane_width"]], out[[0, 1, 3]], atol=1e-3)
            np.testing.assert_allclose([got["cte_variance"], got["yaw_variance"], got["curv_variance"],
                                        got["lane_width_variance"]], out[4:8], rtol=1e-12)
            assert got["curvature"] == out[2] and bool(out[8]) == got["fused_valid"]
    finally:
        ref.ref_pathfinder_destroy(h)


def test_random_single_frames_match_reference(ref):
    """300 independent random frames (fresh filter / tracker each): sparse noise of varying density, random line
    segments, blobs — exercises start-point search, window clamping at the image borders and the fit order switch."""
    rng = np.random.default_rng(777)
    for k in range(300):
        m = np.zeros((3, 80, 160), np.float32)
        for _ in range(int(rng.integers(0, 6))):            # random straight segments in random channels
            ch = int(rng.integers(0, 3))
            x0, x1 = rng.uniform(0, 160, 2)
            y0, y1 = rng.uniform(20, 80, 2)
            n = int(max(abs(x1 - x0), abs(y1 - y0))) + 1
            xs = np.clip(np.round(np.linspace(x0, x1, n)).astype(int), 0, 159)
            ys = np.clip(np.round(np.linspace(y0, y1, n)).astype(int), 0, 79)
            m[ch, ys, xs] = 1.0
            if rng.uniform() < 0.5:
                m[ch, ys, np.clip(xs + 1, 0, 159)] = 1.0
        m[rng.uniform(size=m.shape) < rng.uniform(0, 0.02)] = 1.0
        h = ref.ref_lateral_create(0.5)
        try:
            o = RefOut()
            ref.ref_lateral_update(h, m.ctypes.data, 80, 160, 1920, 1080, C.byref(o))
        finally:
            ref.ref_lateral_destroy(h)
        fo = LT.LaneFilter(0.5).update(m)
        tr = LT.LaneTracker().update(fo.left, fo.right)
        assert bool(o.filt_left_valid) == (fo.left is not None) and bool(o.filt_right_valid) == (fo.right is not None), k
        sl, sr = LT.find_starting_points(m)
        nl = len(LT.sliding_window_search(m, sl, True)[1]) if sl else 0
        nr = len(LT.sliding_window_search(m, sr, False)[1]) if sr else 0
        assert (o.n_left_windows, o.n_right_windows) == (nl, nr), k
        if fo.left is not None:
            np.testing.assert_allclose(np.array(o.filt_left), fo.left, rtol=1e-8, atol=1e-8, err_msg=str(k))
        if fo.right is not None:
            np.testing.assert_allclose(np.array(o.filt_right), fo.right, rtol=1e-8, atol=1e-8, err_msg=str(k))
        assert bool(o.bev_valid) == tr.bev_valid, k
        if tr.bev_valid:
            np.testing.assert_allclose([o.lane_offset, o.yaw_offset, o.curvature], [tr.lane_offset, tr.yaw_offset, tr.curvature],
                                       rtol=1e-8, atol=1e-8, err_msg=str(k))
