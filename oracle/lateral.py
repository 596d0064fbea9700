"""TEST INFRASTRUCTURE ONLY — CPU restatement of the production lateral post-process that follows
EgoLanes: LaneFilter (sliding-window search + poly-fit + temporal smoothing)
and LaneTracker (BEV homography warp, lane-width recovery, curve parameters).

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.

Reference (read-only, restated, not copied):
  VisionPilot/production_release/src/lane_filtering/lane_filter.cpp
      :325-370  findStartingPoints        -> find_starting_points
      :376-590  slidingWindowSearch       -> sliding_window_search
      :56-113   fitPolySimple             -> lstsq_poly            (cv::solve DECOMP_SVD, fp64)
      :116-218  fitPoly                   -> fit_poly
      :232-323  update                    -> LaneFilter.update
  VisionPilot/production_release/src/lane_tracking/lane_tracking.cpp
      :36-300   update                    -> LaneTracker.update
      :305-318  warpPoints                -> warp_points           (cv::perspectiveTransform, CV_32FC2)
      :320-348  genPointsFromCoeffs       -> gen_points
      :350-404  fitPoly2ndOrder           -> fit_poly2
      :406-452  calcLaneOffset / calcYawOffset / calcCurvature

Pinning (two layers):
  1. Against the reference's OWN sources: oracle/build_ref.py compiles the unmodified lane_filter.cpp,
     lane_tracking.cpp, estimator.cpp, poly_fit.cpp and path_finder.cpp (where they lie under /root/reference)
     into oracle/_ref/libref_lateral.so with the OpenCV / Eigen names they use provided by the minimal
     stand-ins under oracle/cvstub/; tests/test_oracle_lateral_vs_reference.py runs 240 frames of stateful
     sequences (dropouts, empty / noise-only / single-row masks) through both: validity flags, window counts
     and BEV points identical, all coefficients and curve parameters within 1e-9.
  2. The two numeric OpenCV functions the stand-in re-implements are themselves pinned against the real
     library through its Python binding: lstsq_poly against cv2.solve(DECOMP_SVD) incl. rank-deficient
     minimum-norm systems, warp_points bit-exact against cv2.perspectiveTransform, the inverse homography
     against cv2.invert (tests/test_oracle_lateral.py).
  PathFinder::update and fitQuadPoly are pinned in layer 1 too (state variances 1e-12, means 1e-3: the
  reference's predict step adds an unseeded +-1e-5 jitter).

Note on RANSAC (lane_filter.cpp:157-191): `best_inliers` starts as ALL points and a candidate model
only replaces it when it has strictly MORE inliers than that — impossible — so the loop never changes
the result and fitPoly is exactly a least-squares fit of all points (order 1 below 30 points, else 2).
The restatement therefore has no random sampler, and neither has the device kernel; layer 1 above confirms
it (the compiled reference, with its unseeded mt19937, agrees on every frame).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import numpy as np

# lane_filter.hpp:29-47
ROI_Y_MIN, ROI_Y_MAX = 40, 79
WIN_H = 4
MIN_PIXELS_FOR_FIT = 4
CONSECUTIVE_EMPTY_THRESHOLD = 12
MIN_WIN_W, MAX_WIN_W = 1, 6
HEIGHT_THRESHOLD = 40
PRIORITY_Y_THRESHOLD = 40

# lane_tracking.hpp:75-79
H_ORIG_TO_BEV = np.array([[-1.79887412e-01, -6.05811422e-01, 6.02998251e+02],
                          [1.85824549e-14, -1.28170839e+00, 8.63871455e+02],
                          [2.95628463e-17, -1.76125061e-03, 1.00000000e+00]], dtype=np.float64)


def _round_half_away(v: float) -> int:
    """std::round on a float."""
    return int(math.floor(abs(v) + 0.5)) * (1 if v >= 0 else -1)


def find_starting_points(masks: np.ndarray) -> Tuple[Optional[Tuple[int, int]], Optional[Tuple[int, int]]]:
    """masks [3][H][W] float (ego_left, ego_right, other).  Returns ((x,y) | None, (x,y) | None)."""
    W = masks.shape[2]
    mid = W // 2
    left = right = None
    for y in range(ROI_Y_MAX, ROI_Y_MIN - 1, -1):
        for x in range(mid - 1, -1, -1):
            if masks[0, y, x] > 0.5:
                left = (x, y)
                break
        if left:
            break
    for y in range(ROI_Y_MAX, ROI_Y_MIN - 1, -1):
        for x in range(mid, W):
            if masks[1, y, x] > 0.5:
                right = (x, y)
                break
        if right:
            break
    return left, right


def sliding_window_search(masks: np.ndarray, start: Tuple[int, int], is_left: bool):
    """Returns (points [(x,y)...] in the reference's push order, windows [(x,y,w,h)...])."""
    H, W = masks.shape[1], masks.shape[2]
    ego = masks[0] if is_left else masks[1]
    other = masks[2]
    pts: List[Tuple[int, int]] = []
    wins: List[Tuple[int, int, int, int]] = []
    f32 = np.float32

    def run(step_y: int):
        cx, cy = start
        if step_y > 0:
            cy += WIN_H
        dir_x, dir_y = f32(0.0), f32(step_y)
        empty = 0
        max_steps = int(H / float(WIN_H))
        for _ in range(max_steps):
            if cx < 0 or cx >= W:
                break
            if step_y < 0 and cy < 0:
                break
            if step_y > 0 and cy >= H:
                break
            cw = MIN_WIN_W if cy < HEIGHT_THRESHOLD else MAX_WIN_W
            if step_y < 0:
                y_lo, y_hi = max(0, cy - WIN_H), cy
            else:
                y_lo, y_hi = cy, min(H, cy + WIN_H)
            x_lo, x_hi = max(0, cx - cw), min(W, cx + cw)
            wins.append((x_lo, y_lo, x_hi - x_lo, y_hi - y_lo))
            strict = cy < PRIORITY_Y_THRESHOLD
            ego_px, oth_px = [], []
            for y in range(y_lo, y_hi):
                for x in range(x_lo, x_hi):
                    if ego[y, x] > 0.5:
                        ego_px.append((x, y))
                    if (not strict) and other[y, x] > 0.5:
                        oth_px.append((x, y))
            chosen = ego_px if len(ego_px) >= 3 else (oth_px if len(oth_px) >= 3 else None)
            if chosen is not None:
                pts.extend(chosen)
                n = len(chosen)
                cxf = f32(sum(p[0] for p in chosen)) / f32(n)     # static_cast<float>(long) / size
                cyf = f32(sum(p[1] for p in chosen)) / f32(n)
                empty = 0
                dx, dy = cxf - f32(cx), cyf - f32(cy)
                ln = f32(math.sqrt(float(dx * dx + dy * dy)))       # std::sqrt(float)
                if ln > f32(0.1):
                    dir_x, dir_y = dx / ln, dy / ln
                cx, cy = _round_half_away(float(cxf)), _round_half_away(float(cyf))
            else:
                if step_y < 0 and cy < H * 0.25:
                    break
                empty += 1
                if empty >= CONSECUTIVE_EMPTY_THRESHOLD:
                    break
                cx += int(float(dir_x * f32(WIN_H)))                 # static_cast<int>: truncation
                cy += int(float(dir_y * f32(WIN_H)))
            if step_y < 0 and cy >= y_hi - 1:
                cy -= WIN_H
            if step_y > 0 and cy <= y_lo + 1:
                cy += WIN_H

    run(-1)
    run(1)
    return pts, wins


def lstsq_poly(ys: np.ndarray, xs: np.ndarray, order: int) -> Optional[np.ndarray]:
    """fitPolySimple: x = c0*y^order + ... ; minimum-norm least squares in fp64 (SVD)."""
    n = len(ys)
    if n <= order:
        return None
    y = np.asarray(ys, dtype=np.float64)
    A = np.stack([y ** k for k in range(order, -1, -1)], axis=1)
    sol, *_ = np.linalg.lstsq(A, np.asarray(xs, dtype=np.float64), rcond=None)
    return sol


def fit_poly(points: List[Tuple[int, int]]) -> Optional[np.ndarray]:
    """fitPoly -> 6 coefficients [c3, c2, c1, c0, min_y, max_y] or None (invalid)."""
    n = len(points)
    if n < MIN_PIXELS_FOR_FIT:
        return None
    ys = np.array([p[1] for p in points], dtype=np.float64)
    xs = np.array([p[0] for p in points], dtype=np.float64)
    order = 1 if n < 30 else 2
    sol = lstsq_poly(ys, xs, order)       # RANSAC never replaces the all-points inlier set (see module doc)
    if sol is None:
        return None
    out = np.zeros(6, dtype=np.float64)
    if order == 1:
        out[2], out[3] = sol
    else:
        out[1], out[2], out[3] = sol
    out[4], out[5] = ys.min(), ys.max()
    return out


@dataclass
class FilterOut:
    left: Optional[np.ndarray] = None       # 6 coefficients (after smoothing) or None
    right: Optional[np.ndarray] = None
    left_start: Tuple[int, int] = (-1, -1)
    right_start: Tuple[int, int] = (-1, -1)
    n_left: int = 0
    n_right: int = 0
    left_pts: list = field(default_factory=list)
    right_pts: list = field(default_factory=list)


class LaneFilter:
    """lane_filter.cpp:232-323 (state: previous fits, smoothing factor 0.5 by default)."""

    def __init__(self, smoothing: float = 0.5):
        self.s = np.float32(smoothing)
        self.prev_left: Optional[np.ndarray] = None
        self.prev_right: Optional[np.ndarray] = None

    def _smooth(self, cur: np.ndarray, prev: Optional[np.ndarray]) -> np.ndarray:
        if prev is None:
            return cur
        # float factor promoted to double: s*cur + (1.0f - s)*prev
        a, b = float(self.s), float(np.float32(1.0) - self.s)
        return a * cur + b * prev

    def update(self, masks: np.ndarray) -> FilterOut:
        o = FilterOut()
        sl, sr = find_starting_points(masks)
        if sl is not None:
            o.left_start = sl
            o.left_pts, _ = sliding_window_search(masks, sl, True)
            o.n_left = len(o.left_pts)
            fit = fit_poly(o.left_pts)
            if fit is not None:
                fit = self._smooth(fit, self.prev_left)
                self.prev_left = fit
                o.left = fit
        else:
            self.prev_left = None
        if sr is not None:
            o.right_start = sr
            o.right_pts, _ = sliding_window_search(masks, sr, False)
            o.n_right = len(o.right_pts)
            fit = fit_poly(o.right_pts)
            if fit is not None:
                fit = self._smooth(fit, self.prev_right)
                self.prev_right = fit
                o.right = fit
        else:
            self.prev_right = None
        return o


# ----------------------------------------------------------------------------------------- LaneTracker
def warp_points(pts: np.ndarray, Hm: np.ndarray) -> np.ndarray:
    """cv::perspectiveTransform on CV_32FC2 points with a 3x3 double matrix: fp64 arithmetic on the
    float inputs, result rounded to float; w == 0 -> (0, 0)."""
    pts = np.asarray(pts, dtype=np.float32).reshape(-1, 2)
    x, y = pts[:, 0].astype(np.float64), pts[:, 1].astype(np.float64)
    m = Hm.reshape(-1)
    w = x * m[6] + y * m[7] + m[8]
    ok = np.abs(w) > np.finfo(np.float64).eps
    wi = np.where(ok, 1.0 / np.where(ok, w, 1.0), 0.0)
    out = np.stack([(x * m[0] + y * m[1] + m[2]) * wi, (x * m[3] + y * m[4] + m[5]) * wi], axis=1)
    return out.astype(np.float32)


def gen_points(c: np.ndarray, step: int = 5) -> np.ndarray:
    """genPointsFromCoeffs: y from min_y to max_y in steps of 5 (double loop variable), float points."""
    out = []
    y = float(c[4])
    while y <= float(c[5]):
        x = c[1] * y * y + c[2] * y + c[3] if c[1] != 0 else c[2] * y + c[3]
        out.append((np.float32(x), np.float32(y)))
        y += step
    return np.array(out, dtype=np.float32).reshape(-1, 2)


def fit_poly2(pts: np.ndarray) -> np.ndarray:
    """fitPoly2ndOrder: 6 coefficients [0, a, b, c, min_y, max_y]; zeros when fewer than 3 points."""
    out = np.zeros(6, dtype=np.float64)
    pts = np.asarray(pts, dtype=np.float32).reshape(-1, 2)
    if len(pts) < 3:
        return out
    y = pts[:, 1].astype(np.float64)
    sol = lstsq_poly(y, pts[:, 0].astype(np.float64), 2)
    out[1], out[2], out[3] = sol
    out[4], out[5] = y.min(), y.max()
    return out


def lane_offset(c, y):
    return c[1] * y * y + c[2] * y + c[3]


def yaw_offset(c, y):
    return math.atan(2 * c[1] * y + c[2])


def curvature(c, y):
    d1 = 2 * c[1] * y + c[2]
    den = math.pow(1 + d1 * d1, 1.5)
    return 0.0 if abs(den) < 1e-6 else abs(2 * c[1]) / den


@dataclass
class TrackOut:
    left: Optional[np.ndarray] = None
    right: Optional[np.ndarray] = None
    center: Optional[np.ndarray] = None
    path_valid: bool = False
    lane_offset: float = 0.0
    yaw_offset: float = 0.0
    curvature: float = 0.0
    bev_lane_offset: float = 0.0
    bev_yaw_offset: float = 0.0
    bev_curvature: float = 0.0
    bev_left: Optional[np.ndarray] = None
    bev_right: Optional[np.ndarray] = None
    bev_center: Optional[np.ndarray] = None
    width_px: float = 0.0
    bev_valid: bool = False
    bev_left_pts: Optional[np.ndarray] = None     # BEVVisuals.bev_left_pts / bev_right_pts (after recovery)
    bev_right_pts: Optional[np.ndarray] = None


class LaneTracker:
    """lane_tracking.cpp:36-300 (state: smoothed BEV lane width)."""

    def __init__(self):
        self.H = H_ORIG_TO_BEV
        self.Hinv = np.linalg.inv(H_ORIG_TO_BEV)
        self.width = 180.0
        self.has_width = False

    def update(self, left: Optional[np.ndarray], right: Optional[np.ndarray], model_wh=(160, 80),
               image_wh=(1920, 1080)) -> TrackOut:
        o = TrackOut(left=None if left is None else left.copy(), right=None if right is None else right.copy())
        sx, sy = image_wh[0] / model_wh[0], image_wh[1] / model_wh[1]

        def upscale(c):
            up = np.zeros(6)
            up[1] = c[1] * sx / (sy * sy)
            up[2] = c[2] * sx / sy
            up[3] = c[3] * sx
            up[4] = c[4] * sy
            up[5] = c[5] * sy
            return up

        lv, rv = left is not None, right is not None
        lb = warp_points(gen_points(upscale(left)), self.H) if lv else np.zeros((0, 2), np.float32)
        rb = warp_points(gen_points(upscale(right)), self.H) if rv else np.zeros((0, 2), np.float32)

        def recover(bev_pts):
            orig = warp_points(bev_pts, self.Hinv)
            # Point2f(p.x / scale_x, p.y / scale_y): float / double -> double, stored as float
            model = np.stack([orig[:, 0].astype(np.float64) / sx, orig[:, 1].astype(np.float64) / sy], axis=1).astype(np.float32)
            return fit_poly2(model)

        if lv and rv:
            if len(lb) and len(rb):
                w = abs(float(rb[-1, 0] - lb[-1, 0]))         # float subtraction, then |.| in double
                self.width = self.width * 0.9 + w * 0.1 if self.has_width else w
                self.has_width = True
        elif (not lv) and rv and self.has_width:
            lb = rb.copy()
            lb[:, 0] = (lb[:, 0].astype(np.float64) - self.width).astype(np.float32)   # p.x -= double
            o.left = recover(lb)
        elif lv and (not rv) and self.has_width:
            rb = lb.copy()
            rb[:, 0] = (rb[:, 0].astype(np.float64) + self.width).astype(np.float32)
            o.right = recover(rb)

        if len(lb) and len(rb):
            n = min(len(lb), len(rb))
            center = ((lb[:n] + rb[:n]) * np.float32(0.5)).astype(np.float32)
            o.bev_center = fit_poly2(center)
            o.bev_left, o.bev_right = fit_poly2(lb), fit_poly2(rb)
            o.bev_lane_offset = lane_offset(o.bev_center, 640.0) - 320.0
            o.bev_yaw_offset = yaw_offset(o.bev_center, 640.0)
            o.bev_curvature = curvature(o.bev_center, 640.0)
            o.center = (o.left + o.right) / 2.0
            o.path_valid = True
            o.lane_offset = lane_offset(o.center, 79.0) - model_wh[0] / 2.0
            o.yaw_offset = yaw_offset(o.center, 79.0)
            o.curvature = curvature(o.center, 79.0)
            o.width_px = self.width
            o.bev_valid = True
            o.bev_left_pts, o.bev_right_pts = lb, rb
        return o


# ----------------------------------------------------------------------------------------- PathFinder
def pixels_to_meters(px: np.ndarray) -> np.ndarray:
    """transformPixelsToMeters (production_release/main.cpp:333-357): 640 px = 40 m, vehicle at (320, 640);
    double arithmetic on the float pixel coordinates, stored as float."""
    px = np.asarray(px, dtype=np.float32).reshape(-1, 2)
    scale = 40.0 / 640.0
    return np.stack([(px[:, 0].astype(np.float64) - 320.0) * scale,
                     (640.0 - px[:, 1].astype(np.float64)) * scale], axis=1).astype(np.float32)


class PathFinder:
    """PathFinder::update (src/path_planning/path_finder.cpp:48-181) on the LaneTracker's BEV points.
    The predict step's process MEAN is drawn from an unseeded U(-1e-5, 1e-5) in the reference
    (path_finder.cpp:59-68); it is taken as 0 here and in the device kernel (five orders of magnitude below
    the measurement noise), the process variance PROC_SD^2 = 0.25 is added as in the reference."""

    def __init__(self, default_lane_width: float = 4.0):
        from . import post
        self._post = post
        self.default_width = default_lane_width
        self.state = post.initial_state(default_lane_width)

    def update(self, left_px: np.ndarray, right_px: np.ndarray, steering_rad: float) -> dict:
        post = self._post
        self.state[:, 1] += 0.5 * 0.5                                   # Estimator::predict
        lm, rm = pixels_to_meters(left_px), pixels_to_meters(right_px)
        lc = post.polyfit(lm[:, 0], lm[:, 1], 2) if len(lm) > 2 else np.full(3, np.nan)
        rc = post.polyfit(rm[:, 0], rm[:, 1], 2) if len(rm) > 2 else np.full(3, np.nan)
        width = self.state[12, 0]
        meas = post.pathfinder_measurement(lc, rc, steering_rad, width, self.default_width)
        self.state = post.estimator_update(self.state, meas)
        l_cte, l_yaw = post.fitted_curve(lc)
        r_cte, r_yaw = post.fitted_curve(rc)
        st = self.state
        return {"left_coeff": lc, "right_coeff": rc, "left_cte": l_cte, "left_yaw_error": l_yaw,
                "right_cte": r_cte, "right_yaw_error": r_yaw, "cte": st[3, 0], "yaw_error": st[7, 0],
                "curvature": steering_rad, "lane_width": st[12, 0], "cte_variance": st[3, 1],
                "yaw_variance": st[7, 1], "curv_variance": st[11, 1], "lane_width_variance": st[12, 1],
                "fused_valid": not (math.isnan(st[3, 0]) or math.isnan(st[7, 0]) or math.isnan(steering_rad))}


def synth_lane_masks(seed: int, H: int = 80, W: int = 160, drop_left=False, drop_right=False, noise=0.01):
    """Plausible EgoLanes masks: two converging ego lines + an outer line, 1-2 px wide, with dropouts and
    salt noise (test input generator, not reference behaviour)."""
    rng = np.random.default_rng(seed)
    m = np.zeros((3, H, W), dtype=np.float32)
    vx = W / 2 + rng.uniform(-10, 10)
    bl, br = W * 0.25 + rng.uniform(-8, 8), W * 0.75 + rng.uniform(-8, 8)
    curve = rng.uniform(-0.004, 0.004)
    for y in range(int(H * 0.3), H):
        t = (y - H * 0.3) / (H * 0.7)
        xl = vx + (bl - vx) * t + curve * (H - y) ** 2
        xr = vx + (br - vx) * t + curve * (H - y) ** 2
        xo = vx + (br + 45 - vx) * t + curve * (H - y) ** 2
        wpx = 1 if t < 0.4 else 2
        for k, (x, drop) in enumerate(((xl, drop_left), (xr, drop_right), (xo, False))):
            if drop or rng.uniform() < 0.12:
                continue
            x0 = int(round(x))
            for d in range(wpx):
                if 0 <= x0 + d < W:
                    m[k, y, x0 + d] = 1.0
    salt = rng.uniform(size=m.shape) < noise
    m[salt] = 1.0
    return m
