// TEST INFRASTRUCTURE ONLY — C entry points around the UNMODIFIED reference classes LaneFilter, LaneTracker
// (VisionPilot/production_release/src/lane_filtering/lane_filter.cpp, src/lane_tracking/lane_tracking.cpp) and
// Estimator (src/path_planning/estimator.cpp), compiled together with them into oracle/_ref/libref_lateral.so
// by oracle/build_ref.py, plus PathFinder / fitQuadPoly (src/path_planning/path_finder.cpp, poly_fit.cpp).  OpenCV and
// Eigen are replaced by the stand-ins under oracle/cvstub/.  Used only by
// tests/test_oracle_lateral_vs_reference.py to pin oracle/lateral.py and oracle/post.py.
#include <array>
#include <cstring>
#include <vector>
#include "lane_filtering/lane_filter.hpp"
#include "lane_tracking/lane_tracking.hpp"
#include "path_planning/estimator.hpp"
#include "path_planning/path_finder.hpp"
#include "path_planning/poly_fit.hpp"

using namespace autoware_pov::vision::egolanes;

struct RefLateral {
  LaneFilter filter;
  LaneTracker tracker;
};

struct RefOut {
  double left[6], right[6], center[6], bev_left[6], bev_right[6], bev_center[6];
  double filt_left[6], filt_right[6];
  double lane_offset, yaw_offset, curvature, bev_lane_offset, bev_yaw_offset, bev_curvature, width_px;
  int left_valid, right_valid, path_valid, bev_valid, filt_left_valid, filt_right_valid;
  int n_left_windows, n_right_windows, n_bev_left, n_bev_right;
  float bev_left_pts[512], bev_right_pts[512];      // (x, y) pairs
};

static void copy6(const std::vector<double>& v, double* dst, int* valid) {
  *valid = v.empty() ? 0 : 1;
  for (int i = 0; i < 6; ++i) dst[i] = (i < static_cast<int>(v.size())) ? v[i] : 0.0;
}

extern "C" {

void* ref_lateral_create(float smoothing) { return new RefLateral{LaneFilter(smoothing), LaneTracker()}; }
void ref_lateral_destroy(void* h) { delete static_cast<RefLateral*>(h); }

// masks: float [3][H][W] (ego_left, ego_right, other_lanes)
void ref_lateral_update(void* h, const float* masks, int H, int W, int img_w, int img_h, RefOut* o) {
  RefLateral* r = static_cast<RefLateral*>(h);
  LaneSegmentation in;
  in.height = H; in.width = W;
  cv::Mat* dst[3] = {&in.ego_left, &in.ego_right, &in.other_lanes};
  for (int c = 0; c < 3; ++c) {
    *dst[c] = cv::Mat(H, W, CV_32FC1);
    std::memcpy(dst[c]->ptr<float>(0), masks + static_cast<size_t>(c) * H * W, sizeof(float) * H * W);
  }
  std::memset(o, 0, sizeof(*o));
  LaneSegmentation filt = r->filter.update(in);
  copy6(filt.left_coeffs, o->filt_left, &o->filt_left_valid);
  copy6(filt.right_coeffs, o->filt_right, &o->filt_right_valid);
  o->n_left_windows = static_cast<int>(filt.left_sliding_windows.size());
  o->n_right_windows = static_cast<int>(filt.right_sliding_windows.size());
  auto res = r->tracker.update(filt, cv::Size(img_w, img_h));
  const LaneSegmentation& out = res.first;
  const DualViewMetrics& m = res.second;
  copy6(out.left_coeffs, o->left, &o->left_valid);
  copy6(out.right_coeffs, o->right, &o->right_valid);
  int dummy;
  copy6(out.center_coeffs, o->center, &dummy);
  copy6(m.bev_visuals.bev_left_coeffs, o->bev_left, &dummy);
  copy6(m.bev_visuals.bev_right_coeffs, o->bev_right, &dummy);
  copy6(m.bev_visuals.bev_center_coeffs, o->bev_center, &dummy);
  o->path_valid = out.path_valid ? 1 : 0;
  o->bev_valid = m.bev_visuals.valid ? 1 : 0;
  o->lane_offset = m.orig_lane_offset; o->yaw_offset = m.orig_yaw_offset; o->curvature = m.orig_curvature;
  o->bev_lane_offset = m.bev_lane_offset; o->bev_yaw_offset = m.bev_yaw_offset; o->bev_curvature = m.bev_curvature;
  o->width_px = m.bev_visuals.last_valid_width_pixels;
  o->n_bev_left = static_cast<int>(m.bev_visuals.bev_left_pts.size());
  o->n_bev_right = static_cast<int>(m.bev_visuals.bev_right_pts.size());
  for (int i = 0; i < o->n_bev_left && i < 256; ++i) { o->bev_left_pts[2 * i] = m.bev_visuals.bev_left_pts[i].x; o->bev_left_pts[2 * i + 1] = m.bev_visuals.bev_left_pts[i].y; }
  for (int i = 0; i < o->n_bev_right && i < 256; ++i) { o->bev_right_pts[2 * i] = m.bev_visuals.bev_right_pts[i].x; o->bev_right_pts[2 * i + 1] = m.bev_visuals.bev_right_pts[i].y; }
}

// Estimator::update with PathFinder's fusion groups on a [14][2] (mean, variance) state, in place
void ref_estimator_update(double* state, const double* meas) {
  Estimator e;
  e.configureFusionGroups({{0, 3}, {5, 7}, {9, 11}});
  std::array<Gaussian, STATE_DIM> s, m;
  for (size_t i = 0; i < STATE_DIM; ++i) { s[i] = {state[2 * i], state[2 * i + 1]}; m[i] = {meas[2 * i], meas[2 * i + 1]}; }
  e.initialize(s);
  e.update(m);
  const auto& r = e.getState();
  for (size_t i = 0; i < STATE_DIM; ++i) { state[2 * i] = r[i].mean; state[2 * i + 1] = r[i].variance; }
}

// fitQuadPoly + FittedCurve (poly_fit.cpp:26-75): pts = n (x, y) float pairs -> coeff[3], cte, yaw_error
void ref_fit_quad(const float* pts, int n, double* coeff, double* cte_yaw) {
  std::vector<cv::Point2f> v;
  for (int i = 0; i < n; ++i) v.push_back(cv::Point2f(pts[2 * i], pts[2 * i + 1]));
  auto c = autoware_pov::vision::path_planning::fitQuadPoly(v);
  autoware_pov::vision::path_planning::FittedCurve fc(c);
  for (int i = 0; i < 3; ++i) coeff[i] = c[i];
  cte_yaw[0] = fc.cte; cte_yaw[1] = fc.yaw_error;
}

void* ref_pathfinder_create(double width) { return new autoware_pov::vision::path_planning::PathFinder(width); }
void ref_pathfinder_destroy(void* h) { delete static_cast<autoware_pov::vision::path_planning::PathFinder*>(h); }
// out[0..7] = cte, yaw_error, curvature, lane_width, variances x4; out[8] = fused_valid; state = [14][2]
void ref_pathfinder_update(void* h, const float* left, int nl, const float* right, int nr, double steering, double* out,
                           double* state) {
  auto* pf = static_cast<autoware_pov::vision::path_planning::PathFinder*>(h);
  std::vector<cv::Point2f> l, r;
  for (int i = 0; i < nl; ++i) l.push_back(cv::Point2f(left[2 * i], left[2 * i + 1]));
  for (int i = 0; i < nr; ++i) r.push_back(cv::Point2f(right[2 * i], right[2 * i + 1]));
  auto o = pf->update(l, r, steering);
  out[0] = o.cte; out[1] = o.yaw_error; out[2] = o.curvature; out[3] = o.lane_width;
  out[4] = o.cte_variance; out[5] = o.yaw_variance; out[6] = o.curv_variance; out[7] = o.lane_width_variance;
  out[8] = o.fused_valid ? 1.0 : 0.0;
  const auto& st = pf->getState();
  for (size_t i = 0; i < STATE_DIM; ++i) { state[2 * i] = st[i].mean; state[2 * i + 1] = st[i].variance; }
}

}  // extern "C"
