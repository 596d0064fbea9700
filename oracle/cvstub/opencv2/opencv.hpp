// TEST INFRASTRUCTURE ONLY — a minimal stand-in for the few OpenCV C++ types the reference's lateral
// post-process touches (cv::Mat of CV_32F / CV_64F, Point, Point2f, Rect, Size, solve(DECOMP_SVD),
// perspectiveTransform, Mat::inv), so that the UNMODIFIED reference sources
//   VisionPilot/production_release/src/lane_filtering/lane_filter.cpp
//   VisionPilot/production_release/src/lane_tracking/lane_tracking.cpp
// compile here (OpenCV's C++ headers are not in this image) into oracle/_ref/libref_lateral.so, against
// which oracle/lateral.py is pinned (tests/test_oracle_lateral_vs_reference.py).  Written from the OpenCV
// documentation of these functions; the numeric behaviour of the two that matter is itself pinned against
// the real library through its Python binding (cv2.solve / cv2.perspectiveTransform,
// tests/test_oracle_lateral.py):
//   solve(A, B, X, DECOMP_SVD)  -> minimum-norm least-squares solution (one-sided Jacobi SVD here)
//   perspectiveTransform        -> fp64 arithmetic on the float points, result rounded to float, w == 0 -> 0
#pragma once
#include <cmath>
#include <cstring>
#include <memory>
#include <vector>

#define CV_32F 5
#define CV_64F 6
#define CV_32FC1 CV_32F
#define CV_64FC1 CV_64F

namespace cv {

template <class T> struct Point_ {
  T x, y;
  Point_() : x(0), y(0) {}
  Point_(T x_, T y_) : x(x_), y(y_) {}
};
template <class T> inline Point_<T> operator+(const Point_<T>& a, const Point_<T>& b) { return Point_<T>(a.x + b.x, a.y + b.y); }
template <class T> inline Point_<T> operator*(const Point_<T>& a, float s) { return Point_<T>(static_cast<T>(a.x * s), static_cast<T>(a.y * s)); }
typedef Point_<int> Point;
typedef Point_<float> Point2f;

struct Size {
  int width, height;
  Size() : width(0), height(0) {}
  Size(int w, int h) : width(w), height(h) {}
  bool operator==(const Size& o) const { return width == o.width && height == o.height; }
};
struct Rect {
  int x, y, width, height;
  Rect() : x(0), y(0), width(0), height(0) {}
  Rect(int x_, int y_, int w, int h) : x(x_), y(y_), width(w), height(h) {}
};

enum { DECOMP_LU = 0, DECOMP_SVD = 1 };

class Mat {
public:
  int rows, cols;
  Mat() : rows(0), cols(0), type_(CV_64F) {}
  Mat(int r, int c, int type) : rows(r), cols(c), type_(type), buf_(new std::vector<unsigned char>(static_cast<size_t>(r) * c * esz(type), 0)) {}
  bool empty() const { return rows == 0 || cols == 0; }
  int type() const { return type_; }
  Mat clone() const {
    Mat m;
    m.rows = rows; m.cols = cols; m.type_ = type_;
    if (buf_) m.buf_.reset(new std::vector<unsigned char>(*buf_));
    return m;
  }
  template <class T> T& at(int i, int j) { return reinterpret_cast<T*>(buf_->data())[static_cast<size_t>(i) * cols + j]; }
  template <class T> const T& at(int i, int j) const { return reinterpret_cast<const T*>(buf_->data())[static_cast<size_t>(i) * cols + j]; }
  template <class T> T& at(int i) { return reinterpret_cast<T*>(buf_->data())[i]; }
  template <class T> const T& at(int i) const { return reinterpret_cast<const T*>(buf_->data())[i]; }
  template <class T> T* ptr(int r = 0) { return reinterpret_cast<T*>(buf_->data()) + static_cast<size_t>(r) * cols; }
  Mat inv(int = DECOMP_LU) const;            // 3x3 CV_64F only (the homography)
  static Mat zeros(int r, int c, int type) { return Mat(r, c, type); }
private:
  static size_t esz(int type) { return type == CV_32F ? 4 : 8; }
  int type_;
  std::shared_ptr<std::vector<unsigned char>> buf_;   // shared like cv::Mat headers; clone() copies
};

// cv::Mat_<double>(3,3) << a, b, c, ...   (comma initialiser)
template <class T> class Mat_ : public Mat {
public:
  Mat_(int r, int c) : Mat(r, c, sizeof(T) == 4 ? CV_32F : CV_64F) {}
};
template <class T> struct MatCommaInit_ {
  Mat_<T> m;
  int n;
  MatCommaInit_(const Mat_<T>& mm, T v) : m(mm), n(0) { m.template at<T>(n++) = v; }
  MatCommaInit_& operator,(T v) { m.template at<T>(n++) = v; return *this; }
  operator Mat() const { return m; }
};
template <class T> inline MatCommaInit_<T> operator<<(const Mat_<T>& m, T v) { return MatCommaInit_<T>(m, v); }

inline Mat Mat::inv(int) const {
  Mat r(3, 3, CV_64F);
  const double a = at<double>(0, 0), b = at<double>(0, 1), c = at<double>(0, 2), d = at<double>(1, 0), e = at<double>(1, 1),
               f = at<double>(1, 2), g = at<double>(2, 0), h = at<double>(2, 1), i = at<double>(2, 2);
  const double det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g), id = 1.0 / det;
  r.at<double>(0, 0) = (e * i - f * h) * id; r.at<double>(0, 1) = (c * h - b * i) * id; r.at<double>(0, 2) = (b * f - c * e) * id;
  r.at<double>(1, 0) = (f * g - d * i) * id; r.at<double>(1, 1) = (a * i - c * g) * id; r.at<double>(1, 2) = (c * d - a * f) * id;
  r.at<double>(2, 0) = (d * h - e * g) * id; r.at<double>(2, 1) = (b * g - a * h) * id; r.at<double>(2, 2) = (a * e - b * d) * id;
  return r;
}

// Minimum-norm least squares  min ||A x - b||  through a one-sided Jacobi SVD of A (n x k, k <= 8), fp64.
inline bool solve(const Mat& A, const Mat& B, Mat& X, int /*flags*/) {
  const int n = A.rows, k = A.cols;
  if (n <= 0 || k <= 0 || k > 8) return false;
  std::vector<double> U(static_cast<size_t>(n) * k), V(static_cast<size_t>(k) * k, 0.0);
  for (int i = 0; i < n; ++i) for (int j = 0; j < k; ++j) U[static_cast<size_t>(i) * k + j] = A.at<double>(i, j);
  for (int j = 0; j < k; ++j) V[static_cast<size_t>(j) * k + j] = 1.0;
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0.0;
    for (int p = 0; p < k - 1; ++p)
      for (int q = p + 1; q < k; ++q) {
        double al = 0, be = 0, ga = 0;
        for (int i = 0; i < n; ++i) {
          const double up = U[static_cast<size_t>(i) * k + p], uq = U[static_cast<size_t>(i) * k + q];
          al += up * up; be += uq * uq; ga += up * uq;
        }
        if (ga == 0.0 || std::fabs(ga) <= 1e-300) continue;
        off = std::fmax(off, std::fabs(ga) / std::sqrt(al * be + 1e-300));
        const double zeta = (be - al) / (2.0 * ga);
        const double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
        const double cs = 1.0 / std::sqrt(1.0 + t * t), sn = cs * t;
        for (int i = 0; i < n; ++i) {
          double& up = U[static_cast<size_t>(i) * k + p]; double& uq = U[static_cast<size_t>(i) * k + q];
          const double a0 = up, b0 = uq;
          up = cs * a0 - sn * b0; uq = sn * a0 + cs * b0;
        }
        for (int i = 0; i < k; ++i) {
          double& vp = V[static_cast<size_t>(i) * k + p]; double& vq = V[static_cast<size_t>(i) * k + q];
          const double a0 = vp, b0 = vq;
          vp = cs * a0 - sn * b0; vq = sn * a0 + cs * b0;
        }
      }
    if (off < 1e-15) break;
  }
  double sv[8], smax = 0.0;
  for (int j = 0; j < k; ++j) {
    double s2 = 0;
    for (int i = 0; i < n; ++i) s2 += U[static_cast<size_t>(i) * k + j] * U[static_cast<size_t>(i) * k + j];
    sv[j] = std::sqrt(s2);
    smax = std::fmax(smax, sv[j]);
  }
  const double thr = smax * 2.220446049250313e-16 * std::fmax(n, k) * 4.0;
  X = Mat(k, 1, CV_64F);
  for (int j = 0; j < k; ++j) {
    if (sv[j] <= thr) continue;                       // null direction: minimum-norm solution leaves it out
    double ub = 0;
    for (int i = 0; i < n; ++i) ub += U[static_cast<size_t>(i) * k + j] * B.at<double>(i, 0);
    const double coef = ub / (sv[j] * sv[j]);          // (u_j . b) / s_j  with u_j = U_j / s_j
    for (int i = 0; i < k; ++i) X.at<double>(i) += coef * V[static_cast<size_t>(i) * k + j];
  }
  return true;
}

inline void perspectiveTransform(const std::vector<Point2f>& src, std::vector<Point2f>& dst, const Mat& H) {
  dst.resize(src.size());
  const double* m = &H.at<double>(0);
  for (size_t i = 0; i < src.size(); ++i) {
    const double x = src[i].x, y = src[i].y;
    double w = x * m[6] + y * m[7] + m[8];
    if (std::fabs(w) > 2.220446049250313e-16) {
      w = 1.0 / w;
      dst[i] = Point2f(static_cast<float>((x * m[0] + y * m[1] + m[2]) * w), static_cast<float>((x * m[3] + y * m[4] + m[5]) * w));
    } else {
      dst[i] = Point2f(0.f, 0.f);
    }
  }
}

}  // namespace cv
