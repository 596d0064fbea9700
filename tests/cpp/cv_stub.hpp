// Minimal stand-in for the parts of OpenCV / the reference headers the adapters touch, so that
// they can be syntax- and link-checked in an image without OpenCV (tests/test_adapters_cpu.py).
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>
#define CV_32FC1 5
namespace cv {
struct Point { int x, y; };
struct Rect { int x, y, width, height; };
struct Mat {
  int rows = 0, cols = 0; size_t step = 0; unsigned char* data = nullptr; int type_ = 0;
  std::vector<unsigned char> store;
  Mat() = default;
  Mat(int r, int c, int type) : rows(r), cols(c), type_(type) {
    const int es = type == CV_32FC1 ? 4 : 3;
    step = static_cast<size_t>(c) * es; store.resize(step * r); data = store.data();
  }
  Mat(const Mat& o) : rows(o.rows), cols(o.cols), step(o.step), type_(o.type_), store(o.store) { data = store.empty() ? o.data : store.data(); }
  Mat& operator=(const Mat& o) { rows = o.rows; cols = o.cols; step = o.step; type_ = o.type_; store = o.store; data = store.empty() ? o.data : store.data(); return *this; }
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
  int channels() const { return type_ == CV_32FC1 ? 1 : 3; }
  int depth() const { return type_ == CV_32FC1 ? 5 : 0; }
  template <class T> T* ptr(int r) { return reinterpret_cast<T*>(data + step * r); }
};
}  // namespace cv
// reference: common/include/inference_backend_base.hpp:14-27
namespace autoware_pov::vision {
class InferenceBackend {
public:
  virtual ~InferenceBackend() = default;
  virtual bool doInference(const cv::Mat& input_image) = 0;
  virtual const float* getRawTensorData() const = 0;
  virtual std::vector<int64_t> getTensorShape() const = 0;
  virtual int getModelInputHeight() const = 0;
  virtual int getModelInputWidth() const = 0;
};
}  // namespace autoware_pov::vision
// reference: production_release/include/inference/lane_segmentation.hpp:16-44 (fields the adapter fills)
namespace autoware_pov::vision::egolanes {
struct LaneSegmentation {
  cv::Mat ego_left, ego_right, other_lanes;
  int height = 0, width = 0;
};
}  // namespace autoware_pov::vision::egolanes
