// egolanes_b200_engine.hpp — header-only adapter with the interface of the reference's
// EgoLanesTensorRTEngine / EgoLanesOnnxEngine on top of libvp_b200.so.
//
// Drop-in for (reference, VisionPilot/production_release):
//   include/inference/tensorrt_engine.hpp:40-148, src/inference/tensorrt_engine.cpp:190-320
//   include/inference/lane_segmentation.hpp:16-44 (LaneSegmentation)
// used by lateralInferenceThread (main.cpp:505-517) behind the compile-time switch at main.cpp:13-27.
//
// Behaviour kept:
//   * ctor (model_path, precision = "fp16", device_id = 0, cache_dir) throws std::runtime_error on failure
//   * inference(bgr, threshold) -> LaneSegmentation with three CV_32FC1 masks (v > threshold ? 1 : 0,
//     tensorrt_engine.cpp:264-305), height/width of the model OUTPUT; empty struct on failure (:255-258)
//   * getRawTensorData() / getTensorShape() ([1,3,H,W]) / getInput*/getOutput* accessors
//   * pre-process convention P0c: INTER_LINEAR resize, BGR->RGB, RGB ImageNet stats (:190-220) on the GPU;
//     the caller's crop of rows >= 420 (main.cpp:497-502) is just a cv::Mat ROI (pointer + step).
// Include after lane_segmentation.hpp (needs cv::Mat and LaneSegmentation).
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../include/vp_b200.h"

namespace autoware_pov::vision::egolanes
{

class EgoLanesB200Engine
{
public:
  EgoLanesB200Engine(const std::string & model_path, const std::string & precision = "fp16",
                     int device_id = 0, const std::string & /*cache_dir*/ = "")
  {
    vp_engine_config cfg{};
    cfg.gpu_id = device_id;
    if (precision != "fp16" && precision != "bf16" && precision != "fp32")
      throw std::runtime_error("EgoLanesB200Engine: unsupported precision '" + precision + "' (fp16 | bf16 | fp32)");
    cfg.dtype = (precision == "bf16") ? VPB_BF16 : VPB_F16;
    cfg.precision = (precision == "fp32") ? VP_PREC_SPLIT : VP_PREC_16;   // fp32 engines: split-fp16 fp32-grade mode
    cfg.resize_mode = VPB_RESIZE_CV_LINEAR;
    cfg.convention = VPB_CONV_BGR_SWAP;
    cfg.n_models = 1;
    cfg.kinds[0] = VP_EGO_LANES;
    cfg.weights[0] = model_path.c_str();
    cfg.fetch_raw = 1;
    cfg.use_graph = 1;
    if (vp_engine_create(&cfg, &engine_) != VPB_OK) {
      throw std::runtime_error(std::string("EgoLanesB200Engine: ") + vp_last_error());
    }
  }
  ~EgoLanesB200Engine() { vp_engine_destroy(engine_); }
  EgoLanesB200Engine(const EgoLanesB200Engine &) = delete;
  EgoLanesB200Engine & operator=(const EgoLanesB200Engine &) = delete;

  LaneSegmentation inference(const cv::Mat & input_image, float threshold = 0.0f)
  {
    if (input_image.empty() ||
        vp_engine_infer(engine_, input_image.data, input_image.rows, input_image.cols,
                        static_cast<int>(input_image.step)) != VPB_OK ||
        vp_engine_output(engine_, 0, &out_) != VPB_OK) {
      return LaneSegmentation{};
    }
    ran_ = true;
    LaneSegmentation result;
    result.height = out_.height;
    result.width = out_.width;
    const int n = out_.height * out_.width;
    cv::Mat * masks[3] = {&result.ego_left, &result.ego_right, &result.other_lanes};
    for (int c = 0; c < 3; ++c) {
      *masks[c] = cv::Mat(out_.height, out_.width, CV_32FC1);
      const float * src = out_.raw_host + static_cast<size_t>(c) * n;
      float * dst = masks[c]->ptr<float>(0);
      for (int i = 0; i < n; ++i) dst[i] = (src[i] > threshold) ? 1.0f : 0.0f;
    }
    return result;
  }

  const float * getRawTensorData() const
  {
    if (!ran_) throw std::runtime_error("Inference has not been run yet. Call inference() first.");
    return out_.raw_host;
  }
  // Not in the reference interface: the same tensor while it is still in HBM, for the on-device
  // lateral post-process (vpb_lane_masks + vpb_lateral_update).
  const float * rawDevice() const
  {
    if (!ran_) throw std::runtime_error("Inference has not been run yet. Call inference() first.");
    return out_.raw_dev;
  }
  std::vector<int64_t> getTensorShape() const
  {
    return {1, static_cast<int64_t>(out_.channels), static_cast<int64_t>(out_.height), static_cast<int64_t>(out_.width)};
  }
  int getInputWidth() const { return 640; }
  int getInputHeight() const { return 320; }
  int getOutputWidth() const { return ran_ ? out_.width : 160; }
  int getOutputHeight() const { return ran_ ? out_.height : 80; }

private:
  vp_engine * engine_{nullptr};
  vp_output out_{};
  bool ran_{false};
};

}  // namespace autoware_pov::vision::egolanes
