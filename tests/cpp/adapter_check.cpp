// Compiles both adapters against the stub headers and links them to libvp_b200.so.
// Without a GPU it checks the reference's error contract (ctor throws std::runtime_error);
// with a GPU and argv = <scene_seg.vpw> <ego_lanes.vpw> it runs one frame through each.
#include <cstdio>
#include <cstring>
#include <memory>
#include "cv_stub.hpp"
#include "../../adapters/b200_backend.hpp"
#include "../../adapters/egolanes_b200_engine.hpp"

using autoware_pov::vision::B200Backend;
using autoware_pov::vision::InferenceBackend;
using autoware_pov::vision::egolanes::EgoLanesB200Engine;

int main(int argc, char** argv) {
  if (argc < 3) {
    int thrown = 0;
    try { B200Backend b("/nonexistent.vpw", "fp16", 0); } catch (const std::runtime_error& e) { ++thrown; std::printf("ctor threw: %s\n", e.what()); }
    try { EgoLanesB200Engine e("/nonexistent.vpw"); } catch (const std::runtime_error& e) { ++thrown; std::printf("ctor threw: %s\n", e.what()); }
    std::printf("ADAPTER_CTOR_THROWS %d\n", thrown);
    return thrown == 2 ? 0 : 1;
  }
  cv::Mat frame(1080, 1920, 0);
  for (size_t i = 0; i < frame.store.size(); ++i) frame.store[i] = static_cast<unsigned char>((i * 2654435761u) >> 24);
  std::unique_ptr<InferenceBackend> be(new B200Backend(argv[1], "fp16", 0, VP_SCENE_SEG));
  bool threw = false;
  try { be->getRawTensorData(); } catch (const std::runtime_error&) { threw = true; }
  if (!threw) return 2;
  if (!be->doInference(frame)) { std::printf("doInference failed: %s\n", vp_last_error()); return 3; }
  auto shp = be->getTensorShape();
  std::printf("SCENESEG_SHAPE %ld %ld %ld %ld first=%f\n", (long)shp[0], (long)shp[1], (long)shp[2], (long)shp[3], be->getRawTensorData()[0]);
  EgoLanesB200Engine eg(argv[2]);
  cv::Mat crop = frame;            // caller-side crop of rows >= 420 (main.cpp:497-502): pointer + rows
  crop.data = frame.data + 420 * frame.step; crop.rows = 1080 - 420; crop.store.clear();
  auto seg = eg.inference(crop, 0.0f);
  auto es = eg.getTensorShape();
  double s = 0; for (int i = 0; i < seg.height * seg.width; ++i) s += seg.ego_left.ptr<float>(0)[i];
  std::printf("EGOLANES_SHAPE %ld %ld %ld %ld mask %dx%d left_sum=%.0f\n", (long)es[0], (long)es[1], (long)es[2], (long)es[3], seg.height, seg.width, s);
  return (shp[1] == 3 && shp[2] == 320 && shp[3] == 640 && es[1] == 3 && es[2] == 80 && es[3] == 160) ? 0 : 4;
}
