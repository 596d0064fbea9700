/* vp_b200_autospeed.h — C-ABI of the AutoSpeed detector in libvp_b200.so.
 *
 * What each entry point replaces in the reference (paths relative to the reference repo):
 *   vp_autospeed_create       AutoSpeedNetworkInfer.__init__ Models/inference/auto_speed_infer.py:6-14 (model load);
 *                             AutoSpeedNetwork.build_model('n', 4) Models/model_components/auto_speed/auto_speed_network.py:63-67;
 *                             C++: AutoSpeedTensorRTEngine ctor VisionPilot/production_release/src/inference/autospeed/tensorrt_engine.cpp
 *   vp_autospeed_infer        AutoSpeedNetworkInfer.inference auto_speed_infer.py:88-108: letterbox to 1024x512 (Pillow BILINEAR,
 *                             gray 114 padding :24-45) -> ToTensor (:50) -> YOLO forward (auto_speed_network.py:46-49) ->
 *                             second sigmoid + confidence 0.6 (:78-80) -> xywh to xyxy (:55-62) -> NMS 0.45 (:64-69) ->
 *                             un-letterbox + clamp (:100-106)
 *   vp_autospeed_detections   the list the helper returns: [[x1, y1, x2, y2, score, class], ...] in source-frame pixels
 *   vp_autospeed_raw          the network's raw prediction tensor [1, 4 + nc, 10752] (auto_speed_head.py:63)
 *
 * The checkpoint is a .vpw file holding the module's state_dict (python -m autoware_vision_pilot_b200.convert);
 * BatchNorm (eps 1e-3) is folded at load.  16-bit operands on the tcgen05 tensor cores, fp32 accumulation, exactly
 * like the reference helper's own half-precision inference (auto_speed_infer.py:50 `.half()`).
 * No CPU fallback: creation fails with VPB_ERR_CUDA without an sm_100 device.
 */
#ifndef VP_B200_AUTOSPEED_H_
#define VP_B200_AUTOSPEED_H_
#include <stddef.h>
#include <stdint.h>
#include "vp_b200_ops.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vp_autospeed vp_autospeed;

int vp_autospeed_create(const char* weights_vpw, int gpu_id, int dtype /* VPB_F16 | VPB_BF16 */, void* stream,
                        vp_autospeed** out);
void vp_autospeed_destroy(vp_autospeed* e);
/* conf_thres / iou_thres of post_process_predictions (defaults 0.6 / 0.45, auto_speed_infer.py:71) */
int vp_autospeed_set_thresholds(vp_autospeed* e, float conf, float iou);

/* Host RGB frame (uint8, 3 interleaved channels, any size), results on the host when it returns:
 * H2D + letterbox + network + decode + NMS + D2H + sync.  fetch_raw != 0 also copies the raw tensor. */
int vp_autospeed_infer(vp_autospeed* e, const uint8_t* frame_host_rgb, int h, int w, int stride, int fetch_raw);
/* Same work for a frame already in device memory, only enqueued on the engine's stream. */
int vp_autospeed_infer_device(vp_autospeed* e, const uint8_t* frame_dev_rgb, int h, int w, int stride);
/* Drain the stream; fetch: 0 nothing, 1 detections, 2 detections + raw tensor to the host buffers. */
int vp_autospeed_sync(vp_autospeed* e, int fetch);

/* det: engine-owned host buffer [n][6] = x1, y1, x2, y2, score, class (descending score, as torchvision.ops.nms
 * orders them); n_candidates (optional) = anchors that passed the confidence filter. Valid until the next inference. */
int vp_autospeed_detections(vp_autospeed* e, const float** det, int* n, int* n_candidates);
/* raw prediction tensor, fp32 planar [channels = 8][anchors = 10752]: cx, cy, w, h (canvas pixels), 4 class scores */
int vp_autospeed_raw(vp_autospeed* e, const float** raw_host, const float** raw_dev, int* channels, int* anchors);
int vp_autospeed_stats(vp_autospeed* e, int* n_launches, double* flops);
/* intermediate tensors for the parity tests ("canvas", "p1".."p5", "p5_ctx", "p5_sppf", "n3".."n5", "head0".."head2")
 * as fp32 NCHW; returns the element count (dst == NULL: size query) */
long vp_autospeed_read_tap(vp_autospeed* e, const char* name, float* dst, long cap, int* c, int* h, int* w);

#ifdef __cplusplus
}
#endif
#endif /* VP_B200_AUTOSPEED_H_ */
