/* vp_b200_ops.h — op-level C-ABI of libvp_b200.so (device pointers in, device pointers out).
 *
 * These are the individual sm_100a kernels the engine (vp_b200.h) strings together.
 * They are exported so that the parity tests can check every stage against the
 * oracle in isolation, and so that a host written in another language can build its
 * own graph.  All pointers are DEVICE pointers unless a name ends in _host.
 * All functions return 0 on success, <0 on error (vpb_last_error() has the text).
 *
 * Activation tensors are NHWC 16-bit (fp16 or bf16, chosen by `dtype`), channel
 * stride `ld*` a multiple of 8 elements, base address 16-byte aligned.
 * Weight tensors for the GEMM convolutions are [taps][Cout][Cin] 16-bit (K-major).
 *
 * Each entry cites the reference op it replaces (paths relative to the reference repo).
 */
#ifndef VP_B200_OPS_H_
#define VP_B200_OPS_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { VPB_OK = 0, VPB_ERR_ARG = -1, VPB_ERR_CUDA = -2, VPB_ERR_STATE = -3, VPB_ERR_IO = -4 };
enum { VPB_F16 = 0, VPB_BF16 = 1 };
enum { VPB_ACT_NONE = 0, VPB_ACT_GELU = 1, VPB_ACT_SILU = 2, VPB_ACT_SIGMOID = 3 };
/* conv epilogue modes */
enum {
  VPB_EPI_STORE = 0,  /* out = act(acc + bias)                                   */
  VPB_EPI_ADD = 1,    /* out = act(acc + bias) + res          (MBConv residual, ConvT + skip) */
  VPB_EPI_MULADD = 2, /* out = act(acc + bias) * res + res    (scene_context.py:56)           */
  VPB_EPI_FINAL = 3   /* fp32 planar logits + uint8 class map (heads' last conv + P8 post)    */
};
/* class-map rule for VPB_EPI_FINAL */
enum {
  VPB_FINAL_NONE = 0,     /* raw tensor only (Scene3D depth, scene_3d_infer.py:54-56)          */
  VPB_FINAL_ARGMAX = 1,   /* first-max-wins argmax (scene_seg_infer.py:52-55)                  */
  VPB_FINAL_THRESH = 2,   /* v > 0 ? 1 : 0 on channel 0 (domain_seg_infer.py:57-58)            */
  VPB_FINAL_EGOLANES = 3  /* other>right>left priority id {2,1,0,255}
                             (cuda_visualization_kernels.cu:45-75)                             */
};

const char* vpb_last_error(void);
void vpb_set_error(const char* fmt, ...);

/* Implicit-GEMM convolution on tcgen05 tensor cores.
 *   taps = 9 : Conv2d 3x3 stride 1 pad 1   (scene_neck.py:13-24, scene_seg_head.py:13-19, ...)
 *   taps = 1 : Conv2d 1x1                  (skip links scene_neck.py:12, EfficientNet pointwise)
 *   phases = 4 (taps must be 1): ConvTranspose2d k2 s2 (scene_neck.py:11); phase p=(a*2+b)
 *              writes output pixel (2h+a, 2w+b); out/res then have spatial size 2H x 2W.
 *   w      : [taps*phases][Cout][Cin] 16-bit,   bias: fp32 [Cout] or NULL
 *   in     : [H][W][ldi]  (Cin valid channels, ldi >= Cin, both multiples of 8)
 *   out    : [Ho][Wo][ldo], res (modes ADD/MULADD): [Ho][Wo][ldr]
 *   FINAL  : out_f32 planar [Cout][H][W] fp32, out_cls [H][W] uint8 (may be NULL)
 */
typedef struct {
  int dtype;
  int H, W, Cin, ldi;
  int Cout, taps, phases;
  int act, mode, final_kind;
  const void* in;
  const void* w;
  const float* bias;
  void* out;
  int ldo;
  const void* res;
  int ldr;
  float* out_f32;
  uint8_t* out_cls;
  int bn; /* 0 = auto */
} vpb_conv_args;
int vpb_conv_gemm(const vpb_conv_args* a, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VP_B200_OPS_H_ */
