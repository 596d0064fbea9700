// conv_gemm.cu — implicit-GEMM convolutions on Blackwell tcgen05 tensor cores.
//
// Replaces (reference, all fp32 library calls):
//   nn.Conv2d 3x3 s1 p1   Models/model_components/scene_neck.py:13-24, scene_seg_head.py:13-19,
//                         scene_3d_head.py:13-20, ego_lanes_head.py:13-15, scene_context.py:20-22
//   nn.Conv2d 1x1         scene_neck.py:12,17,22 (skip links) and EfficientNet-B0 pointwise convs
//   nn.ConvTranspose2d k2 s2   scene_neck.py:11,16,21, scene_seg_head.py:11,16
//
// Six main loops share one epilogue family.
//
// (1) conv_gemm_kernel — "tile" formulation, any of the three op types.
//     Activations are NHWC 16-bit.  For an output tile of 128 pixels (a TH x TW patch) and BN
//     output channels,  D[pixel, n] = sum_tap sum_c In[pixel + off(tap), c] * W[tap][n][c]  is a GEMM
//     with M = 128, N = BN, K = taps*Cin.  The A operand for one (tap, 64-channel chunk) is one 4-D
//     TMA box {64 ch, TW, TH, 1} at the shifted coordinate; out-of-range rows/columns are
//     zero-filled by the TMA unit, which IS the convolution's zero padding.
//     ConvTranspose = 4 phase GEMMs (output pixel (2h+a, 2w+b)); the neck's skip link
//     ConvT(in) + Conv1x1(skip) is ONE GEMM: the skip tensor is a second K segment read at output
//     resolution through a 5-D TMA view [h][a][w][b][c] into the same accumulator.
//
// (2) conv3x3_lin_kernel — "linear padded" formulation for the 3x3 layers that carry > 90 % of the
//     FLOPs.  Input and output live in HBM as zero-bordered images [(H+2)*(W+2)][C]; the GEMM M
//     index is the linear padded pixel index p, so tap (dy,dx) is the constant shift
//     (dy-1)*(W+2)+(dx-1).  One TMA box of 130 consecutive pixels per (chunk, dy) serves the three
//     dx taps: the MMA's A descriptor simply starts dx rows (dx*128 B) further into the same
//     shared-memory segment.  Measured on B200 (scripts/check_base_offset.py): the 128-B swizzle
//     phase is taken from the absolute shared-memory address, so the descriptor's base_offset
//     field must stay 0 for these row-shifted views (setting it to dx gives wrong results).
//     A traffic from L2 drops 9 -> 3.05 loads per chunk and TMA issue count halves; B (weights)
//     stream through their own ring.  Border pixels are written as zeros, so the output is again
//     a valid zero-bordered image for the next layer.  Used for 3x3 layers with < 96 tiles.
//
// (3) conv3x3_pair_kernel — formulation (2) on a CTA pair (cluster of 2, tcgen05.mma.cta_group::2,
//     M = 256): each CTA stages its own pixels and HALF of every weight tile.  The dominant kernel
//     (every 3x3 layer with >= 96 tiles); see the comment above the kernel and
//     profiles/r1_smem_operand_model.md for why halving the weight operand per SM is what counts.
//
// (4) conv3x3_splitk_kernel — formulation (2) with the K chunks of one output tile divided over a
//     cluster of 2-4 CTAs; fp32 partials are reduced through distributed shared memory in rank order
//     (deterministic).  Used for the 10x20 context layers (<= 32 output tiles, 36-72 K groups).
//
// (5) convt_ws_kernel — weight-stationary ConvTranspose (+ skip) on a CTA pair; since (6) only on the layer-by-layer
//     graphs (VPB_UPCONV=0, split-fp16 mode falls back to (1)).
//
// (6) upconv_pair_kernel — ConvTranspose2d(k2,s2) [+ Conv1x1 skip] and the Conv3x3 + GELU that follows, composed into
//     ONE GEMM over the low-resolution tensor (weights from upconv_compose.cu): four 2x2 taps per output phase + nine
//     skip taps, CTA pair, M = 256.  Dominant by device time in the 16-bit engines.
//
// All: operands land in shared memory in the 128-byte-swizzled K-major layout tcgen05.mma reads
// through descriptors; fp32 accumulators in TMEM, two of them, so the epilogue of tile i overlaps
// the main loop of tile i+1; persistent grid; warp-specialised:
//   warp 0       TMA producer (one elected lane)
//   warp 1       tcgen05.mma issuer (one elected lane)
//   warps 2..17  epilogue (4 per TMEM lane quadrant, splitting the 16-column chunks):
//                tcgen05.ld -> +bias (staged in smem) -> activation -> residual -> 16-bit NHWC
//                stores (or fp32 planar logits + class map for the heads' last conv).
//                The first ncu capture (profiles/r1_conv_v1_ncu.md) showed the kernel bound by this
//                stage, hence 16 warps, a branch-free GELU and no per-element global loads.
#include "common.cuh"
#include "conv_gemm.cuh"
#include "ops_internal.h"
#include <cstdio>
#include <cstring>
#include <algorithm>

namespace vpb {

static constexpr int kEpiWarps = 16;
static constexpr int kThreads = 64 + kEpiWarps * 32;
static constexpr int kMaxStages = 8;
static constexpr int kUpcStages = 12;       // upconv_pair_kernel ring (stages of 20-32 KB)
static constexpr int kATileBytes = 128 * 128;  // 128 pixels x 64 ch x 2 B
static constexpr int kAccStride = 256;         // TMEM columns between the two accumulators
// 227 KB opt-in limit covers static + dynamic shared memory; keep 4 KB for the static part.
static constexpr int kMaxDynSmem = 227 * 1024 - 4096;
// linear kernel
static constexpr int kSegRows = 130;                 // 128 pixels + one halo pixel each side
static constexpr int kSegBytes = 17 * 1024;          // slot size (130*128 = 16640 B used)
// slot size with ms M sub-tiles: (128*ms + 2) rows of 128 B, rounded up to 1 KB
__host__ __device__ constexpr uint32_t seg_slot_bytes(int ms) { return ((128u * ms + 2u) * 128u + 1023u) & ~1023u; }
static constexpr int kMaxRing = 16;

// ------------------------------------------------------------------------------------------------
// Shared epilogue: one accumulator (128 rows x BN fp32 columns in TMEM) -> outputs.
// ------------------------------------------------------------------------------------------------
struct EpiPix {
  bool ok;        // compute and store this row
  bool zero;      // store zeros instead (border pixel of a padded output)
  uint32_t ooff;  // element offset of the pixel in the output tensor   (pixel * ldo; < 2^31, checked by the plan)
  uint32_t roff;  // element offset of the pixel in the residual tensor (pixel * ldr)
  uint32_t fpix;  // pixel index in the planar fp32 / class outputs (FINAL)
};

// NC = 16-column chunks handled per loop iteration.  The epilogue warps are latency-bound (4 warps per
// scheduler, one dependent chain each: ncu shows ~16 cycles per issued instruction), so with NC = 2 both
// TMEM loads are issued before the single wait and every later phase works on two independent chunks
// the scheduler can interleave.  Measured (gpurun_out/bench_conv_v14_nc{1,2}.txt): NC = 2 is 2-3 % SLOWER —
// 18 warps leave 96 registers per thread (5 warps on two of the SM sub-partitions) and the second chunk
// spills; only NC = 1 is instantiated.
template <class E, int NC, bool TILEK>
__device__ __forceinline__ void epilogue_chunks(const ConvKParams& p, uint32_t t_row, int n0,
                                                const float* sbias, int part, const EpiPix& px) {
  const bool split = TILEK && p.split;       // split-fp16 mode and the post-residual activation exist in the tile kernel only
  const int nchunks = p.BN >> 4;
  typename E::T* out = reinterpret_cast<typename E::T*>(p.out);
  const typename E::T* res = reinterpret_cast<const typename E::T*>(p.res);
  typename E::T* out_lo = reinterpret_cast<typename E::T*>(p.out_lo);                // split-fp16 mode only
  const typename E::T* res_lo = reinterpret_cast<const typename E::T*>(p.res_lo);
  for (int chunk0 = part; chunk0 < nchunks; chunk0 += 4 * NC) {
    uint32_t rr[NC][16];
#pragma unroll
    for (int c = 0; c < NC; ++c) tmem_ld16(t_row + (chunk0 + 4 * c) * 16, rr[c]);
    tmem_ld_wait();
    float v[NC][16];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const float4* sb4 = reinterpret_cast<const float4*>(sbias + (chunk0 + 4 * c) * 16);   // 4 x LDS.128
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 b4 = sb4[i];
        const float2 lo = fadd2(make_float2(__uint_as_float(rr[c][4 * i]), __uint_as_float(rr[c][4 * i + 1])), make_float2(b4.x, b4.y));
        const float2 hi = fadd2(make_float2(__uint_as_float(rr[c][4 * i + 2]), __uint_as_float(rr[c][4 * i + 3])), make_float2(b4.z, b4.w));
        v[c][4 * i] = lo.x; v[c][4 * i + 1] = lo.y; v[c][4 * i + 2] = hi.x; v[c][4 * i + 3] = hi.y;
      }
    }
    // activation switch hoisted out of the element loop (act(0) == 0 for GELU/SiLU keeps the
    // channel padding zero; sigmoid is masked explicitly)
    if (p.act == ACT_GELU) {
#pragma unroll
      for (int c = 0; c < NC; ++c)
#pragma unroll
        for (int i = 0; i < 8; ++i) {   // packed fp32x2: two elements per FFMA2 / FMUL2
          const float2 g = act_gelu2(make_float2(v[c][2 * i], v[c][2 * i + 1]));
          v[c][2 * i] = g.x; v[c][2 * i + 1] = g.y;
        }
    } else if (p.act == ACT_SILU) {
#pragma unroll
      for (int c = 0; c < NC; ++c)
#pragma unroll
        for (int i = 0; i < 16; ++i) v[c][i] = act_silu(v[c][i]);
    } else if (p.act == ACT_SIGMOID) {
#pragma unroll
      for (int c = 0; c < NC; ++c)
#pragma unroll
        for (int i = 0; i < 16; ++i) v[c][i] = (n0 + (chunk0 + 4 * c) * 16 + i < p.Cout) ? act_sigmoid(v[c][i]) : 0.f;
    }
    if (p.mode == VPB_EPI_FINAL) {
      if (px.ok && chunk0 == 0) {
        const uint32_t plane = static_cast<uint32_t>(p.H * p.W);
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (i < p.Cout) p.out_f32[i * plane + px.fpix] = v[0][i];
        if (p.out_cls) {
          uint8_t cls = 0;
          if (p.final_kind == VPB_FINAL_ARGMAX) {
            float best = v[0][0];
#pragma unroll
            for (int i = 1; i < 16; ++i)
              if (i < p.Cout && v[0][i] > best) { best = v[0][i]; cls = static_cast<uint8_t>(i); }
          } else if (p.final_kind == VPB_FINAL_THRESH) {
            cls = v[0][0] > 0.f ? 1 : 0;
          } else if (p.final_kind == VPB_FINAL_EGOLANES) {
            cls = (v[0][2] > 0.f) ? 2 : (v[0][1] > 0.f) ? 1 : (v[0][0] > 0.f) ? 0 : 255;
          }
          p.out_cls[px.fpix] = cls;
        }
      }
    } else if (px.ok) {
      if (p.mode == VPB_EPI_ADD || p.mode == VPB_EPI_MULADD) {
        uint4 rv[NC][2];
        bool rok[NC][2];
#pragma unroll
        for (int c = 0; c < NC; ++c) {          // all residual loads in flight before the first use
          const int n = n0 + (chunk0 + 4 * c) * 16;
          const typename E::T* rp = res + (px.roff + n);
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            rok[c][j] = (n + 8 * j < p.ldr) && (n + 8 * j < p.nlim);
            rv[c][j] = rok[c][j] ? *reinterpret_cast<const uint4*>(rp + 8 * j) : make_uint4(0, 0, 0, 0);
          }
        }
        uint4 rl[NC][2];
#pragma unroll
        for (int c = 0; c < NC; ++c)
#pragma unroll
          for (int j = 0; j < 2; ++j)
            rl[c][j] = (split && rok[c][j])
                           ? *reinterpret_cast<const uint4*>(res_lo + (px.roff + n0 + (chunk0 + 4 * c) * 16) + 8 * j)
                           : make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int c = 0; c < NC; ++c)
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            if (!rok[c][j]) continue;
            const uint32_t rw[4] = {rv[c][j].x, rv[c][j].y, rv[c][j].z, rv[c][j].w};
            const uint32_t rwl[4] = {rl[c][j].x, rl[c][j].y, rl[c][j].z, rl[c][j].w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              float2 f = unpack2<E>(rw[i]);
              if (split) { const float2 fl = unpack2<E>(rwl[i]); f.x += fl.x; f.y += fl.y; }
              float& a = v[c][8 * j + 2 * i];
              float& b = v[c][8 * j + 2 * i + 1];
              if (p.mode == VPB_EPI_ADD) { a += f.x; b += f.y; }
              else { a = fmaf(a, f.x, f.x); b = fmaf(b, f.y, f.y); }
              if (TILEK && p.act2 == ACT_SILU) { a = act_silu(a); b = act_silu(b); }
            }
          }
      }
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const int n = n0 + (chunk0 + 4 * c) * 16;
        typename E::T* op = out + (px.ooff + n);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          if (n + 8 * j < p.nlim) {
            uint4 o;
            o.x = pack2<E>(v[c][8 * j + 0], v[c][8 * j + 1]);
            o.y = pack2<E>(v[c][8 * j + 2], v[c][8 * j + 3]);
            o.z = pack2<E>(v[c][8 * j + 4], v[c][8 * j + 5]);
            o.w = pack2<E>(v[c][8 * j + 6], v[c][8 * j + 7]);
            *reinterpret_cast<uint4*>(op + 8 * j) = o;
            if (split) {          // low half: what the 16-bit rounding of the high half lost
              const uint32_t ow[4] = {o.x, o.y, o.z, o.w};
              uint32_t lw[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float2 h = unpack2<E>(ow[i]);
                lw[i] = pack2<E>(v[c][8 * j + 2 * i] - h.x, v[c][8 * j + 2 * i + 1] - h.y);
              }
              *reinterpret_cast<uint4*>(out_lo + (px.ooff + n) + 8 * j) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
            }
          }
        }
      }
    } else if (px.zero) {
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const int n = n0 + (chunk0 + 4 * c) * 16;
        typename E::T* op = out + (px.ooff + n);
#pragma unroll
        for (int j = 0; j < 2; ++j)
          if (n + 8 * j < p.nlim) *reinterpret_cast<uint4*>(op + 8 * j) = make_uint4(0, 0, 0, 0);
      }
    }
  }
}

// Lean form of the above for the two dominant cases — plain store with no activation (ConvTranspose + skip,
// 1x1 projections) or GELU (every 3x3 of context / neck / heads) — when the whole N tile lies inside the
// output row (no per-store predicates).  The epilogue is latency-bound on its instruction chain (ncu: ~16
// cycles per issued instruction, 77 instructions per chunk in the generic form, most of them uniform
// mode/activation tests and address arithmetic): here pointers advance by constants and nothing is tested
// inside the loop.
// ACT: ACT_NONE | ACT_GELU | ACT_SILU (the encoder's expand / head convolutions); ADD: residual added after the
// activation (MBConv projection + skip: out = conv + bias + res), the whole N tile inside the residual row as well.
// CLS / gb: as epilogue_to_smem — a border pixel of the upconv kernel reads its bias row from global memory.
template <class E, int ACT, bool ADD = false, bool CLS = false>
__device__ __forceinline__ void epilogue_store_fast(const ConvKParams& p, uint32_t t_row, int n0,
                                                    const float* sbias, int part, const EpiPix& px,
                                                    const float* gb = nullptr) {
  const int nchunks = p.BN >> 4;
  typename E::T* op = reinterpret_cast<typename E::T*>(p.out) + (px.ooff + n0 + part * 16);
  const typename E::T* rp = ADD ? reinterpret_cast<const typename E::T*>(p.res) + (px.roff + n0 + part * 16) : nullptr;
  const float4* sb4 = reinterpret_cast<const float4*>(sbias + part * 16);
  const float4* gb4 = (CLS && gb) ? reinterpret_cast<const float4*>(gb + part * 16) : nullptr;
  uint32_t ta = t_row + part * 16;
  const bool ok = px.ok, zero = px.zero;
  auto finish = [&](const uint32_t (&rr)[16], const float4* sb, typename E::T* o, const typename E::T* r, int goff = 0) {
    uint4 r0 = make_uint4(0, 0, 0, 0), r1 = r0;
    if (ADD && ok) { r0 = reinterpret_cast<const uint4*>(r)[0]; r1 = reinterpret_cast<const uint4*>(r)[1]; }   // in flight early
    float2 v[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float4 b4;
      if (CLS && gb4) b4 = __ldg(gb4 + goff + i);
      else b4 = sb[i];
      v[2 * i] = fadd2(make_float2(__uint_as_float(rr[4 * i]), __uint_as_float(rr[4 * i + 1])), make_float2(b4.x, b4.y));
      v[2 * i + 1] = fadd2(make_float2(__uint_as_float(rr[4 * i + 2]), __uint_as_float(rr[4 * i + 3])), make_float2(b4.z, b4.w));
    }
    if (ACT == ACT_GELU) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = act_gelu2(v[i]);
    } else if (ACT == ACT_SILU) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = make_float2(act_silu(v[i].x), act_silu(v[i].y));
    }
    if (ADD) {
      const uint32_t rw[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = fadd2(v[i], unpack2<E>(rw[i]));
    }
    uint4 o0, o1;
    o0.x = pack2<E>(v[0].x, v[0].y); o0.y = pack2<E>(v[1].x, v[1].y); o0.z = pack2<E>(v[2].x, v[2].y); o0.w = pack2<E>(v[3].x, v[3].y);
    o1.x = pack2<E>(v[4].x, v[4].y); o1.y = pack2<E>(v[5].x, v[5].y); o1.z = pack2<E>(v[6].x, v[6].y); o1.w = pack2<E>(v[7].x, v[7].y);
    if (zero) { o0 = make_uint4(0, 0, 0, 0); o1 = o0; }
    if (ok || zero) {
      reinterpret_cast<uint4*>(o)[0] = o0;
      reinterpret_cast<uint4*>(o)[1] = o1;
    }
  };
  int chunk = part;
  // two chunks per TMEM wait while at least two remain (the drain is bound by TMEM read latency / bandwidth,
  // profiles/r1_convt_timeline.md), then the odd one
  for (; chunk + 4 < nchunks; chunk += 8, op += 128, sb4 += 32, ta += 128) {
    uint32_t ra[16], rb[16];
    tmem_ld16(ta, ra);                     // .sync.aligned: every lane takes part, stores are predicated
    tmem_ld16(ta + 64, rb);
    tmem_ld_wait();
    finish(ra, sb4, op, rp);
    finish(rb, sb4 + 16, op + 64, rp + 64, 16);
    if (ADD) rp += 128;
    if (CLS && gb4) gb4 += 32;
  }
  if (chunk < nchunks) {
    uint32_t ra[16];
    tmem_ld16(ta, ra);
    tmem_ld_wait();
    finish(ra, sb4, op, rp);
  }
}

// Staged form of epilogue_store_fast (tile kernel, ConvTranspose / 1x1 with plain store or GELU): the accumulator
// row of a lane is ONE pixel, so direct stores are 32 separate 16-byte requests per instruction (a different 128-byte
// line per lane) — profiles/r1_convt_timeline.md measured 4.5-6.7 k cycles per 128 x 256 tile for that, the L1 store
// path's one-request-per-cycle limit, and it is what bounds the ConvTranspose layers (they write 4x the pixels they
// read).  Here every 16-column chunk goes to shared memory in the 128-byte-swizzled box layout ([128 pixels][64 ch]
// per slab: 4 wavefronts per warp store, the minimum) and one thread then issues a TMA store per 64-channel slab.
// chunk0 / nck >= 0: this warp drains the nck CONSECUTIVE chunks starting at chunk0 (slab-group form used by
// convt_ws_kernel, where the four quadrant warps of a slab run their own barrier and store); nck < 0: the interleaved
// assignment chunk = part, part + 4, ... of the tile kernel.
// CLS / gb: the upconv kernel's border pixels take their bias row from global memory (gb != nullptr: this lane's pixel is on
// the image border, where the folded ConvTranspose bias sees fewer 3x3 taps) instead of the staged interior row.
template <class E, bool GELU, bool CLS = false>
__device__ __forceinline__ void epilogue_to_smem(const ConvKParams& p, uint32_t t_row, const float* sbias, int part,
                                                 uint32_t slab0, int row, int chunk0 = 0, int nck = -1,
                                                 const float* gb = nullptr) {
  const int nchunks = nck >= 0 ? chunk0 + nck : (p.BN >> 4);
  const uint32_t rbase = slab0 + static_cast<uint32_t>(row) * 128u;
  const uint32_t rx = static_cast<uint32_t>(row & 7);
  auto finish = [&](const uint32_t (&rr)[16], int chunk) {
    const float4* sb = reinterpret_cast<const float4*>(sbias + chunk * 16);
    float2 v[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float4 b4;
      if (CLS && gb) b4 = __ldg(reinterpret_cast<const float4*>(gb + chunk * 16) + i);
      else b4 = sb[i];
      v[2 * i] = fadd2(make_float2(__uint_as_float(rr[4 * i]), __uint_as_float(rr[4 * i + 1])), make_float2(b4.x, b4.y));
      v[2 * i + 1] = fadd2(make_float2(__uint_as_float(rr[4 * i + 2]), __uint_as_float(rr[4 * i + 3])), make_float2(b4.z, b4.w));
    }
    if (GELU) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = act_gelu2(v[i]);
    }
    uint4 o0, o1;
    o0.x = pack2<E>(v[0].x, v[0].y); o0.y = pack2<E>(v[1].x, v[1].y); o0.z = pack2<E>(v[2].x, v[2].y); o0.w = pack2<E>(v[3].x, v[3].y);
    o1.x = pack2<E>(v[4].x, v[4].y); o1.y = pack2<E>(v[5].x, v[5].y); o1.z = pack2<E>(v[6].x, v[6].y); o1.w = pack2<E>(v[7].x, v[7].y);
    const uint32_t slab = rbase + static_cast<uint32_t>(chunk >> 2) * (128u * 128u);
    const uint32_t j = static_cast<uint32_t>(chunk & 3) * 2u;          // 16-byte piece index inside the 128-byte row
    st_shared_v4(slab + ((j ^ rx) << 4), o0);
    st_shared_v4(slab + (((j + 1u) ^ rx) << 4), o1);
  };
  const int step = nck >= 0 ? 1 : 4;
  int chunk = nck >= 0 ? chunk0 : part;
  for (; chunk + step < nchunks; chunk += 2 * step) {
    uint32_t ra[16], rb[16];
    tmem_ld16(t_row + chunk * 16, ra);
    tmem_ld16(t_row + (chunk + step) * 16, rb);
    tmem_ld_wait();
    finish(ra, chunk);
    finish(rb, chunk + step);
  }
  if (chunk < nchunks) {
    uint32_t ra[16];
    tmem_ld16(t_row + chunk * 16, ra);
    tmem_ld_wait();
    finish(ra, chunk);
  }
}

// TILEK: the tile kernel (encoder 1x1 convolutions) also gets the lean SiLU and residual-add forms; the 3x3 kernels do
// not instantiate them (they never use those modes and the extra paths cost them registers / spills).
template <class E, bool TILEK = false>
__device__ __forceinline__ void epilogue_tile(const ConvKParams& p, uint32_t t_row, int n0,
                                              const float* sbias, int part, const EpiPix& px) {
  // NC = 2 (two chunks per iteration) measured 2-3 % slower and makes ptxas spill in every kernel that
  // contains it (96-register cap), so only the one-chunk form is instantiated.
  const bool whole = !p.split && n0 + p.BN <= p.nlim && p.act2 == ACT_NONE;
  if (whole && p.mode == VPB_EPI_STORE && p.act == ACT_GELU) epilogue_store_fast<E, ACT_GELU>(p, t_row, n0, sbias, part, px);
  else if (whole && p.mode == VPB_EPI_STORE && p.act == ACT_NONE) epilogue_store_fast<E, ACT_NONE>(p, t_row, n0, sbias, part, px);
  else if (TILEK && whole && p.mode == VPB_EPI_STORE && p.act == ACT_SILU) epilogue_store_fast<E, ACT_SILU>(p, t_row, n0, sbias, part, px);
  else if (TILEK && whole && p.mode == VPB_EPI_ADD && p.act == ACT_NONE && n0 + p.BN <= p.ldr) epilogue_store_fast<E, ACT_NONE, true>(p, t_row, n0, sbias, part, px);
  else epilogue_chunks<E, 1, TILEK>(p, t_row, n0, sbias, part, px);
}

__device__ __forceinline__ void stage_bias(const ConvKParams& p, float* dst, int etid, int n0) {
  if (etid < p.BN) {
    const int nn = n0 + etid;
    dst[etid] = (p.bias && nn < p.Cout) ? __ldg(p.bias + nn) : 0.f;
  }
  asm volatile("bar.sync 1, 512;" ::: "memory");
}

// ------------------------------------------------------------------------------------------------
// (1) tile formulation
// ------------------------------------------------------------------------------------------------
template <class E>
__global__ void __launch_bounds__(kThreads, 1)
conv_gemm_kernel(const __grid_constant__ ConvMaps maps, const ConvKParams p) {
  const CUtensorMap& mapA = maps.A;
  const CUtensorMap& mapB = maps.B;
  const CUtensorMap& mapA2 = maps.A2;
  const CUtensorMap& mapB2 = maps.B2;
  const CUtensorMap& mapO = maps.O;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[kMaxStages];
  __shared__ __align__(8) uint64_t bar_empty[kMaxStages];
  __shared__ __align__(8) uint64_t bar_tfull[2];
  __shared__ __align__(8) uint64_t bar_tempty[2];
  __shared__ uint32_t tmem_holder;
  __shared__ __align__(16) float s_bias[2][256];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t b_tile_bytes = static_cast<uint32_t>(p.BN) * 128u;
  // fuse4 (ConvTranspose): one stage = the A tile + the weight tiles of all four phases, so the
  // activations are fetched once per K chunk and the four phase accumulators (4 x 128 TMEM
  // columns) are filled together
  const int nb_tiles = p.fuse4 ? 4 : 1;
  const uint32_t stage_bytes = kATileBytes + nb_tiles * b_tile_bytes;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapA);
    tma_prefetch_desc(&mapB);
    if (p.kchunks2) { tma_prefetch_desc(&mapA2); tma_prefetch_desc(&mapB2); }
    if (p.tma_store) tma_prefetch_desc(&mapO);
    if (p.split) { tma_prefetch_desc(&maps.Alo); tma_prefetch_desc(&maps.Blo); }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&bar_tfull[s]), 1);
      mbar_init(smem_u32(&bar_tempty[s]), kEpiWarps);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(smem_u32(&tmem_holder), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_holder;
  pdl_launch_dependents();   // let the next kernel's prologue overlap this kernel
  pdl_wait();                // predecessor's outputs (our inputs) are complete and visible from here on
  if (p.trace && blockIdx.x == 0 && threadIdx.x == 0) p.trace[255] = clock64();

  // split-fp16 mode: every K chunk is walked three times (A_hi W_hi, A_lo W_hi, A_hi W_lo) into the same accumulator
  const int nseg = p.split ? 3 : 1;
  const int kiters = p.taps * p.kchunks;
  const int kiters_all = (kiters + p.kchunks2) * nseg;   // + the fused skip link's K chunks
  const int tiles_per_phase = p.tiles_n * p.tiles_h * p.tiles_w;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    // The whole warp runs the loop converged and one ELECTed lane issues: that lets the compiler
    // emit the uniform-datapath instructions (UTMALDG / UTCHMMA / UTCBAR) directly instead of the
    // per-instruction BRA.U.ANY loops it needs inside an `if (lane == 0)` region.
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int ph = static_cast<int>(fast_div(tile, p.mg_tpp));
      int r = tile - ph * tiles_per_phase;
      const int rn = static_cast<int>(fast_div(r, p.mg_tn));
      const int nt = r - rn * p.tiles_n;
      const int thi = static_cast<int>(fast_div(rn, p.mg_tw));
      const int twi = rn - thi * p.tiles_w;
      const int h0 = thi * p.TH, w0 = twi * p.TW, n0 = nt * p.BN;
      for (int t = 0; t < p.taps; ++t) {
        const int dy = (p.taps == 9) ? (t / 3 - 1) : 0;
        const int dx = (p.taps == 9) ? (t % 3 - 1) : 0;
        const int wsel = (p.phases > 1) ? ph : t;
        for (int c = 0; c < p.kchunks; ++c) {
          for (int seg = 0; seg < nseg; ++seg) {
            const CUtensorMap* mA = seg == 1 ? &maps.Alo : &mapA;
            const CUtensorMap* mB = seg == 2 ? &maps.Blo : &mapB;
            mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1u);
            if (elect_one()) {
              const uint32_t full = smem_u32(&bar_full[stage]);
              const uint32_t sa = smem_base + stage * stage_bytes;
              if (p.trace && blockIdx.x == 0) {
                const int ti = (tile - blockIdx.x) / gridDim.x, k = t * p.kchunks + c;
                if (ti < 16 && k < 4) p.trace[ti * 16 + k] = clock64();
              }
              mbar_arrive_expect_tx(full, stage_bytes);
              tma_load_4d(sa, mA, full, c * 64, w0 * p.stride + dx, h0 * p.stride + dy, 0);
              if (p.fuse4) {
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4)
                  tma_load_3d(sa + kATileBytes + q4 * b_tile_bytes, mB, full, c * 64, n0, q4);
              } else {
                tma_load_3d(sa + kATileBytes, mB, full, c * 64, n0, wsel);
              }
            }
            __syncwarp();
            if (++stage == p.stages) { stage = 0; phase ^= 1u; }
          }
        }
      }
      // fused skip link: the same output pixels seen in the second input (at output resolution);
      // for a ConvTranspose phase (a,b) that is the pixel set (2h+a, 2w+b), addressed through a
      // 5-D view [h][a][w][b][c] of the tensor so a plain tiled box picks every second pixel
      for (int c2 = 0; c2 < p.kchunks2; ++c2) {
        for (int seg = 0; seg < nseg; ++seg) {
          const CUtensorMap* mA = seg == 1 ? &maps.A2lo : &mapA2;
          const CUtensorMap* mB = seg == 2 ? &maps.B2lo : &mapB2;
          mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1u);
          if (elect_one()) {
            const uint32_t full = smem_u32(&bar_full[stage]);
            const uint32_t sa = smem_base + stage * stage_bytes;
            mbar_arrive_expect_tx(full, stage_bytes);
            tma_load_5d(sa, mA, full, c2 * 64, ph & 1, w0, ph >> 1, h0);
            tma_load_3d(sa + kATileBytes, mB, full, c2 * 64, n0, 0);
          }
          __syncwarp();
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (converged warp, elected lane)
    const uint32_t idesc = umma_idesc(E::kUmmaFmt, 128, p.BN);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      // fuse4 uses all 512 TMEM columns for one tile (single accumulator stage)
      const int as = p.fuse4 ? 0 : (it & 1);
      const uint32_t aphase = p.fuse4 ? (it & 1) : ((it >> 1) & 1);
      mbar_wait(smem_u32(&bar_tempty[as]), aphase ^ 1u);
      tc_fence_after();
      if (p.trace && blockIdx.x == 0 && it < 16 && lane == 0) p.trace[it * 16 + 9] = clock64();
      const uint32_t d_tmem = tmem_base + as * kAccStride;
      for (int k = 0; k < kiters_all; ++k) {
        const int kl = p.split ? static_cast<int>(fast_div(k, 0x55555556u)) : k;   // logical K chunk (k / 3 in split mode)
        const int c = kl % p.kchunks;
        const int kvalid = kl < kiters ? min(64, p.Cin - c * 64) : min(64, p.Cin2 - (kl - kiters) * 64);
        const int ksteps = (kvalid + 15) >> 4;
        mbar_wait(smem_u32(&bar_full[stage]), phase);
        tc_fence_after();
        if (elect_one()) {
          if (p.trace && blockIdx.x == 0 && it < 16) {
            if (k < 4) p.trace[it * 16 + 4 + k] = clock64();
            if (k == kiters_all - 1) p.trace[it * 16 + 8] = clock64();
          }
          const uint32_t sa = smem_base + stage * stage_bytes;
          const uint64_t adesc = umma_desc_k128(sa);
          for (int q4 = 0; q4 < nb_tiles; ++q4) {
            const uint64_t bdesc = umma_desc_k128(sa + kATileBytes + q4 * b_tile_bytes);
            const uint32_t dt = d_tmem + q4 * 128;
            // +32 B along K inside the 128-B swizzle row == +2 in the encoded start address
            if (ksteps == 4) {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                umma_f16(dt, adesc + 2 * kk, bdesc + 2 * kk, idesc, (k | kk) != 0 ? 1u : 0u);
            } else {
              for (int kk = 0; kk < ksteps; ++kk)
                umma_f16(dt, adesc + 2 * kk, bdesc + 2 * kk, idesc, (k | kk) != 0 ? 1u : 0u);
            }
          }
          umma_commit(smem_u32(&bar_empty[stage]));
          if (k == kiters_all - 1) umma_commit(smem_u32(&bar_tfull[as]));
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue
    const int q = warp & 3;            // TMEM lane quadrant this warp may read
    const int part = (warp - 2) >> 2;  // which interleaved set of 16-column chunks (0..3)
    const int etid = threadIdx.x - 64; // 0..511 among the epilogue threads
    const int row = q * 32 + lane;     // accumulator row == pixel within the tile
    const int lh = row >> p.tw_shift;
    const int lw = row & (p.TW - 1);
    const int Wo = (p.phases > 1) ? 2 * p.W : p.W;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const int as = p.fuse4 ? 0 : (it & 1);
      const uint32_t aphase = p.fuse4 ? (it & 1) : ((it >> 1) & 1);
      const int ph0 = p.fuse4 ? 0 : static_cast<int>(fast_div(tile, p.mg_tpp));
      const int r = tile - ph0 * tiles_per_phase;
      const int rn = static_cast<int>(fast_div(r, p.mg_tn));
      const int nt = r - rn * p.tiles_n;
      const int thi = static_cast<int>(fast_div(rn, p.mg_tw));
      const int twi = rn - thi * p.tiles_w;
      const int h = thi * p.TH + lh, w = twi * p.TW + lw, n0 = nt * p.BN;

      stage_bias(p, s_bias[as], etid, n0);
      if (p.trace && blockIdx.x == 0 && it < 16 && etid == 0) p.trace[it * 16 + 10] = clock64();
      mbar_wait(smem_u32(&bar_tfull[as]), aphase);
      tc_fence_after();
      if (p.trace && blockIdx.x == 0 && it < 16 && etid == 0) p.trace[it * 16 + 11] = clock64();
      if (p.tma_store) {
        // staged epilogue: accumulator -> swizzled shared-memory slabs -> one TMA store per 64-channel slab
        const uint32_t slab0 = smem_base + static_cast<uint32_t>(p.stages) * stage_bytes;
        const int h0 = thi * p.TH, w0 = twi * p.TW;
        for (int q4 = 0; q4 < nb_tiles; ++q4) {
          const int ph = ph0 + q4;
          if (etid == 0) bulk_wait_read0();                 // the previous tile's stores have read the slabs
          asm volatile("bar.sync 1, 512;" ::: "memory");
          const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * kAccStride + q4 * 128;
          if (p.act == ACT_GELU) epilogue_to_smem<E, true>(p, t_row, s_bias[as], part, slab0, row);
          else epilogue_to_smem<E, false>(p, t_row, s_bias[as], part, slab0, row);
          fence_proxy_async();                              // generic-proxy smem writes -> visible to the TMA unit
          if (q4 == nb_tiles - 1) tc_fence_before();
          asm volatile("bar.sync 1, 512;" ::: "memory");
          if (etid == 0) {
            for (int sl = 0; sl < (p.BN >> 6); ++sl)
              tma_store_5d(&mapO, slab0 + sl * (128 * 128), n0 + sl * 64, ph & 1, w0, ph >> 1, h0);
            bulk_commit();
          }
        }
        if (lane == 0) mbar_arrive(smem_u32(&bar_tempty[as]));    // accumulator drained (stores still in flight)
        continue;
      }
      for (int q4 = 0; q4 < nb_tiles; ++q4) {
        const int ph = ph0 + q4;
        const int oh = (p.phases > 1) ? 2 * h + (ph >> 1) : h;
        const int ow = (p.phases > 1) ? 2 * w + (ph & 1) : w;
        EpiPix px;
        px.ok = (h < p.H) && (w < p.W);
        px.zero = false;
        px.ooff = static_cast<uint32_t>((oh + p.out_pad) * (Wo + 2 * p.out_pad) + (ow + p.out_pad)) * p.ldo;
        px.roff = static_cast<uint32_t>((oh + p.res_pad) * (Wo + 2 * p.res_pad) + (ow + p.res_pad)) * p.ldr;
        px.fpix = static_cast<uint32_t>(h * p.W + w);
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * kAccStride + q4 * 128;
        epilogue_tile<E, true>(p, t_row, n0, s_bias[as], part, px);
      }
      tc_fence_before();
      __syncwarp();
      if (p.trace && blockIdx.x == 0 && it < 16 && etid == 0) p.trace[it * 16 + 12] = clock64();
      if (lane == 0) mbar_arrive(smem_u32(&bar_tempty[as]));
    }
  }

  if (p.tma_store && threadIdx.x == 64) bulk_wait0();      // this thread issued the stores: all of them have completed
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// (2) linear padded formulation (3x3 only)
// ------------------------------------------------------------------------------------------------
template <class E>
__global__ void __launch_bounds__(kThreads, 1)
conv3x3_lin_kernel(const __grid_constant__ CUtensorMap mapA,
                   const __grid_constant__ CUtensorMap mapB, const ConvKParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full[kMaxRing], a_empty[kMaxRing];
  __shared__ __align__(8) uint64_t b_full[kMaxRing], b_empty[kMaxRing];
  __shared__ __align__(8) uint64_t bar_tfull[2], bar_tempty[2];
  __shared__ uint32_t tmem_holder;
  __shared__ __align__(16) float s_bias[2][256];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t b_bytes = static_cast<uint32_t>(p.BN) * 128u;
  const uint32_t seg_bytes = seg_slot_bytes(p.ms);
  const uint32_t acc_cols = p.BN <= 64 ? 64u : 128u;   // TMEM columns between the M sub-tiles' accumulators
  const uint32_t b_base = smem_base + p.na * seg_bytes;
  const int tile_px = 128 * p.ms;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapA);
    tma_prefetch_desc(&mapB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.na; ++s) { mbar_init(smem_u32(&a_full[s]), 1); mbar_init(smem_u32(&a_empty[s]), 1); }
    for (int s = 0; s < p.nb; ++s) { mbar_init(smem_u32(&b_full[s]), 1); mbar_init(smem_u32(&b_empty[s]), 1); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&bar_tfull[s]), 1);
      mbar_init(smem_u32(&bar_tempty[s]), kEpiWarps);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(smem_u32(&tmem_holder), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_holder;
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (converged warp, elected lane)
    int sa = 0, sb = 0;
    uint32_t pa = 0, pb = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int mt = static_cast<int>(fast_div(tile, p.mg_tn)), nt = tile - mt * p.tiles_n;
      const int p0 = mt * tile_px, n0 = nt * p.BN;
      for (int c = 0; c < p.kchunks; ++c) {
        for (int dy = 0; dy < 3; ++dy) {
          mbar_wait(smem_u32(&a_empty[sa]), pa ^ 1u);
          if (elect_one()) {
            const uint32_t af = smem_u32(&a_full[sa]);
            mbar_arrive_expect_tx(af, kSegRows * 128 * p.ms);
            // 130 consecutive padded pixels starting one pixel left of the tile in row (dy-1);
            // with several M sub-tiles further 130-row boxes land 128 rows apart (the two overlap
            // rows are written twice with identical bytes) — TMA boxes are limited to 256 rows.
            const int r0 = p0 + (dy - 1) * p.WP - 1;
            for (int j = 0; j < p.ms; ++j)
              tma_load_2d(smem_base + sa * seg_bytes + j * 128 * 128, &mapA, af, c * 64, r0 + j * 128);
          }
          __syncwarp();
          if (++sa == p.na) { sa = 0; pa ^= 1u; }
          if (p.gb == 3) {
            // one weight stage = the three dx taps of this kernel row (one barrier handshake
            // per 12 MMAs instead of per 4)
            mbar_wait(smem_u32(&b_empty[sb]), pb ^ 1u);
            if (elect_one()) {
              const uint32_t bf = smem_u32(&b_full[sb]);
              mbar_arrive_expect_tx(bf, 3 * b_bytes);
#pragma unroll
              for (int dx = 0; dx < 3; ++dx)
                tma_load_3d(b_base + (sb * 3 + dx) * b_bytes, &mapB, bf, c * 64, n0, dy * 3 + dx);
            }
            __syncwarp();
            if (++sb == p.nb) { sb = 0; pb ^= 1u; }
          } else {
            for (int dx = 0; dx < 3; ++dx) {
              mbar_wait(smem_u32(&b_empty[sb]), pb ^ 1u);
              if (elect_one()) {
                const uint32_t bf = smem_u32(&b_full[sb]);
                mbar_arrive_expect_tx(bf, b_bytes);
                tma_load_3d(b_base + sb * b_bytes, &mapB, bf, c * 64, n0, dy * 3 + dx);
              }
              __syncwarp();
              if (++sb == p.nb) { sb = 0; pb ^= 1u; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (converged warp, elected lane)
    const uint32_t idesc = umma_idesc(E::kUmmaFmt, 128, p.BN);
    int sa = 0, sb = 0;
    uint32_t pa = 0, pb = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      mbar_wait(smem_u32(&bar_tempty[as]), aphase ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * kAccStride;
      uint32_t first = 1;
      for (int c = 0; c < p.kchunks; ++c) {
        const int kvalid = min(64, p.Cin - c * 64);
        const int ksteps = (kvalid + 15) >> 4;
        for (int dy = 0; dy < 3; ++dy) {
          const bool last = (c == p.kchunks - 1) && (dy == 2);
          mbar_wait(smem_u32(&a_full[sa]), pa);
          const uint32_t seg = smem_base + sa * seg_bytes;
          if (p.gb == 3) {
            mbar_wait(smem_u32(&b_full[sb]), pb);
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
              for (int dx = 0; dx < 3; ++dx) {
                // tap view: same segment, dx pixel rows further in (swizzle phase follows the address)
                const uint64_t adesc = umma_desc_k128(seg + dx * 128, p.desc_bo ? static_cast<uint32_t>(dx) : 0u);
                const uint64_t bdesc = umma_desc_k128(b_base + (sb * 3 + dx) * b_bytes);
                for (int half = 0; half < p.ms; ++half) {
                  // second M sub-tile: 128 rows (16 KB, encoded +1024) further, its own accumulator
                  const uint64_t ad = adesc + static_cast<uint64_t>(half) * 1024u;
                  const uint32_t dt = d_tmem + half * acc_cols;
                  if (ksteps == 4) {
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) umma_f16(dt, ad + 2 * kk, bdesc + 2 * kk, idesc, (first && kk == 0) ? 0u : 1u);
                  } else {
                    for (int kk = 0; kk < ksteps; ++kk) umma_f16(dt, ad + 2 * kk, bdesc + 2 * kk, idesc, (first && kk == 0) ? 0u : 1u);
                  }
                }
                first = 0;
              }
              umma_commit(smem_u32(&b_empty[sb]));
              umma_commit(smem_u32(&a_empty[sa]));
              if (last) umma_commit(smem_u32(&bar_tfull[as]));
            }
            __syncwarp();
            first = 0;
            if (++sb == p.nb) { sb = 0; pb ^= 1u; }
          } else {
            for (int dx = 0; dx < 3; ++dx) {
              mbar_wait(smem_u32(&b_full[sb]), pb);
              tc_fence_after();
              if (elect_one()) {
                const uint64_t adesc = umma_desc_k128(seg + dx * 128, p.desc_bo ? static_cast<uint32_t>(dx) : 0u);
                const uint64_t bdesc = umma_desc_k128(b_base + sb * b_bytes);
                for (int half = 0; half < p.ms; ++half) {
                  const uint64_t ad = adesc + static_cast<uint64_t>(half) * 1024u;
                  const uint32_t dt = d_tmem + half * acc_cols;
                  if (ksteps == 4) {
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) umma_f16(dt, ad + 2 * kk, bdesc + 2 * kk, idesc, (first && kk == 0) ? 0u : 1u);
                  } else {
                    for (int kk = 0; kk < ksteps; ++kk) umma_f16(dt, ad + 2 * kk, bdesc + 2 * kk, idesc, (first && kk == 0) ? 0u : 1u);
                  }
                }
                umma_commit(smem_u32(&b_empty[sb]));
                if (dx == 2) {
                  umma_commit(smem_u32(&a_empty[sa]));
                  if (last) umma_commit(smem_u32(&bar_tfull[as]));
                }
              }
              __syncwarp();
              first = 0;
              if (++sb == p.nb) { sb = 0; pb ^= 1u; }
            }
          }
          if (++sa == p.na) { sa = 0; pa ^= 1u; }
        }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue
    const int q = warp & 3;
    const int part = (warp - 2) >> 2;
    const int etid = threadIdx.x - 64;
    const int row = q * 32 + lane;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int mt = static_cast<int>(fast_div(tile, p.mg_tn)), nt = tile - mt * p.tiles_n;
      const int n0 = nt * p.BN;
      stage_bias(p, s_bias[as], etid, n0);
      mbar_wait(smem_u32(&bar_tfull[as]), aphase);
      tc_fence_after();
      for (int half = 0; half < p.ms; ++half) {
        // a 16-wide N tile (the heads' last conv) has one column chunk: instead of leaving three of the
        // four warps of a quadrant idle, the M sub-tiles are dealt out to them (ncu: dec10 was bound by
        // 4 of 16 epilogue warps walking all sub-tiles serially)
        int epart = part;
        if (p.BN == 16) { if (part != (half & 3)) continue; epart = 0; }
        const int pp = mt * tile_px + half * 128 + row;          // linear padded pixel index
        const int y = static_cast<int>(fast_div(pp, p.mg_wp)), x = pp - y * p.WP;
        const bool inside = (pp < p.NP) && y >= 1 && y <= p.H && x >= 1 && x <= p.W;
        EpiPix px;
        px.ok = inside;
        px.zero = (pp < p.NP) && !inside && p.out_pad;
        const uint32_t upix = static_cast<uint32_t>((y - 1) * p.W + (x - 1));   // unpadded index (if inside)
        px.ooff = (p.out_pad ? static_cast<uint32_t>(pp) : upix) * p.ldo;
        px.roff = (p.res_pad ? static_cast<uint32_t>(pp) : upix) * p.ldr;
        px.fpix = upix;
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * kAccStride + half * acc_cols;
        epilogue_tile<E>(p, t_row, n0, s_bias[as], epart, px);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bar_tempty[as]));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// (3) linear padded formulation on a CTA PAIR (cta_group::2, cluster of two CTAs on one TPC)
//
// One pair tile = two adjacent M tiles (CTA rank r owns tile 2*pm + r) x one N tile.  Each CTA stages
// its own activation segment and HALF of every weight tile (BN/2 rows); the leader issues
// tcgen05.mma.cta_group::2 with M = 256 and both CTAs' TMEM receive their 128 rows x BN columns.
// Per SM this halves the weight bytes that have to be staged and read per MMA cycle — the quantity
// that bounds the 1-CTA kernel (ring depth < TMA round trip at BN = 256, 128 B/clk smem reads at
// BN = 128) — while pixels, accumulators and the epilogue stay exactly as in conv3x3_lin_kernel.
//
// Barriers: full barriers live in the leader and collect the bytes of both CTAs' loads
// (cp.async.bulk.tensor ... cta_group::2 signals the leader's barrier); empty / accumulator-full
// barriers exist in both CTAs and are signalled together by one multicast tcgen05.commit; the
// accumulator-empty barrier lives in the leader and counts the epilogue warps of both CTAs.
// ------------------------------------------------------------------------------------------------
template <class E>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
conv3x3_pair_kernel(const __grid_constant__ CUtensorMap mapA,
                    const __grid_constant__ CUtensorMap mapB, const ConvKParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full[kMaxRing], a_empty[kMaxRing];
  __shared__ __align__(8) uint64_t b_full[kMaxRing], b_empty[kMaxRing];
  __shared__ __align__(8) uint64_t bar_tfull[2], bar_tempty[2];
  __shared__ uint32_t tmem_holder;
  __shared__ __align__(16) float s_bias[2][256];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cid = blockIdx.x >> 1, ncl = gridDim.x >> 1;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bh_bytes = static_cast<uint32_t>(p.BN >> 1) * 128u;   // this CTA's half of a weight tile
  const uint32_t seg_bytes = seg_slot_bytes(p.ms);
  const uint32_t acc_cols = p.BN <= 64 ? 64u : 128u;
  const uint32_t b_base = smem_base + p.na * seg_bytes;
  const int tile_px = 128 * p.ms;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapA);
    tma_prefetch_desc(&mapB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.na; ++s) { mbar_init(smem_u32(&a_full[s]), 1); mbar_init(smem_u32(&a_empty[s]), 1); }
    for (int s = 0; s < p.nb; ++s) { mbar_init(smem_u32(&b_full[s]), 1); mbar_init(smem_u32(&b_empty[s]), 1); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&bar_tfull[s]), 1);
      mbar_init(smem_u32(&bar_tempty[s]), 2 * kEpiWarps);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc2(smem_u32(&tmem_holder), 512);
    tmem_relinquish2();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // peer barriers initialised before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = tmem_holder;
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs)
    int sa = 0, sb = 0;
    uint32_t pa = 0, pb = 0;
    for (int tile = cid; tile < p.total_tiles; tile += ncl) {
      const int pm = static_cast<int>(fast_div(tile, p.mg_tn)), nt = tile - pm * p.tiles_n;
      const int p0 = (pm * 2 + static_cast<int>(rank)) * tile_px;
      const int n0 = nt * p.BN + static_cast<int>(rank) * (p.BN >> 1);
      for (int c = 0; c < p.kchunks; ++c) {
        for (int dy = 0; dy < 3; ++dy) {
          mbar_wait(smem_u32(&a_empty[sa]), pa ^ 1u);
          if (elect_one()) {
            const uint32_t af = smem_u32(&a_full[sa]) & kPeerBitMask;
            if (rank == 0) mbar_arrive_expect_tx(af, 2 * kSegRows * 128 * p.ms);
            const int r0 = p0 + (dy - 1) * p.WP - 1;
            for (int j = 0; j < p.ms; ++j)
              tma_load_2d_pair(smem_base + sa * seg_bytes + j * 128 * 128, &mapA, af, c * 64, r0 + j * 128);
          }
          __syncwarp();
          if (++sa == p.na) { sa = 0; pa ^= 1u; }
          mbar_wait(smem_u32(&b_empty[sb]), pb ^ 1u);
          if (elect_one()) {
            const uint32_t bf = smem_u32(&b_full[sb]) & kPeerBitMask;
            if (rank == 0) mbar_arrive_expect_tx(bf, 2 * 3 * bh_bytes);
#pragma unroll
            for (int dx = 0; dx < 3; ++dx)
              tma_load_3d_pair(b_base + (sb * 3 + dx) * bh_bytes, &mapB, bf, c * 64, n0, dy * 3 + dx);
          }
          __syncwarp();
          if (++sb == p.nb) { sb = 0; pb ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (leader CTA only)
    if (rank == 0) {
      const uint32_t idesc = umma_idesc(E::kUmmaFmt, 256, p.BN);
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      int it = 0;
      for (int tile = cid; tile < p.total_tiles; tile += ncl, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(smem_u32(&bar_tempty[as]), aphase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * kAccStride;
        uint32_t first = 1;
        for (int c = 0; c < p.kchunks; ++c) {
          const int kvalid = min(64, p.Cin - c * 64);
          const int ksteps = (kvalid + 15) >> 4;
          for (int dy = 0; dy < 3; ++dy) {
            const bool last = (c == p.kchunks - 1) && (dy == 2);
            mbar_wait(smem_u32(&a_full[sa]), pa);
            mbar_wait(smem_u32(&b_full[sb]), pb);
            tc_fence_after();
            const uint32_t seg = smem_base + sa * seg_bytes;
            if (elect_one()) {
#pragma unroll
              for (int dx = 0; dx < 3; ++dx) {
                const uint64_t adesc = umma_desc_k128(seg + dx * 128);
                const uint64_t bdesc = umma_desc_k128(b_base + (sb * 3 + dx) * bh_bytes);
                for (int half = 0; half < p.ms; ++half) {
                  const uint64_t ad = adesc + static_cast<uint64_t>(half) * 1024u;
                  const uint32_t dt = d_tmem + half * acc_cols;
                  if (ksteps == 4) {
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) umma_f16_2cta(dt, ad + 2 * kk, bdesc + 2 * kk, idesc, (first && kk == 0) ? 0u : 1u);
                  } else {
                    for (int kk = 0; kk < ksteps; ++kk) umma_f16_2cta(dt, ad + 2 * kk, bdesc + 2 * kk, idesc, (first && kk == 0) ? 0u : 1u);
                  }
                }
                first = 0;
              }
              umma_commit_pair(smem_u32(&b_empty[sb]));
              umma_commit_pair(smem_u32(&a_empty[sa]));
              if (last) umma_commit_pair(smem_u32(&bar_tfull[as]));
            }
            __syncwarp();
            first = 0;
            if (++sb == p.nb) { sb = 0; pb ^= 1u; }
            if (++sa == p.na) { sa = 0; pa ^= 1u; }
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (both CTAs, own 128 x ms rows)
    const int q = warp & 3;
    const int part = (warp - 2) >> 2;
    const int etid = threadIdx.x - 64;
    const int row = q * 32 + lane;
    int it = 0;
    for (int tile = cid; tile < p.total_tiles; tile += ncl, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int pm = static_cast<int>(fast_div(tile, p.mg_tn)), nt = tile - pm * p.tiles_n;
      const int mt = pm * 2 + static_cast<int>(rank);
      const int n0 = nt * p.BN;
      stage_bias(p, s_bias[as], etid, n0);
      mbar_wait(smem_u32(&bar_tfull[as]), aphase);
      tc_fence_after();
      for (int half = 0; half < p.ms; ++half) {
        const int pp = mt * tile_px + half * 128 + row;
        const int y = static_cast<int>(fast_div(pp, p.mg_wp)), x = pp - y * p.WP;
        const bool inside = (pp < p.NP) && y >= 1 && y <= p.H && x >= 1 && x <= p.W;
        EpiPix px;
        px.ok = inside;
        px.zero = (pp < p.NP) && !inside && p.out_pad;
        const uint32_t upix = static_cast<uint32_t>((y - 1) * p.W + (x - 1));   // unpadded index (if inside)
        px.ooff = (p.out_pad ? static_cast<uint32_t>(pp) : upix) * p.ldo;
        px.roff = (p.res_pad ? static_cast<uint32_t>(pp) : upix) * p.ldr;
        px.fpix = upix;
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * kAccStride + half * acc_cols;
        epilogue_tile<E>(p, t_row, n0, s_bias[as], part, px);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(smem_u32(&bar_tempty[as]), 0);
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // the leader's MMAs read the peer's shared memory until the very end
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// (4) linear padded formulation with the K loop split over a cluster (small-M layers)
//
// The 10x20 / 20x40 layers (context, first neck block) have 3-8 M tiles but 36-180 (chunk, kernel-row)
// groups in their K loop: with one CTA per output tile at most 48-96 SMs work and each walks a long
// latency-bound K loop.  Here a cluster of S = 2..4 CTAs shares ONE output tile (128 pixels x BN):
// CTA r accumulates K chunks [r*kc/S, (r+1)*kc/S) into its own TMEM accumulator, writes the fp32 partial
// to its shared memory, and after a cluster barrier CTA r sums the S partials of ITS 128/S rows through
// distributed shared memory in rank order (deterministic) and runs the epilogue (bias, activation,
// residual, zero border).  One tile per CTA (not persistent), ms = 1, gb = 3, BN <= 128.
// ------------------------------------------------------------------------------------------------
template <class E>
__global__ void __launch_bounds__(kThreads, 1)
conv3x3_splitk_kernel(const __grid_constant__ CUtensorMap mapA,
                      const __grid_constant__ CUtensorMap mapB, const ConvKParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full[kMaxRing], a_empty[kMaxRing];
  __shared__ __align__(8) uint64_t b_full[kMaxRing], b_empty[kMaxRing];
  __shared__ __align__(8) uint64_t bar_tfull;
  __shared__ uint32_t tmem_holder;
  __shared__ __align__(16) float s_bias[256];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int S = p.splitk;
  const uint32_t rank = cluster_ctarank();
  const int tile = blockIdx.x / S;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t b_bytes = static_cast<uint32_t>(p.BN) * 128u;
  const uint32_t seg_bytes = seg_slot_bytes(1);
  const uint32_t b_base = smem_base + p.na * seg_bytes;
  const int mt = static_cast<int>(fast_div(tile, p.mg_tn)), nt = tile - mt * p.tiles_n;
  const int p0 = mt * 128, n0 = nt * p.BN;
  const int c_beg = static_cast<int>(rank) * p.kchunks / S, c_end = (static_cast<int>(rank) + 1) * p.kchunks / S;
  const int pstride = p.BN + 4;                       // floats per row of the partial buffer (bank spreading)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapA);
    tma_prefetch_desc(&mapB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.na; ++s) { mbar_init(smem_u32(&a_full[s]), 1); mbar_init(smem_u32(&a_empty[s]), 1); }
    for (int s = 0; s < p.nb; ++s) { mbar_init(smem_u32(&b_full[s]), 1); mbar_init(smem_u32(&b_empty[s]), 1); }
    mbar_init(smem_u32(&bar_tfull), 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(smem_u32(&tmem_holder), 128);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_holder;
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    int sa = 0, sb = 0;
    uint32_t pa = 0, pb = 0;
    for (int c = c_beg; c < c_end; ++c) {
      for (int dy = 0; dy < 3; ++dy) {
        mbar_wait(smem_u32(&a_empty[sa]), pa ^ 1u);
        if (elect_one()) {
          const uint32_t af = smem_u32(&a_full[sa]);
          mbar_arrive_expect_tx(af, kSegRows * 128);
          tma_load_2d(smem_base + sa * seg_bytes, &mapA, af, c * 64, p0 + (dy - 1) * p.WP - 1);
        }
        __syncwarp();
        if (++sa == p.na) { sa = 0; pa ^= 1u; }
        mbar_wait(smem_u32(&b_empty[sb]), pb ^ 1u);
        if (elect_one()) {
          const uint32_t bf = smem_u32(&b_full[sb]);
          mbar_arrive_expect_tx(bf, 3 * b_bytes);
#pragma unroll
          for (int dx = 0; dx < 3; ++dx)
            tma_load_3d(b_base + (sb * 3 + dx) * b_bytes, &mapB, bf, c * 64, n0, dy * 3 + dx);
        }
        __syncwarp();
        if (++sb == p.nb) { sb = 0; pb ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    const uint32_t idesc = umma_idesc(E::kUmmaFmt, 128, p.BN);
    int sa = 0, sb = 0;
    uint32_t pa = 0, pb = 0;
    uint32_t first = 1;
    for (int c = c_beg; c < c_end; ++c) {
      const int kvalid = min(64, p.Cin - c * 64);
      const int ksteps = (kvalid + 15) >> 4;
      for (int dy = 0; dy < 3; ++dy) {
        const bool last = (c == c_end - 1) && (dy == 2);
        mbar_wait(smem_u32(&a_full[sa]), pa);
        mbar_wait(smem_u32(&b_full[sb]), pb);
        tc_fence_after();
        const uint32_t seg = smem_base + sa * seg_bytes;
        if (elect_one()) {
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) {
            const uint64_t adesc = umma_desc_k128(seg + dx * 128);
            const uint64_t bdesc = umma_desc_k128(b_base + (sb * 3 + dx) * b_bytes);
            for (int kk = 0; kk < ksteps; ++kk)
              umma_f16(tmem_base, adesc + 2 * kk, bdesc + 2 * kk, idesc, (first && kk == 0) ? 0u : 1u);
            first = 0;
          }
          umma_commit(smem_u32(&b_empty[sb]));
          umma_commit(smem_u32(&a_empty[sa]));
          if (last) umma_commit(smem_u32(&bar_tfull));
        }
        __syncwarp();
        first = 0;
        if (++sb == p.nb) { sb = 0; pb ^= 1u; }
        if (++sa == p.na) { sa = 0; pa ^= 1u; }
      }
    }
  } else {
    // ------------------------------------------------------------ accumulator -> fp32 partial in shared memory
    const int q = warp & 3;
    const int part = (warp - 2) >> 2;
    const int etid = threadIdx.x - 64;
    const int row = q * 32 + lane;
    stage_bias(p, s_bias, etid, n0);
    if (c_end > c_beg) mbar_wait(smem_u32(&bar_tfull), 0);   // every MMA of this CTA has retired: the ring may be overwritten
    tc_fence_after();
    float* part_buf = reinterpret_cast<float*>(smem_raw + (smem_base - smem_u32(smem_raw)));
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const int nchunks = p.BN >> 4;
    for (int chunk = part; chunk < nchunks; chunk += 4) {
      uint32_t rr[16];
      tmem_ld16(t_row + chunk * 16, rr);
      tmem_ld_wait();
      if (c_end <= c_beg) {                 // no K chunk for this rank (S > kchunks): contributes zeros
#pragma unroll
        for (int i = 0; i < 16; ++i) rr[i] = 0u;
      }
      float4* dst = reinterpret_cast<float4*>(part_buf + row * pstride + chunk * 16);
#pragma unroll
      for (int i = 0; i < 4; ++i)
        dst[i] = make_float4(__uint_as_float(rr[4 * i]), __uint_as_float(rr[4 * i + 1]), __uint_as_float(rr[4 * i + 2]),
                             __uint_as_float(rr[4 * i + 3]));
    }
    tc_fence_before();
  }

  __syncthreads();
  cluster_sync_all();                         // all S partials are in shared memory and visible cluster-wide

  if (warp >= 2) {
    // ------------------------------------------------------------ reduce my rows over the cluster + epilogue
    const int etid = threadIdx.x - 64;
    const int rows_per = (128 + S - 1) / S;
    const int r_beg = static_cast<int>(rank) * rows_per, r_end = min(128, r_beg + rows_per);
    const int c4n = p.BN >> 2;                                      // float4 groups per row
    const uint32_t part_addr = smem_base;
    typename E::T* out = reinterpret_cast<typename E::T*>(p.out);
    const typename E::T* res = reinterpret_cast<const typename E::T*>(p.res);
    for (int item = etid; item < (r_end - r_beg) * c4n; item += kEpiWarps * 32) {
      const int rl = item / c4n, c4 = item - rl * c4n;
      const int row = r_beg + rl;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      const uint32_t addr = part_addr + static_cast<uint32_t>(row * pstride + c4 * 4) * 4u;
      for (int s = 0; s < S; ++s) {                                 // fixed order: deterministic
        const float4 v = ld_dsmem_f4(addr, static_cast<uint32_t>(s));
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
      const int pp = p0 + row;
      if (pp >= p.NP) continue;
      const int y = static_cast<int>(fast_div(pp, p.mg_wp)), x = pp - y * p.WP;
      const bool inside = y >= 1 && y <= p.H && x >= 1 && x <= p.W;
      const int n = n0 + c4 * 4;
      if (n >= p.nlim) continue;
      const uint32_t upix = static_cast<uint32_t>((y - 1) * p.W + (x - 1));
      typename E::T* op = out + ((p.out_pad ? static_cast<uint32_t>(pp) : upix) * p.ldo + n);
      if (!inside) {
        if (p.out_pad) *reinterpret_cast<uint2*>(op) = make_uint2(0u, 0u);
        continue;
      }
      const float4 b4 = *reinterpret_cast<const float4*>(s_bias + c4 * 4);
      float v[4] = {acc.x + b4.x, acc.y + b4.y, acc.z + b4.z, acc.w + b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (p.act == ACT_GELU) v[i] = act_gelu(v[i]);
        else if (p.act == ACT_SILU) v[i] = act_silu(v[i]);
        else if (p.act == ACT_SIGMOID) v[i] = (n + i < p.Cout) ? act_sigmoid(v[i]) : 0.f;
      }
      if ((p.mode == VPB_EPI_ADD || p.mode == VPB_EPI_MULADD) && n < p.ldr) {
        const typename E::T* rp = res + ((p.res_pad ? static_cast<uint32_t>(pp) : upix) * p.ldr + n);
        const uint2 rv = *reinterpret_cast<const uint2*>(rp);
        const float2 f0 = unpack2<E>(rv.x), f1 = unpack2<E>(rv.y);
        const float f[4] = {f0.x, f0.y, f1.x, f1.y};
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = (p.mode == VPB_EPI_ADD) ? v[i] + f[i] : fmaf(v[i], f[i], f[i]);
      }
      *reinterpret_cast<uint2*>(op) = make_uint2(pack2<E>(v[0], v[1]), pack2<E>(v[2], v[3]));
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                         // peers may still be reading this CTA's partial
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}


// ------------------------------------------------------------------------------------------------
// (5) weight-stationary ConvTranspose2d(k2, s2) [+ fused 1x1 skip link] on a CTA pair
//
// The tile kernel re-streams the weights of a (phase, N tile) for every pixel tile: for upsample_layer_3
// (256 -> 256 at 80x160) that is 128 KB of weights + 48 KB of skip-link operands against 64 KB of activations per tile —
// ncu: tensor pipe 20 %, tile period 12.8 k cycles (profiles/r2_ncu_convt_post.md).  Here a CTA PAIR (cluster of 2,
// tcgen05.mma.cta_group::2, M = 256 = two pixel tiles) owns ONE weight set (phase, N tile) for the whole launch: every K
// chunk of it (and of the skip link) is loaded once, HALF of its rows into each CTA, and only the activation tiles stream
// through a ring; pixel-tile pairs of a set are dealt round-robin to the clusters that share it.  The in-kernel timeline
// of the single-CTA form of this kernel (gpurun_out/r2k_trace_ws_up3.txt) showed it MMA-issue bound at ~140 cycles per
// M128 x N128 MMA (8 KB of operands per 64 cycles of math against the ~64 B/clk shared-memory operand path); the pair
// halves the weight bytes per CTA, so N = 256 (4 + 4 KB per 128 cycles) runs at the full rate.  Accumulators
// double-buffered in TMEM, epilogue per CTA through swizzled shared-memory slabs and TMA stores (epilogue_to_smem).
// ------------------------------------------------------------------------------------------------
template <class E>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
convt_ws_kernel(const __grid_constant__ ConvMaps maps, const ConvKParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full[kMaxStages], a_empty[kMaxStages];
  __shared__ __align__(8) uint64_t b_full;
  __shared__ __align__(8) uint64_t bar_tfull[2], bar_tempty[2];
  __shared__ uint32_t tmem_holder;
  __shared__ __align__(16) float s_bias[256];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cid = blockIdx.x >> 1, ncl = gridDim.x >> 1;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bh_bytes = static_cast<uint32_t>(p.BN >> 1) * 128u;      // this CTA's half of a weight chunk
  const int nB = p.kchunks + p.kchunks2;
  const uint32_t b_base = smem_base + static_cast<uint32_t>(p.stages) * kATileBytes;
  const uint32_t slab0 = b_base + static_cast<uint32_t>(nB) * bh_bytes;
  // this cluster's weight set and its share of the pixel-tile pairs
  const int S = p.phases * p.tiles_n;
  const int set = cid % S, j0 = cid / S;
  const int nper = (ncl - set + S - 1) / S;
  const int ph = static_cast<int>(fast_div(set, p.mg_tn)), nt = set - ph * p.tiles_n, n0 = nt * p.BN;
  const int npt = p.tiles_h * p.tiles_w, npairs = (npt + 1) >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.A); tma_prefetch_desc(&maps.B); tma_prefetch_desc(&maps.O);
    if (p.kchunks2) { tma_prefetch_desc(&maps.A2); tma_prefetch_desc(&maps.B2); }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(smem_u32(&a_full[s]), 1); mbar_init(smem_u32(&a_empty[s]), 1); }
    mbar_init(smem_u32(&b_full), 1);
    for (int s = 0; s < 2; ++s) { mbar_init(smem_u32(&bar_tfull[s]), 1); mbar_init(smem_u32(&bar_tempty[s]), 2 * kEpiWarps); }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc2(smem_u32(&tmem_holder), 512);
    tmem_relinquish2();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // peer barriers initialised before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = tmem_holder;
  pdl_launch_dependents();
  pdl_wait();
  const bool tr = p.trace && blockIdx.x == 0;
  if (tr && threadIdx.x == 0) p.trace[255] = clock64();

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs: own pixels, own half of B)
    if (elect_one()) {
      const uint32_t bf = smem_u32(&b_full) & kPeerBitMask;
      if (rank == 0) mbar_arrive_expect_tx(bf, 2u * static_cast<uint32_t>(nB) * bh_bytes);
      const int nr = n0 + static_cast<int>(rank) * (p.BN >> 1);
      for (int c = 0; c < p.kchunks; ++c) tma_load_3d_pair(b_base + c * bh_bytes, &maps.B, bf, c * 64, nr, ph);
      for (int c2 = 0; c2 < p.kchunks2; ++c2) tma_load_3d_pair(b_base + (p.kchunks + c2) * bh_bytes, &maps.B2, bf, c2 * 64, nr, 0);
    }
    __syncwarp();
    int stage = 0;
    uint32_t phase = 0;
    for (int pp = j0; pp < npairs; pp += nper) {
      const int pt = 2 * pp + static_cast<int>(rank);       // may be == npt for an odd count: every row out of bounds -> zeros
      const int thi = static_cast<int>(fast_div(pt, p.mg_tw)), twi = pt - thi * p.tiles_w;
      const int h0 = thi * p.TH, w0 = twi * p.TW;
      for (int k = 0; k < nB; ++k) {
        mbar_wait(smem_u32(&a_empty[stage]), phase ^ 1u);
        if (elect_one()) {
          const uint32_t full = smem_u32(&a_full[stage]) & kPeerBitMask;
          if (rank == 0) mbar_arrive_expect_tx(full, 2u * kATileBytes);
          if (k < p.kchunks) tma_load_4d_pair(smem_base + stage * kATileBytes, &maps.A, full, k * 64, w0, h0, 0);
          else tma_load_5d_pair(smem_base + stage * kATileBytes, &maps.A2, full, (k - p.kchunks) * 64, ph & 1, w0, ph >> 1, h0);
          if (tr) { const int ti = (pp - j0) / nper; if (ti < 8 && (k == 0 || k == nB - 1)) p.trace[ti * 16 + (k ? 1 : 0)] = clock64(); }
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (leader CTA only)
    if (rank == 0) {
      const uint32_t idesc = umma_idesc(E::kUmmaFmt, 256, p.BN);
      mbar_wait(smem_u32(&b_full), 0);
      if (tr && lane == 0) p.trace[254] = clock64();
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int pp = j0; pp < npairs; pp += nper, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(smem_u32(&bar_tempty[as]), aphase ^ 1u);
        tc_fence_after();
        if (tr && lane == 0 && it < 8) p.trace[it * 16 + 9] = clock64();
        const uint32_t d_tmem = tmem_base + as * kAccStride;
        for (int k = 0; k < nB; ++k) {
          const int kvalid = k < p.kchunks ? min(64, p.Cin - k * 64) : min(64, p.Cin2 - (k - p.kchunks) * 64);
          const int ksteps = (kvalid + 15) >> 4;
          mbar_wait(smem_u32(&a_full[stage]), phase);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t adesc = umma_desc_k128(smem_base + stage * kATileBytes);
            const uint64_t bdesc = umma_desc_k128(b_base + k * bh_bytes);
            if (tr && it < 8 && k == 0) p.trace[it * 16 + 4] = clock64();
            for (int kk = 0; kk < ksteps; ++kk) umma_f16_2cta(d_tmem, adesc + 2 * kk, bdesc + 2 * kk, idesc, (k | kk) != 0 ? 1u : 0u);
            umma_commit_pair(smem_u32(&a_empty[stage]));
            if (k == nB - 1) { umma_commit_pair(smem_u32(&bar_tfull[as])); if (tr && it < 8) p.trace[it * 16 + 8] = clock64(); }
          }
          __syncwarp();
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (both CTAs: own 128 pixels x BN channels)
    const int q = warp & 3;
    const int part = (warp - 2) >> 2;
    const int etid = threadIdx.x - 64;
    const int row = q * 32 + lane;
    stage_bias(p, s_bias, etid, n0);          // one N tile per cluster: staged once
    // slab groups: nslab = BN / 64 in {1, 2, 4}; part p drains chunks [p * nslab, (p + 1) * nslab) = slab p * nslab / 4
    const int nslab = p.BN >> 6;
    const int slab = (part * nslab) >> 2;
    const int gthreads = 512 / nslab;
    const bool gleader = (q == 0) && (lane == 0) && (part == (slab * 4) / nslab);
    int it = 0;
    for (int pp = j0; pp < npairs; pp += nper, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int pt = 2 * pp + static_cast<int>(rank);
      const int thi = static_cast<int>(fast_div(pt, p.mg_tw)), twi = pt - thi * p.tiles_w;
      mbar_wait(smem_u32(&bar_tfull[as]), aphase);
      tc_fence_after();
      if (tr && gleader && slab == 0 && it < 8) p.trace[it * 16 + 10] = clock64();
      // every 64-channel slab has its own group of warps, its own barrier and its own store: a slab's store is in
      // flight while the other slabs are still being drained, and only that slab's previous store is waited for
      // (the single-buffer form cost ~1.5 k cycles of waiting per tile, gpurun_out/r2l_trace_ws_up3.txt)
      if (gleader) bulk_wait_read0();
      named_bar_sync(2 + slab, gthreads);
      if (tr && gleader && slab == 0 && it < 8) p.trace[it * 16 + 11] = clock64();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * kAccStride;
      if (p.act == ACT_GELU) epilogue_to_smem<E, true>(p, t_row, s_bias, part, slab0, row, part * nslab, nslab);
      else epilogue_to_smem<E, false>(p, t_row, s_bias, part, slab0, row, part * nslab, nslab);
      fence_proxy_async();
      tc_fence_before();
      named_bar_sync(2 + slab, gthreads);
      if (tr && gleader && slab == 0 && it < 8) p.trace[it * 16 + 12] = clock64();
      if (gleader && pt < npt) {
        tma_store_5d(&maps.O, slab0 + slab * (128 * 128), n0 + slab * 64, ph & 1, twi * p.TW, ph >> 1, thi * p.TH);
        bulk_commit();
        if (tr && slab == 0 && it < 8) p.trace[it * 16 + 13] = clock64();
      }
      if (lane == 0) mbar_arrive_cluster(smem_u32(&bar_tempty[as]), 0);
    }
    if (gleader) { bulk_wait0(); if (tr && slab == 0) p.trace[253] = clock64(); }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // the leader's MMAs read the peer's shared memory until the very end
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, 512);
  }
}


// ------------------------------------------------------------------------------------------------
// (6) "upconv": ConvTranspose2d(k2, s2) [+ Conv1x1(skip)] followed by Conv3x3 + bias + GELU, as ONE GEMM on a CTA pair
//
// The reference sums the ConvTranspose and the skip link without an activation and feeds the sum to a 3x3 convolution
// (scene_neck.py:30-37, scene_seg_head.py:25-33), so the two linear layers compose exactly: output phase (a, b) — the
// pixels (2h+a, 2w+b) — is a 2x2 convolution of the LOW-resolution input with weights
//     Wf[a,b][ty,tx] = sum over the 3x3 taps (dy,dx) that land on low-res offset (ty-1+a, tx-1+b) of  W3[dy,dx] . Wt[a',b']
// plus a 3x3 convolution of the skip tensor with W3[dy,dx] . Wskip, plus a bias that depends only on which 3x3 taps fall
// inside the image (9 border classes).  K per output pixel drops from 9*Cmid (+ Cin + C2 for the ConvTranspose) to
// 4*Cin + 9*C2, the upsampled tensor is never written or read, and one kernel replaces two (vpb_upconv_compose builds the
// weights at load time; DESIGN.md 3e has the algebra and the flop table).
// Tile = (phase, N tile, pair of 128-pixel low-res tiles); stage = this CTA's 128 x 64 activation box + HALF of the
// BN x 64 weight box (tcgen05.mma.cta_group::2, M = 256); accumulators double-buffered in TMEM; epilogue per 64-channel
// slab through swizzled shared memory and a TMA store into the [h][a][w][b][c] view of the output (as convt_ws_kernel).
// ------------------------------------------------------------------------------------------------
template <class E>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
upconv_pair_kernel(const __grid_constant__ ConvMaps maps, const ConvKParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[kUpcStages], bar_empty[kUpcStages];
  __shared__ __align__(8) uint64_t bar_tfull[2], bar_tempty[2];
  __shared__ uint32_t tmem_holder;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cid = blockIdx.x >> 1, ncl = gridDim.x >> 1;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bh_bytes = static_cast<uint32_t>(p.BN >> 1) * 128u;      // this CTA's half of a weight box
  const uint32_t stage_bytes = kATileBytes + bh_bytes;
  const int nslab = p.BN >> 6;
  const uint32_t slab0 = smem_base + static_cast<uint32_t>(p.stages) * stage_bytes;
  // tma_store == 0 (direct-store epilogue): no slabs, the ring gets their shared memory
  float* s_bias = reinterpret_cast<float*>(smem_raw + (slab0 - smem_u32(smem_raw)) + (p.tma_store ? nslab * (128 * 128) : 0));   // [tiles_n * BN]
  const int npt = p.tiles_h * p.tiles_w, npairs = (npt + 1) >> 1;
  const int kc1 = p.kchunks, kc2 = p.kchunks2;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.A); tma_prefetch_desc(&maps.B); tma_prefetch_desc(&maps.O);
    if (kc2) { tma_prefetch_desc(&maps.A2); tma_prefetch_desc(&maps.B2); }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(smem_u32(&bar_full[s]), 1); mbar_init(smem_u32(&bar_empty[s]), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(smem_u32(&bar_tfull[s]), 1); mbar_init(smem_u32(&bar_tempty[s]), 2 * kEpiWarps); }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc2(smem_u32(&tmem_holder), 512);
    tmem_relinquish2();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // peer barriers initialised before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = tmem_holder;
  pdl_launch_dependents();
  pdl_wait();
  const bool tr = p.trace && blockIdx.x == 0;
  if (tr && threadIdx.x == 0) p.trace[255] = clock64();

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs: own pixels, own half of the weights)
    int stage = 0;
    uint32_t phase = 0;
    int ti = 0;
    for (int tile = cid; tile < p.total_tiles; tile += ncl, ++ti) {
      const int set = static_cast<int>(fast_div(tile, p.mg_tpp)), pp = tile - set * npairs;
      const int ph = static_cast<int>(fast_div(set, p.mg_tn)), nt = set - ph * p.tiles_n;
      const int pt = 2 * pp + static_cast<int>(rank);       // may be == npt for an odd count: every row out of bounds -> zeros
      const int thi = static_cast<int>(fast_div(pt, p.mg_tw)), twi = pt - thi * p.tiles_w;
      const int h0 = thi * p.TH, w0 = twi * p.TW, pa = ph >> 1, pb = ph & 1;
      const int nr = nt * p.BN + static_cast<int>(rank) * (p.BN >> 1);
      // the four low-resolution taps of this phase: offsets (ty - 1 + a, tx - 1 + b)
      for (int t = 0; t < 4; ++t) {
        const int oy = (t >> 1) - 1 + pa, ox = (t & 1) - 1 + pb;
        for (int c = 0; c < kc1; ++c) {
          mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1u);
          if (elect_one()) {
            const uint32_t full = smem_u32(&bar_full[stage]) & kPeerBitMask;
            const uint32_t sa = smem_base + stage * stage_bytes;
            if (rank == 0) mbar_arrive_expect_tx(full, 2u * stage_bytes);
            tma_load_4d_pair(sa, &maps.A, full, c * 64, w0 + ox, h0 + oy, 0);
            tma_load_3d_pair(sa + kATileBytes, &maps.B, full, c * 64, nr, ph * 4 + t);
            if (tr && ti < 8) { if (t == 0 && c == 0) p.trace[ti * 16] = clock64(); if (t == 3 && c == kc1 - 1) p.trace[ti * 16 + 1] = clock64(); }
          }
          __syncwarp();
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
      }
      // the nine taps of the skip tensor (output resolution, [h][a][w][b][c] view): hi-res row 2h + a + dy - 1 is low-res
      // row h + floor(u / 2) of row phase u & 1, u = a + dy - 1; rows / columns outside the image are zero-filled by TMA
      if (kc2) {
        for (int t2 = 0; t2 < 9; ++t2) {
          const int dy = t2 / 3, dx = t2 - dy * 3;
          const int u = pa + dy - 1, v = pb + dx - 1;
          for (int c2 = 0; c2 < kc2; ++c2) {
            mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1u);
            if (elect_one()) {
              const uint32_t full = smem_u32(&bar_full[stage]) & kPeerBitMask;
              const uint32_t sa = smem_base + stage * stage_bytes;
              if (rank == 0) mbar_arrive_expect_tx(full, 2u * stage_bytes);
              tma_load_5d_pair(sa, &maps.A2, full, c2 * 64, v & 1, w0 + (v >> 1), u & 1, h0 + (u >> 1));
              tma_load_3d_pair(sa + kATileBytes, &maps.B2, full, c2 * 64, nr, t2);
              if (tr && ti < 8 && t2 == 8 && c2 == kc2 - 1) p.trace[ti * 16 + 2] = clock64();
            }
            __syncwarp();
            if (++stage == p.stages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (leader CTA only)
    if (rank == 0) {
      const uint32_t idesc = umma_idesc(E::kUmmaFmt, 256, p.BN);
      const int n1 = 4 * kc1, nK = n1 + 9 * kc2;
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = cid; tile < p.total_tiles; tile += ncl, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(smem_u32(&bar_tempty[as]), aphase ^ 1u);
        tc_fence_after();
        if (tr && lane == 0 && it < 8) p.trace[it * 16 + 9] = clock64();
        const uint32_t d_tmem = tmem_base + as * kAccStride;
        int c = 0;                    // K chunk inside the current tap
        for (int k = 0; k < nK; ++k) {
          if (k == n1) c = 0;
          const int kvalid = k < n1 ? min(64, p.Cin - c * 64) : min(64, p.Cin2 - c * 64);
          if (++c == (k < n1 ? kc1 : kc2)) c = 0;
          const int ksteps = (kvalid + 15) >> 4;
          mbar_wait(smem_u32(&bar_full[stage]), phase);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t sa = smem_base + stage * stage_bytes;
            const uint64_t adesc = umma_desc_k128(sa);
            const uint64_t bdesc = umma_desc_k128(sa + kATileBytes);
            if (tr && it < 8) { if (k == 0) p.trace[it * 16 + 4] = clock64(); if (k == nK - 1) p.trace[it * 16 + 8] = clock64(); }
            if (ksteps == 4) {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) umma_f16_2cta(d_tmem, adesc + 2 * kk, bdesc + 2 * kk, idesc, (k | kk) != 0 ? 1u : 0u);
            } else {
              for (int kk = 0; kk < ksteps; ++kk) umma_f16_2cta(d_tmem, adesc + 2 * kk, bdesc + 2 * kk, idesc, (k | kk) != 0 ? 1u : 0u);
            }
            umma_commit_pair(smem_u32(&bar_empty[stage]));
            if (k == nK - 1) umma_commit_pair(smem_u32(&bar_tfull[as]));
          }
          __syncwarp();
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (both CTAs: own 128 pixels x BN channels)
    const int q = warp & 3;
    const int part = (warp - 2) >> 2;
    const int etid = threadIdx.x - 64;
    const int row = q * 32 + lane;
    const int lh = row >> p.tw_shift, lw = row & (p.TW - 1);
    // interior bias row (class 4) of every N tile, staged once
    for (int n = etid; n < p.tiles_n * p.BN; n += 512) s_bias[n] = n < p.Cout ? __ldg(p.bias + 4 * p.Cout + n) : 0.f;
    asm volatile("bar.sync 1, 512;" ::: "memory");
    const int slab = (part * nslab) >> 2;
    const int gthreads = 512 / nslab;
    const bool gleader = (q == 0) && (lane == 0) && (part == (slab * 4) / nslab);
    int it = 0;
    for (int tile = cid; tile < p.total_tiles; tile += ncl, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int set = static_cast<int>(fast_div(tile, p.mg_tpp)), pp = tile - set * npairs;
      const int ph = static_cast<int>(fast_div(set, p.mg_tn)), nt = set - ph * p.tiles_n, n0 = nt * p.BN;
      const int pt = 2 * pp + static_cast<int>(rank);
      const int thi = static_cast<int>(fast_div(pt, p.mg_tw)), twi = pt - thi * p.tiles_w;
      const int pa = ph >> 1, pb = ph & 1;
      const int h = thi * p.TH + lh, w = twi * p.TW + lw;
      // border class of this lane's output pixel (2h + a, 2w + b): first / interior / last row and column
      const int cy = (pa == 0 && h == 0) ? 0 : (pa == 1 && h == p.H - 1) ? 2 : 1;
      const int cx = (pb == 0 && w == 0) ? 0 : (pb == 1 && w == p.W - 1) ? 2 : 1;
      const int cls = cy * 3 + cx;
      const float* gb = cls != 4 ? p.bias + cls * p.Cout + n0 : nullptr;
      mbar_wait(smem_u32(&bar_tfull[as]), aphase);
      tc_fence_after();
      if (tr && gleader && slab == 0 && it < 8) p.trace[it * 16 + 10] = clock64();
      if (!p.tma_store) {
        // direct stores from registers: slower per tile than the staged TMA store, but hidden behind the next tile's
        // MMAs (double-buffered accumulators), and the ring is 2-4 stages deeper without the slabs
        const int oh = 2 * h + pa, ow = 2 * w + pb, Wo = 2 * p.W;
        EpiPix px;
        px.ok = (h < p.H) && (w < p.W) && (pt < npt);
        px.zero = false;
        px.ooff = static_cast<uint32_t>((oh + p.out_pad) * (Wo + 2 * p.out_pad) + (ow + p.out_pad)) * p.ldo;
        px.roff = 0; px.fpix = 0;
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * kAccStride;
        if (p.act == ACT_GELU) epilogue_store_fast<E, ACT_GELU, false, true>(p, t_row, n0, s_bias + n0, part, px, gb);
        else epilogue_store_fast<E, ACT_NONE, false, true>(p, t_row, n0, s_bias + n0, part, px, gb);
        tc_fence_before();
        __syncwarp();
        if (tr && gleader && slab == 0 && it < 8) p.trace[it * 16 + 12] = clock64();
        if (lane == 0) mbar_arrive_cluster(smem_u32(&bar_tempty[as]), 0);
        continue;
      }
      if (gleader) bulk_wait_read0();          // this slab's previous store has read the shared memory
      named_bar_sync(2 + slab, gthreads);
      if (tr && gleader && slab == 0 && it < 8) p.trace[it * 16 + 11] = clock64();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * kAccStride;
      if (p.act == ACT_GELU) epilogue_to_smem<E, true, true>(p, t_row, s_bias + n0, part, slab0, row, part * nslab, nslab, gb);
      else epilogue_to_smem<E, false, true>(p, t_row, s_bias + n0, part, slab0, row, part * nslab, nslab, gb);
      fence_proxy_async();
      tc_fence_before();
      named_bar_sync(2 + slab, gthreads);
      if (tr && gleader && slab == 0 && it < 8) p.trace[it * 16 + 12] = clock64();
      if (gleader && pt < npt) {
        tma_store_5d(&maps.O, slab0 + slab * (128 * 128), n0 + slab * 64, pb, twi * p.TW, pa, thi * p.TH);
        bulk_commit();
      }
      if (lane == 0) mbar_arrive_cluster(smem_u32(&bar_tempty[as]), 0);
    }
    if (gleader && p.tma_store) { bulk_wait0(); if (tr && slab == 0) p.trace[253] = clock64(); }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // the leader's MMAs read the peer's shared memory until the very end
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, 512);
  }
}

// ------------------------------------------------------------------ host side

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
            cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int device_sm_count() {
  int dev = 0, n = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  return n > 0 ? n : 148;
}

static inline size_t b_bytes_ws(int bn) { return static_cast<size_t>(bn) * 128; }

static int pick_bn(int Cout) {
  if (Cout <= 256) return (Cout + 15) / 16 * 16;
  int best = 256, best_waste = 1 << 30;
  for (int bn = 256; bn >= 128; bn -= 16) {
    const int waste = (Cout + bn - 1) / bn * bn - Cout;
    if (waste < best_waste) { best_waste = waste; best = bn; }
  }
  return best;
}

int conv_plan_build(const vpb_conv_args* a, ConvPlan* plan) {
  if (!a || !plan) return VPB_ERR_ARG;
  if (a->Cin <= 0 || (a->Cin & 7) || (a->ldi & 7) || a->ldi < a->Cin) {
    vpb_set_error("conv: Cin=%d ldi=%d must be multiples of 8 (ldi >= Cin)", a->Cin, a->ldi);
    return VPB_ERR_ARG;
  }
  const int cstride = a->stride == 2 ? 2 : 1;
  if (a->stride != 0 && a->stride != 1 && a->stride != 2) { vpb_set_error("conv: stride %d unsupported", a->stride); return VPB_ERR_ARG; }
  if (cstride == 2 && (a->algo == VPB_ALGO_LINEAR || a->phases != 1 || a->in_pad || a->in2 || a->in_lo)) {
    vpb_set_error("conv: stride 2 needs the TILE algorithm on an unpadded input, phases = 1, no second input");
    return VPB_ERR_ARG;
  }
  if (a->ldw != 0 && (a->ldw < a->Cin || (a->ldw & 7) || a->algo == VPB_ALGO_LINEAR)) {
    vpb_set_error("conv: ldw=%d must be a multiple of 8 >= Cin (TILE algorithm)", a->ldw);
    return VPB_ERR_ARG;
  }
  if (a->act2 != ACT_NONE && a->act2 != ACT_SILU) { vpb_set_error("conv: act2 supports SiLU only"); return VPB_ERR_ARG; }
  const bool upc = a->taps == 4 && a->phases == 4;       // fused ConvTranspose -> Conv3x3 (upconv_pair_kernel)
  if (upc && (a->algo == VPB_ALGO_LINEAR || a->in_lo || a->mode != VPB_EPI_STORE || (a->act != ACT_NONE && a->act != ACT_GELU) ||
              !a->bias || (a->Cout & 15) || a->out_slice || cstride != 1 || a->ldw || (a->in2 && a->taps2 != 9))) {
    vpb_set_error("conv: upconv (taps=4, phases=4) needs the TILE algorithm, 16-bit mode, STORE, act NONE|GELU, bias[9][Cout], Cout%%16==0, taps2=9 with in2");
    return VPB_ERR_ARG;
  }
  if (!((a->taps == 9 && a->phases == 1) || (a->taps == 1 && (a->phases == 1 || a->phases == 4)) || upc)) {
    vpb_set_error("conv: unsupported taps=%d phases=%d", a->taps, a->phases);
    return VPB_ERR_ARG;
  }
  const bool lin = a->algo == VPB_ALGO_LINEAR;
  if (lin && (a->taps != 9 || !a->in_pad)) {
    vpb_set_error("conv: the linear-padded algorithm needs a 3x3 conv on a zero-bordered input");
    return VPB_ERR_ARG;
  }
  if (a->mode == VPB_EPI_FINAL) {
    if (a->Cout > 16 || !a->out_f32) {
      vpb_set_error("conv: FINAL mode needs Cout<=16 and out_f32");
      return VPB_ERR_ARG;
    }
  } else {
    if (!a->out || (a->ldo & 7) || a->ldo < a->Cout) {
      vpb_set_error("conv: bad out/ldo=%d (Cout=%d)", a->ldo, a->Cout);
      return VPB_ERR_ARG;
    }
    if ((a->mode == VPB_EPI_ADD || a->mode == VPB_EPI_MULADD) && (!a->res || (a->ldr & 7))) {
      vpb_set_error("conv: residual mode needs res with ldr%%8==0");
      return VPB_ERR_ARG;
    }
  }
  const bool split = a->in_lo != nullptr;
  if (split) {
    if (lin || !a->w_lo || (a->mode != VPB_EPI_FINAL && !a->out_lo) ||
        ((a->mode == VPB_EPI_ADD || a->mode == VPB_EPI_MULADD) && !a->res_lo) || (a->in2 && (!a->in2_lo || !a->w2_lo))) {
      vpb_set_error("conv: split-fp16 mode needs the TILE algorithm and the low halves of every tensor given (w_lo, out_lo, res_lo, in2_lo, w2_lo)");
      return VPB_ERR_ARG;
    }
  }
  if (a->in2) {
    if (lin || (a->taps != 1 && !upc) || !a->w2 || a->Cin2 <= 0 || (a->Cin2 & 7) || (a->ld2 & 7) || a->ld2 < a->Cin2) {
      vpb_set_error("conv: second input needs the TILE algorithm, taps=1, w2, Cin2/ld2 multiples of 8 (Cin2=%d ld2=%d)",
                    a->Cin2, a->ld2);
      return VPB_ERR_ARG;
    }
  }
  {
    // the epilogue addresses pixels with 32-bit element offsets
    const long ho = a->phases == 4 ? 2L * a->H + 2 : a->H + 2, wo = a->phases == 4 ? 2L * a->W + 2 : a->W + 2;
    if (ho * wo * std::max(a->ldo, a->ldr) >= (1L << 31) || static_cast<long>(a->H) * a->W * 16 >= (1L << 31)) {
      vpb_set_error("conv: tensor too large for 32-bit element offsets");
      return VPB_ERR_ARG;
    }
  }
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) {
    vpb_set_error("conv: cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
    return VPB_ERR_CUDA;
  }
  memset(plan, 0, sizeof(*plan));
  ConvKParams& p = plan->p;
  p.H = a->H; p.W = a->W; p.Cin = a->Cin; p.Cout = a->Cout;
  p.taps = a->taps; p.phases = a->phases;
  p.in_pad = a->in_pad ? 1 : 0; p.out_pad = a->out_pad ? 1 : 0; p.res_pad = a->res_pad ? 1 : 0;
  p.lin = lin ? 1 : 0;
  p.split = split ? 1 : 0;
  p.stride = cstride; p.act2 = a->act2;
  p.nlim = a->out_slice ? std::min(a->ldo, (a->Cout + 7) / 8 * 8) : a->ldo;
  p.out_lo = a->out_lo; p.res_lo = a->res_lo;
  p.desc_bo = a->dbg_base_offset ? 1 : 0;
  p.WP = a->W + 2; p.NP = (a->H + 2) * (a->W + 2);
  p.BN = a->bn > 0 ? a->bn : pick_bn(a->Cout);
  if (a->bn <= 0 && a->phases == 4 && !upc && a->Cout > 64 && p.BN % 64) {
    // ConvTranspose: N tiles made of whole 64-channel slabs (the TMA-store epilogue writes one box per slab)
    int best = 256, best_waste = 1 << 30;
    for (int bn = 256; bn >= 128; bn -= 64) {
      const int waste = (a->Cout + bn - 1) / bn * bn - a->Cout;
      if (waste < best_waste) { best_waste = waste; best = bn; }
    }
    p.BN = best;
  }
  // split-K over a cluster for the smallest-M layers (context convs at 10x20: <= 32 tiles at BN = 128)
  p.splitk = 0;
  if (lin && a->mode != VPB_EPI_FINAL && a->dbg_splitk >= 0 && (a->bn <= 0 || a->dbg_splitk >= 2)) {
    int bn = 128;
    if (a->Cout < 128) bn = (a->Cout + 15) / 16 * 16;
    else {
      int best_waste = 1 << 30;
      for (int b = 128; b >= 64; b -= 16) {
        const int waste = (a->Cout + b - 1) / b * b - a->Cout;
        if (waste < best_waste) { best_waste = waste; bn = b; }
      }
    }
    if (a->bn > 0) bn = std::min(a->bn, 128);
    const int tiles = ((p.NP + 127) / 128) * ((a->Cout + bn - 1) / bn);
    const int kc = (a->Cin + 63) / 64;
    int S = 0;
    if (a->dbg_splitk >= 2) S = std::min(std::min(a->dbg_splitk, 8), kc);
    // measured (gpurun_out/bench_conv_v18_splitk.txt): context_layer_6 (30 tiles) 20.1 -> 11.6 us, but the 48-tile
    // 20x40 layers are faster on the CTA-pair kernel (dec0 36.9 vs 41.9 us) -> only the 10x20 layers
    else if (a->Cout >= 128 && tiles <= 32) S = std::min(std::min(4, device_sm_count() / tiles), kc / 2);
    if (S >= 2) { p.splitk = S; p.BN = bn; }
  }
  bool convt_fused_bn = false;
  if (a->bn <= 0 && a->phases == 4 && !upc && p.BN > 128) {
    // ConvTranspose: the fused form (all four phases per tile, 4 accumulators) needs an N tile <= 128;
    // it only pays with >= 2 waves of fused tiles, otherwise keep the wide tile (measured: up3 at
    // BN=256 20 us, at BN=128 unfused 26 us)
    int best = 128, best_waste = 1 << 30;
    for (int bn = 128; bn >= 64; bn -= 16) {
      const int waste = (a->Cout + bn - 1) / bn * bn - a->Cout;
      if (waste < best_waste) { best_waste = waste; best = bn; }
    }
    long sp = -1;
    for (int tw = 128; tw >= 8; tw >>= 1) {
      const int th = 128 / tw;
      const long cost = static_cast<long>((a->H + th - 1) / th) * ((a->W + tw - 1) / tw);
      if (sp < 0 || cost < sp) sp = cost;
    }
    if (a->dbg_ms == 2 || (a->dbg_ms != 1 && !a->in2 && sp * ((a->Cout + best - 1) / best) >= 2 * device_sm_count()))
      { p.BN = best; convt_fused_bn = true; }
  }
  if (!convt_fused_bn && !p.splitk && !upc && a->bn <= 0 && a->Cout >= 128) {
    // small-M layers (context, first neck blocks): trade N-tile width for CTA count so that the
    // persistent grid covers more of the 148 SMs (weights are re-streamed from L2, activations
    // are tiny)
    const long m_tiles = lin ? (static_cast<long>(a->H + 2) * (a->W + 2) + 127) / 128
                             : (static_cast<long>(a->H) * a->W + 127) / 128 * a->phases;
    const int bn_floor = a->phases == 4 ? 128 : 64;   // ConvT measured better at 128 even with 80 tiles (up0)
    while (p.BN > bn_floor && m_tiles * ((a->Cout + p.BN - 1) / p.BN) < 96) {
      const int nb2 = (p.BN / 2 + 15) / 16 * 16;
      if ((a->Cout + nb2 - 1) / nb2 * nb2 - a->Cout > a->Cout / 8) break;   // too much N padding
      p.BN = nb2;
    }
  }
  if (p.BN % 16 || p.BN > 256 || p.BN < 16) {
    vpb_set_error("conv: bad BN %d", p.BN);
    return VPB_ERR_ARG;
  }
  p.tiles_n = (a->Cout + p.BN - 1) / p.BN;
  p.kchunks = (a->Cin + 63) / 64;
  p.Cin2 = a->in2 ? a->Cin2 : 0;
  p.kchunks2 = (p.Cin2 + 63) / 64;
  bool wave_split = false;
  if (lin && !p.splitk && a->bn <= 0 && a->dbg_pair >= 0 && p.BN == 256 && a->Cout % 128 == 0) {
    // wave quantisation on the pair grid (74 clusters): a 256-wide N tile that needs e.g. 1.4 waves
    // costs 2 waves; 128-wide tiles (half the work each) may pack better.  Measured on decode_layer_4
    // (104 pair tiles): BN=256 1153 TF/s, BN=128 1253 TF/s; 128-wide tiles are ~8 % less efficient
    // otherwise (decode_layer_5/6 keep 256).
    const int ncl = std::max(1, device_sm_count() / 2);
    const int pm = ((p.NP + 127) / 128 + 1) / 2;
    const int w256 = (pm * (a->Cout / 256) + ncl - 1) / ncl, w128 = (pm * (a->Cout / 128) + ncl - 1) / ncl;
    if (pm * (a->Cout / 256) >= 96 && w128 * 128 * 1.08 < w256 * 256.0) { p.BN = 128; wave_split = true; }
  }
  p.tiles_n = (a->Cout + p.BN - 1) / p.BN;
  const size_t b_bytes = static_cast<size_t>(p.BN) * 128;
  if (lin) {
    // two M sub-tiles per CTA (256 pixels, two accumulators sharing every weight tile) when the
    // accumulators fit twice (BN <= 128) and the layer still gives >= 2 waves of tiles
    p.ms = 1;
    if (!p.splitk && a->dbg_ms != 1 && p.BN <= 128 && ((p.NP + 255) / 256) * p.tiles_n >= 2 * device_sm_count()) p.ms = 2;
    if (!p.splitk && a->dbg_ms != 1 && a->dbg_ms != 2 && p.BN <= 64 && ((p.NP + 511) / 512) * p.tiles_n >= 2 * device_sm_count()) p.ms = 4;
    if (!p.splitk && a->dbg_ms == 2 && p.BN <= 128) p.ms = 2;
    if (!p.splitk && a->dbg_ms == 4 && p.BN <= 64) p.ms = 4;
    if (wave_split) p.ms = 1;
    const size_t seg = seg_slot_bytes(p.ms);
    p.tiles_m = (p.NP + 128 * p.ms - 1) / (128 * p.ms);
    // CTA pair (cta_group::2, M = 256) whenever the layer still fills most of the SMs — each CTA then
    // stages only half of every weight tile (measured: dec6 1082 -> 1396 TF/s, dec4 889 -> 1117)
    p.pair = (!p.splitk && a->dbg_pair >= 0 && p.BN >= 64 && p.tiles_m >= 2 &&
              (a->dbg_pair == 1 || p.tiles_m * p.tiles_n >= 96)) ? 1 : 0;
    p.total_tiles = (p.pair ? (p.tiles_m + 1) / 2 : p.tiles_m) * p.tiles_n;
    // weight ring: slots of gb tiles; gb = 3 (a whole kernel row per barrier) whenever three tiles
    // (half tiles for a pair) stay <= 48 KB
    p.gb = (p.pair || p.splitk || (a->dbg_gb != 1 && p.BN <= 144)) ? 3 : 1;
    const size_t slot = (p.pair ? b_bytes / 2 : b_bytes) * p.gb;
    const size_t budget = kMaxDynSmem - 1024;
    // One ring round trip costs ~3.1 k cycles whatever the box size (profiles/r1_tma_ring_microbench.md), so
    // what matters is how many (activation segment + kernel-row weight slot) GROUPS are in flight: balance
    // the two rings (the K-heavy small layers were running with 4 segments against 6+ weight slots)
    const int groups = static_cast<int>(budget / (seg + (p.gb == 3 ? slot : 3 * slot)));
    p.na = std::max(2, std::min(kMaxRing, groups));
    while (p.na > 2 && static_cast<size_t>(p.na) * seg + 2 * slot > budget) --p.na;
    p.nb = static_cast<int>(std::min<size_t>(kMaxRing, (budget - static_cast<size_t>(p.na) * seg) / slot));
    if (p.nb < 2) { vpb_set_error("conv: no room for the weight ring"); return VPB_ERR_ARG; }
    plan->smem_bytes = static_cast<size_t>(p.na) * seg + p.nb * slot + 1024;
    if (p.splitk) plan->smem_bytes = std::max(plan->smem_bytes, static_cast<size_t>(128) * (p.BN + 4) * 4 + 1024);
    p.TW = 128; p.TH = 1; p.tw_shift = 7;
  } else {
    // spatial tile: minimise padded pixels, prefer wide tiles
    int best_tw = 128; long best_cost = -1;
    for (int tw = 128; tw >= 8; tw >>= 1) {
      const int th = 128 / tw;
      const long cost = static_cast<long>((a->H + th - 1) / th) * ((a->W + tw - 1) / tw);
      if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_tw = tw; }
    }
    p.TW = best_tw; p.TH = 128 / best_tw;
    p.tw_shift = 0; while ((1 << p.tw_shift) < p.TW) ++p.tw_shift;
    p.tiles_h = (a->H + p.TH - 1) / p.TH;
    p.tiles_w = (a->W + p.TW - 1) / p.TW;
    // ConvTranspose: all four phases of a pixel tile in one CTA tile (A fetched once per K chunk,
    // four 128-column accumulators) whenever the N tile is at most 128 wide
    // ... and only when the fused grid still gives >= 2 waves (measured: with fewer tiles the lost
    // parallelism and the single-buffered accumulators cost more than the saved A traffic)
    // weight-stationary ConvTranspose kernel: a (phase, N tile) weight set of <= 96 KB resident per CTA
    p.wstat = 0;
    p.upc = upc ? 1 : 0; p.taps2 = a->in2 ? a->taps2 : 0;
    if (upc) {
      // N tile: fewest (waves x cycles per K chunk) on the 74 clusters; an M256 x N MMA of one K chunk costs ~512 / 384 /
      // 256 cycles at N = 256 / 128 / 64 (narrow tiles are bound by the shared-memory operand path, conv_gemm.cu header)
      const int npairs = (p.tiles_h * p.tiles_w + 1) / 2, ncl = std::max(1, device_sm_count() / 2);
      int best = 0; long best_t = -1;
      for (int bn : {256, 128, 64}) {
        if (a->bn > 0 ? bn != a->bn : bn > (a->Cout + 63) / 64 * 64) continue;
        const int tn = (a->Cout + bn - 1) / bn;
        const long waves = (static_cast<long>(npairs) * 4 * tn + ncl - 1) / ncl;
        const long t = waves * (bn == 256 ? 512 : bn == 128 ? 384 : 256);
        if (best_t < 0 || t < best_t) { best_t = t; best = bn; }
      }
      if (!best) { vpb_set_error("conv: upconv N tile must be 64, 128 or 256 (bn=%d)", a->bn); return VPB_ERR_ARG; }
      p.BN = best; p.tiles_n = (a->Cout + best - 1) / best;
    }
    if (a->phases == 4 && !upc && !split && a->bn <= 0 && a->mode == VPB_EPI_STORE && (a->act == ACT_NONE || a->act == ACT_GELU) &&
        a->dbg_gb != 2 && a->dbg_ms == 0) {
      // CTA pair: each CTA holds HALF of the set's rows; widest N tile whose half set fits next to the slabs and a ring of
      // >= 3 activation tiles, with at least one pixel-tile pair per cluster
      const int nB = p.kchunks + p.kchunks2, npairs = (p.tiles_h * p.tiles_w + 1) / 2, ncl = std::max(1, device_sm_count() / 2);
      for (int bn : {256, 128}) {
        if (bn > (a->Cout + 127) / 128 * 128 || a->dbg_pair < 0) continue;
        const size_t wbytes = static_cast<size_t>(nB) * (bn / 2) * 128, slabs = static_cast<size_t>(bn / 64) * 128 * 128;
        const int tn = (a->Cout + bn - 1) / bn, S = 4 * tn;
        const long ring = static_cast<long>(kMaxDynSmem) - 1024 - static_cast<long>(wbytes) - static_cast<long>(slabs);
        if (ring < 3 * kATileBytes || S > ncl) continue;
        if (static_cast<long>(npairs) * S < ncl) continue;
        // measured per layer (gpurun_out/r2l_bench_conv.txt, tile kernel -> this kernel): upsample_layer_1 16.9 -> 14.6 us,
        // _3 17.3 -> 12.9, _4 23.1 -> 18.3, but _2 (512 -> 512 at 40x80: 9 K chunks against a 7-slot ring, 3 rounds of
        // 36-MMA tiles) 16.7 -> 18.3: long K loops with many pixel-tile pairs stay on the tile kernel
        if (nB >= 8 && npairs >= 8) continue;
        p.wstat = 1; p.BN = bn; p.tiles_n = tn;
        p.stages = static_cast<int>(std::min<long>(kMaxStages, ring / kATileBytes));
        break;
      }
    }
    p.fuse4 = (!p.wstat && !p.upc && a->phases == 4 && p.BN <= 128 && a->dbg_ms != 1 && !a->in2 && !split &&
               (p.tiles_h * p.tiles_w * p.tiles_n >= 2 * device_sm_count() || a->dbg_ms == 2)) ? 1 : 0;
    p.total_tiles = p.tiles_h * p.tiles_w * p.tiles_n * (p.fuse4 ? 1 : p.phases);
    const size_t stage_bytes = kATileBytes + b_bytes * (p.fuse4 ? 4 : 1);
    // staged TMA-store epilogue: plain store (ConvTranspose, with or without the fused skip link) or GELU, N tile made
    // of whole 64-channel slabs; dbg_gb == 2 forces the direct-store epilogue (A/B comparison)
    p.tma_store = (p.wstat || (p.upc && (a->dbg_gb == 3 || a->Cout % p.BN)) || (!split && a->mode == VPB_EPI_STORE && (a->act == ACT_NONE || a->act == ACT_GELU) && p.BN % 64 == 0 &&
                               a->phases == 4 && a->dbg_gb != 2)) ? 1 : 0;
    const size_t slab_bytes = p.tma_store ? static_cast<size_t>(p.BN / 64) * 128 * 128 : 0;
    int stages = static_cast<int>((kMaxDynSmem - 1024 - slab_bytes) / stage_bytes);
    if (stages < 2 && p.tma_store && !p.wstat && !p.upc) { p.tma_store = 0; stages = static_cast<int>((kMaxDynSmem - 1024) / stage_bytes); }
    if (p.upc) {
      const size_t st = kATileBytes + static_cast<size_t>(p.BN / 2) * 128, bias_b = (static_cast<size_t>(p.tiles_n) * p.BN * 4 + 1023) / 1024 * 1024;
      const long ring = static_cast<long>(kMaxDynSmem) - 1024 - static_cast<long>(slab_bytes) - static_cast<long>(bias_b);
      p.stages = static_cast<int>(std::min<long>(kUpcStages, ring / static_cast<long>(st)));
      if (p.stages < 3) { vpb_set_error("conv: upconv ring does not fit"); return VPB_ERR_ARG; }
      plan->smem_bytes = p.stages * st + slab_bytes + bias_b + 1024;
      p.total_tiles = ((p.tiles_h * p.tiles_w + 1) / 2) * p.tiles_n * 4;      // pixel-tile pairs x (phase, N tile)
    } else if (p.wstat) {
      plan->smem_bytes = static_cast<size_t>(p.stages) * kATileBytes + static_cast<size_t>(p.kchunks + p.kchunks2) * b_bytes_ws(p.BN / 2) +
                         static_cast<size_t>(p.BN / 64) * 128 * 128 + 1024;
      p.total_tiles = ((p.tiles_h * p.tiles_w + 1) / 2) * p.tiles_n * 4;      // pixel-tile pairs x weight sets
    } else {
      p.stages = std::max(2, std::min(stages, kMaxStages));
      plan->smem_bytes = p.stages * stage_bytes + (p.tma_store ? slab_bytes : 0) + 1024;
    }
  }
  p.mg_tn = fast_div_magic(p.tiles_n); p.mg_tw = fast_div_magic(lin ? 1 : p.tiles_w);
  p.mg_tpp = fast_div_magic(lin ? 1 : p.upc ? (p.tiles_h * p.tiles_w + 1) / 2 : p.tiles_n * p.tiles_h * p.tiles_w); p.mg_wp = fast_div_magic(p.WP);
  p.act = a->act; p.mode = a->mode; p.final_kind = a->final_kind;
  p.bias = a->bias; p.out = a->out; p.ldo = a->ldo; p.res = a->res; p.ldr = a->ldr;
  p.out_f32 = a->out_f32; p.out_cls = a->out_cls;
  p.trace = a->dbg_trace;
  plan->dtype = a->dtype;
  plan->grid = (p.wstat || p.upc) ? 2 * std::min(p.total_tiles, device_sm_count() / 2)
             : p.splitk ? p.total_tiles * p.splitk
             : p.pair ? 2 * std::min(p.total_tiles, device_sm_count() / 2)
                      : std::min(p.total_tiles, device_sm_count());
  plan->flops = 2.0 * a->H * a->W * static_cast<double>(a->Cout) * a->phases * (a->Cin * a->taps + p.Cin2 * (upc ? 9 : 1));

  const CUtensorMapDataType dt =
      a->dtype == VPB_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUresult r;
  if (lin) {
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(a->Cin), static_cast<cuuint64_t>(p.NP)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(a->ldi) * 2};
    cuuint32_t box[2] = {64, static_cast<cuuint32_t>(kSegRows)};
    cuuint32_t es[2] = {1, 1};
    r = enc(&plan->mapA, dt, 2, const_cast<void*>(a->in), dims, strides, box, es,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    // a zero-bordered input is addressed through its interior: base at pixel (1,1), padded pitch
    auto encA = [&](const void* ptr, CUtensorMap* m) {
      const int pad = p.in_pad;
      const int Hin = (cstride == 2 && a->in_h > 0) ? a->in_h : a->H, Win = (cstride == 2 && a->in_w > 0) ? a->in_w : a->W;
      const size_t pitch = static_cast<size_t>(Win + 2 * pad) * a->ldi * 2;
      const uint8_t* base = static_cast<const uint8_t*>(ptr) + pad * pitch + static_cast<size_t>(pad) * a->ldi * 2;
      cuuint64_t dims[4] = {static_cast<cuuint64_t>(a->Cin), static_cast<cuuint64_t>(Win),
                            static_cast<cuuint64_t>(Hin), 1};
      cuuint64_t strides[3] = {static_cast<cuuint64_t>(a->ldi) * 2, pitch, pitch * (Hin + 2 * pad)};
      // stride 2: the box spans 2*TW x 2*TH input pixels and is traversed with element stride 2 (TW x TH loaded)
      cuuint32_t box[4] = {64, static_cast<cuuint32_t>(p.TW * cstride), static_cast<cuuint32_t>(p.TH * cstride), 1};
      cuuint32_t es[4] = {1, static_cast<cuuint32_t>(cstride), static_cast<cuuint32_t>(cstride), 1};
      return enc(m, dt, 4, const_cast<uint8_t*>(base), dims, strides, box, es,
                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                 CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    };
    r = encA(a->in, &plan->mapA);
    if (r == CUDA_SUCCESS && split) r = encA(a->in_lo, &plan->mapAlo);
  }
  if (r != CUDA_SUCCESS) {
    vpb_set_error("conv: cuTensorMapEncodeTiled(A) failed: %d", static_cast<int>(r));
    return VPB_ERR_CUDA;
  }
  {
    auto encB = [&](const void* ptr, CUtensorMap* m) {
      const size_t ldw = a->ldw > 0 ? a->ldw : a->Cin;
      cuuint64_t dims[3] = {static_cast<cuuint64_t>(a->Cin), static_cast<cuuint64_t>(a->Cout),
                            static_cast<cuuint64_t>(a->taps * a->phases)};
      cuuint64_t strides[2] = {static_cast<cuuint64_t>(ldw) * 2, static_cast<cuuint64_t>(ldw) * 2 * a->Cout};
      cuuint32_t box[3] = {64, static_cast<cuuint32_t>((p.pair || p.wstat || p.upc) ? p.BN / 2 : p.BN), 1};
      cuuint32_t es[3] = {1, 1, 1};
      return enc(m, dt, 3, const_cast<void*>(ptr), dims, strides, box, es,
                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    };
    r = encB(a->w, &plan->mapB);
    if (r == CUDA_SUCCESS && split) r = encB(a->w_lo, &plan->mapBlo);
    if (r != CUDA_SUCCESS) {
      vpb_set_error("conv: cuTensorMapEncodeTiled(B) failed: %d", static_cast<int>(r));
      return VPB_ERR_CUDA;
    }
  }
  plan->mapA2 = plan->mapA;
  plan->mapB2 = plan->mapB;
  plan->mapO = plan->mapA;
  if (!split) { plan->mapAlo = plan->mapA; plan->mapBlo = plan->mapB; }
  plan->mapA2lo = plan->mapAlo;
  plan->mapB2lo = plan->mapBlo;
  if (p.tma_store) {
    // output viewed as [h][a][w][b][c] (ConvTranspose phase (a,b) of pixel tile (h0,w0) is one tiled box)
    const int s2 = a->phases == 4 ? 2 : 1;
    const int pad = p.out_pad;
    const size_t px = static_cast<size_t>(a->ldo) * 2;
    const size_t pitch = static_cast<size_t>(a->W * s2 + 2 * pad) * px;
    uint8_t* base = static_cast<uint8_t*>(a->out) + pad * pitch + pad * px;
    cuuint64_t dims[5] = {static_cast<cuuint64_t>(p.nlim), static_cast<cuuint64_t>(s2), static_cast<cuuint64_t>(a->W),
                          static_cast<cuuint64_t>(s2), static_cast<cuuint64_t>(a->H)};
    cuuint64_t strides[4] = {px, px * s2, pitch, pitch * s2};
    cuuint32_t box[5] = {64, 1, static_cast<cuuint32_t>(p.TW), 1, static_cast<cuuint32_t>(p.TH)};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    r = enc(&plan->mapO, dt, 5, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      vpb_set_error("conv: cuTensorMapEncodeTiled(O) failed: %d", static_cast<int>(r));
      return VPB_ERR_CUDA;
    }
  }
  if (a->in2) {
    // second input at output resolution, viewed as [h][a][w][b][c] (s = 2 for a ConvTranspose,
    // a = b = 0 and s = 1 otherwise); box = one pixel tile of one phase
    auto encA2 = [&](const void* ptr, CUtensorMap* m) {
      const int s2 = a->phases == 4 ? 2 : 1;
      const int pad = a->in2_pad ? 1 : 0;
      const size_t px = static_cast<size_t>(a->ld2) * 2;                       // bytes per pixel
      const size_t pitch = static_cast<size_t>(a->W * s2 + 2 * pad) * px;      // bytes per image row
      const uint8_t* base = static_cast<const uint8_t*>(ptr) + pad * pitch + pad * px;
      cuuint64_t dims[5] = {static_cast<cuuint64_t>(a->Cin2), static_cast<cuuint64_t>(s2),
                            static_cast<cuuint64_t>(a->W), static_cast<cuuint64_t>(s2),
                            static_cast<cuuint64_t>(a->H)};
      cuuint64_t strides[4] = {px, px * s2, pitch, pitch * s2};
      cuuint32_t box[5] = {64, 1, static_cast<cuuint32_t>(p.TW), 1, static_cast<cuuint32_t>(p.TH)};
      cuuint32_t es[5] = {1, 1, 1, 1, 1};
      return enc(m, dt, 5, const_cast<uint8_t*>(base), dims, strides, box, es,
                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                 CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    };
    r = encA2(a->in2, &plan->mapA2);
    if (r == CUDA_SUCCESS && split) r = encA2(a->in2_lo, &plan->mapA2lo);
    if (r != CUDA_SUCCESS) {
      vpb_set_error("conv: cuTensorMapEncodeTiled(A2) failed: %d", static_cast<int>(r));
      return VPB_ERR_CUDA;
    }
    auto encB2 = [&](const void* ptr, CUtensorMap* m) {
      cuuint64_t bd[3] = {static_cast<cuuint64_t>(a->Cin2), static_cast<cuuint64_t>(a->Cout), static_cast<cuuint64_t>(p.upc ? 9 : 1)};
      cuuint64_t bs[2] = {static_cast<cuuint64_t>(a->Cin2) * 2, static_cast<cuuint64_t>(a->Cin2) * 2 * a->Cout};
      cuuint32_t bb[3] = {64, static_cast<cuuint32_t>((p.wstat || p.upc) ? p.BN / 2 : p.BN), 1};
      cuuint32_t be[3] = {1, 1, 1};
      return enc(m, dt, 3, const_cast<void*>(ptr), bd, bs, bb, be,
                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    };
    r = encB2(a->w2, &plan->mapB2);
    if (r == CUDA_SUCCESS && split) r = encB2(a->w2_lo, &plan->mapB2lo);
    if (r != CUDA_SUCCESS) {
      vpb_set_error("conv: cuTensorMapEncodeTiled(B2) failed: %d", static_cast<int>(r));
      return VPB_ERR_CUDA;
    }
  }
  return VPB_OK;
}

int conv_plan_launch(const ConvPlan* plan, cudaStream_t stream) {
  {
    std::lock_guard<std::mutex> g(init_mutex());
    bool* done = device_flag(kInitConv);
    if (!*done) {
      VPB_CUDA_OK(cudaFuncSetAttribute(conv_gemm_kernel<F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
      VPB_CUDA_OK(cudaFuncSetAttribute(conv_gemm_kernel<BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
      VPB_CUDA_OK(cudaFuncSetAttribute(conv3x3_lin_kernel<F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
      VPB_CUDA_OK(cudaFuncSetAttribute(conv3x3_lin_kernel<BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
      VPB_CUDA_OK(cudaFuncSetAttribute(conv3x3_pair_kernel<F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
      VPB_CUDA_OK(cudaFuncSetAttribute(conv3x3_pair_kernel<BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
      VPB_CUDA_OK(cudaFuncSetAttribute(conv3x3_splitk_kernel<F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
      VPB_CUDA_OK(cudaFuncSetAttribute(conv3x3_splitk_kernel<BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
      VPB_CUDA_OK(cudaFuncSetAttribute(convt_ws_kernel<F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
      VPB_CUDA_OK(cudaFuncSetAttribute(convt_ws_kernel<BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
      VPB_CUDA_OK(cudaFuncSetAttribute(upconv_pair_kernel<F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
      VPB_CUDA_OK(cudaFuncSetAttribute(upconv_pair_kernel<BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
      *done = true;
    }
  }
  const bool bf = plan->dtype == VPB_BF16;
  const dim3 g(plan->grid), b(kThreads);
  if (plan->p.lin && plan->p.splitk) {
    if (bf) VPB_CUDA_OK(launch_k_cluster(conv3x3_splitk_kernel<BF16>, g, b, plan->smem_bytes, stream, plan->p.splitk, plan->mapA, plan->mapB, plan->p));
    else VPB_CUDA_OK(launch_k_cluster(conv3x3_splitk_kernel<F16>, g, b, plan->smem_bytes, stream, plan->p.splitk, plan->mapA, plan->mapB, plan->p));
  } else if (plan->p.lin && plan->p.pair) {
    if (bf) VPB_CUDA_OK(launch_k(conv3x3_pair_kernel<BF16>, g, b, plan->smem_bytes, stream, plan->mapA, plan->mapB, plan->p));
    else VPB_CUDA_OK(launch_k(conv3x3_pair_kernel<F16>, g, b, plan->smem_bytes, stream, plan->mapA, plan->mapB, plan->p));
  } else if (plan->p.lin) {
    if (bf) VPB_CUDA_OK(launch_k(conv3x3_lin_kernel<BF16>, g, b, plan->smem_bytes, stream, plan->mapA, plan->mapB, plan->p));
    else VPB_CUDA_OK(launch_k(conv3x3_lin_kernel<F16>, g, b, plan->smem_bytes, stream, plan->mapA, plan->mapB, plan->p));
  } else {
    ConvMaps maps;
    maps.A = plan->mapA; maps.B = plan->mapB; maps.A2 = plan->mapA2; maps.B2 = plan->mapB2; maps.O = plan->mapO;
    maps.Alo = plan->mapAlo; maps.Blo = plan->mapBlo; maps.A2lo = plan->mapA2lo; maps.B2lo = plan->mapB2lo;
    if (plan->p.upc) {
      if (bf) VPB_CUDA_OK(launch_k(upconv_pair_kernel<BF16>, g, b, plan->smem_bytes, stream, maps, plan->p));
      else VPB_CUDA_OK(launch_k(upconv_pair_kernel<F16>, g, b, plan->smem_bytes, stream, maps, plan->p));
    } else if (plan->p.wstat) {
      if (bf) VPB_CUDA_OK(launch_k(convt_ws_kernel<BF16>, g, b, plan->smem_bytes, stream, maps, plan->p));
      else VPB_CUDA_OK(launch_k(convt_ws_kernel<F16>, g, b, plan->smem_bytes, stream, maps, plan->p));
    } else if (bf) VPB_CUDA_OK(launch_k(conv_gemm_kernel<BF16>, g, b, plan->smem_bytes, stream, maps, plan->p));
    else VPB_CUDA_OK(launch_k(conv_gemm_kernel<F16>, g, b, plan->smem_bytes, stream, maps, plan->p));
  }
  return VPB_OK;
}

}  // namespace vpb

extern "C" int vpb_conv_gemm(const vpb_conv_args* a, void* stream) {
  vpb::ConvPlan plan;
  int rc = vpb::conv_plan_build(a, &plan);
  if (rc != VPB_OK) return rc;
  return vpb::conv_plan_launch(&plan, static_cast<cudaStream_t>(stream));
}
