// conv_gemm.cu — implicit-GEMM convolution on Blackwell tcgen05 tensor cores.
//
// Replaces (reference, all fp32 library calls):
//   nn.Conv2d 3x3 s1 p1   Models/model_components/scene_neck.py:13-24, scene_seg_head.py:13-19,
//                         scene_3d_head.py:13-20, ego_lanes_head.py:13-15, scene_context.py:20-22
//   nn.Conv2d 1x1         scene_neck.py:12,17,22 (skip links) and EfficientNet-B0 pointwise convs
//   nn.ConvTranspose2d k2 s2   scene_neck.py:11,16,21, scene_seg_head.py:11,16
//
// Formulation.  Activations are NHWC 16-bit.  For an output tile of 128 pixels
// (a TH x TW spatial patch) and BN output channels,
//     D[pixel, n] = sum_{tap} sum_{c} In[pixel + offset(tap), c] * W[tap][n][c]
// is a GEMM with M = 128, N = BN, K = taps * Cin.  The A operand for one (tap, 64-channel
// chunk) is a single 4-D TMA box {64 ch, TW, TH, 1} at the shifted coordinate; out-of-range
// rows/columns are zero-filled by the TMA unit, which IS the convolution's zero padding.
// The box lands in shared memory as 128 rows x 128 B with the 128-byte swizzle, exactly
// the K-major layout tcgen05.mma consumes.  B (weights) is a {64, BN, 1} box of the
// [tap][Cout][Cin] tensor.  Accumulators live in TMEM (fp32), double-buffered so the
// epilogue of tile i overlaps the main loop of tile i+1.
//
// Warp roles (320 threads, persistent, 1 CTA / SM):
//   warp 0      TMA producer (one elected lane)
//   warp 1      tcgen05.mma issuer (one elected lane)
//   warps 2..9  epilogue: tcgen05.ld -> bias/activation/residual -> 16-bit NHWC stores
//               (or fp32 planar logits + class map for the heads' last conv).
#include "common.cuh"
#include "conv_gemm.cuh"
#include <cstdio>
#include <cstring>
#include <algorithm>

namespace vpb {

static constexpr int kThreads = 320;
static constexpr int kEpiWarps = 8;
static constexpr int kMaxStages = 8;
static constexpr int kATileBytes = 128 * 128;  // 128 pixels x 64 ch x 2 B
static constexpr int kAccStride = 256;         // TMEM columns between the two accumulators
// 227 KB opt-in limit covers static + dynamic shared memory; keep 2 KB for the static part.
static constexpr int kMaxDynSmem = 227 * 1024 - 2048;

template <class E>
__global__ void __launch_bounds__(kThreads, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap mapA,
                 const __grid_constant__ CUtensorMap mapB, const ConvKParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[kMaxStages];
  __shared__ __align__(8) uint64_t bar_empty[kMaxStages];
  __shared__ __align__(8) uint64_t bar_tfull[2];
  __shared__ __align__(8) uint64_t bar_tempty[2];
  __shared__ uint32_t tmem_holder;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t b_tile_bytes = static_cast<uint32_t>(p.BN) * 128u;
  const uint32_t stage_bytes = kATileBytes + b_tile_bytes;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapA);
    tma_prefetch_desc(&mapB);
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

  const int kiters = p.taps * p.kchunks;
  const int tiles_per_phase = p.tiles_n * p.tiles_h * p.tiles_w;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int ph = tile / tiles_per_phase;
        int r = tile - ph * tiles_per_phase;
        const int nt = r % p.tiles_n;
        r /= p.tiles_n;
        const int twi = r % p.tiles_w;
        const int thi = r / p.tiles_w;
        const int h0 = thi * p.TH, w0 = twi * p.TW, n0 = nt * p.BN;
        for (int t = 0; t < p.taps; ++t) {
          const int dy = (p.taps == 9) ? (t / 3 - 1) : 0;
          const int dx = (p.taps == 9) ? (t % 3 - 1) : 0;
          const int wsel = (p.phases > 1) ? ph : t;
          for (int c = 0; c < p.kchunks; ++c) {
            mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1u);
            const uint32_t full = smem_u32(&bar_full[stage]);
            const uint32_t sa = smem_base + stage * stage_bytes;
            mbar_arrive_expect_tx(full, stage_bytes);
            tma_load_4d(sa, &mapA, full, c * 64, w0 + dx, h0 + dy, 0);
            tma_load_3d(sa + kATileBytes, &mapB, full, c * 64, n0, wsel);
            if (++stage == p.stages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = umma_idesc(E::kUmmaFmt, 128, p.BN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(smem_u32(&bar_tempty[as]), aphase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * kAccStride;
        for (int k = 0; k < kiters; ++k) {
          const int c = k % p.kchunks;
          const int kvalid = min(64, p.Cin - c * 64);
          const int ksteps = (kvalid + 15) >> 4;
          mbar_wait(smem_u32(&bar_full[stage]), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * stage_bytes;
          const uint64_t adesc = umma_desc_k128(sa);
          const uint64_t bdesc = umma_desc_k128(sa + kATileBytes);
          for (int kk = 0; kk < ksteps; ++kk) {
            // +32 B along K inside the 128-B swizzle row == +2 in the encoded start address
            umma_f16(d_tmem, adesc + 2 * kk, bdesc + 2 * kk, idesc, (k | kk) != 0 ? 1u : 0u);
          }
          umma_commit(smem_u32(&bar_empty[stage]));
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
        umma_commit(smem_u32(&bar_tfull[as]));
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue
    const int q = warp & 3;           // TMEM lane quadrant this warp may read
    const int half = (warp - 2) >> 2; // which interleaved set of 16-column chunks
    const int row = q * 32 + lane;    // accumulator row == pixel within the tile
    const int lh = row >> p.tw_shift;
    const int lw = row & (p.TW - 1);
    const int nchunks = p.BN >> 4;
    const int Ho = (p.phases > 1) ? 2 * p.H : p.H;
    const int Wo = (p.phases > 1) ? 2 * p.W : p.W;
    (void)Ho;
    typename E::T* out = reinterpret_cast<typename E::T*>(p.out);
    const typename E::T* res = reinterpret_cast<const typename E::T*>(p.res);
    int it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int ph = tile / tiles_per_phase;
      int r = tile - ph * tiles_per_phase;
      const int nt = r % p.tiles_n;
      r /= p.tiles_n;
      const int twi = r % p.tiles_w;
      const int thi = r / p.tiles_w;
      const int h = thi * p.TH + lh, w = twi * p.TW + lw, n0 = nt * p.BN;
      const bool pix_ok = (h < p.H) && (w < p.W);
      const int oh = (p.phases > 1) ? 2 * h + (ph >> 1) : h;
      const int ow = (p.phases > 1) ? 2 * w + (ph & 1) : w;
      const size_t opix = static_cast<size_t>(oh) * Wo + ow;

      mbar_wait(smem_u32(&bar_tfull[as]), aphase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * kAccStride;

      for (int chunk = half; chunk < nchunks; chunk += 2) {
        uint32_t rr[16];
        tmem_ld16(t_row + chunk * 16, rr);
        tmem_ld_wait();
        const int n = n0 + chunk * 16;
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int nn = n + i;
          float x = __uint_as_float(rr[i]);
          if (nn < p.Cout) {
            if (p.bias) x += __ldg(p.bias + nn);
            x = apply_act(x, p.act);
          } else {
            x = 0.f;
          }
          v[i] = x;
        }
        if (p.mode == VPB_EPI_FINAL) {
          if (pix_ok && chunk == 0) {
            const size_t plane = static_cast<size_t>(p.H) * p.W;
            const size_t pix = static_cast<size_t>(h) * p.W + w;
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (i < p.Cout) p.out_f32[i * plane + pix] = v[i];
            if (p.out_cls) {
              uint8_t cls = 0;
              if (p.final_kind == VPB_FINAL_ARGMAX) {
                float best = v[0];
#pragma unroll
                for (int i = 1; i < 16; ++i)
                  if (i < p.Cout && v[i] > best) { best = v[i]; cls = static_cast<uint8_t>(i); }
              } else if (p.final_kind == VPB_FINAL_THRESH) {
                cls = v[0] > 0.f ? 1 : 0;
              } else if (p.final_kind == VPB_FINAL_EGOLANES) {
                cls = (v[2] > 0.f) ? 2 : (v[1] > 0.f) ? 1 : (v[0] > 0.f) ? 0 : 255;
              }
              p.out_cls[pix] = cls;
            }
          }
        } else if (pix_ok) {
          if (p.mode == VPB_EPI_ADD || p.mode == VPB_EPI_MULADD) {
            const typename E::T* rp = res + opix * p.ldr + n;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              if (n + 8 * j < p.ldr && n + 8 * j < p.ldo) {
                const uint4 rv = *reinterpret_cast<const uint4*>(rp + 8 * j);
                const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const float2 f = unpack2<E>(rw[i]);
                  float& a = v[8 * j + 2 * i];
                  float& b = v[8 * j + 2 * i + 1];
                  if (p.mode == VPB_EPI_ADD) { a += f.x; b += f.y; }
                  else { a = a * f.x + f.x; b = b * f.y + f.y; }
                }
              }
            }
          }
          typename E::T* op = out + opix * p.ldo + n;
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            if (n + 8 * j < p.ldo) {
              uint4 o;
              o.x = pack2<E>(v[8 * j + 0], v[8 * j + 1]);
              o.y = pack2<E>(v[8 * j + 2], v[8 * j + 3]);
              o.z = pack2<E>(v[8 * j + 4], v[8 * j + 5]);
              o.w = pack2<E>(v[8 * j + 6], v[8 * j + 7]);
              *reinterpret_cast<uint4*>(op + 8 * j) = o;
            }
          }
        }
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
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}

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
  if (!((a->taps == 9 && a->phases == 1) || (a->taps == 1 && (a->phases == 1 || a->phases == 4)))) {
    vpb_set_error("conv: unsupported taps=%d phases=%d", a->taps, a->phases);
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
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) {
    vpb_set_error("conv: cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
    return VPB_ERR_CUDA;
  }
  memset(plan, 0, sizeof(*plan));
  ConvKParams& p = plan->p;
  p.H = a->H; p.W = a->W; p.Cin = a->Cin; p.Cout = a->Cout;
  p.taps = a->taps; p.phases = a->phases;
  // spatial tile: minimise padded pixels, prefer wide tiles
  int best_tw = 128; long best_cost = -1;
  for (int tw = 128; tw >= 8; tw >>= 1) {
    const int th = 128 / tw;
    const long cost = static_cast<long>((a->H + th - 1) / th) * ((a->W + tw - 1) / tw);
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_tw = tw; }
  }
  p.TW = best_tw; p.TH = 128 / best_tw;
  p.tw_shift = 0; while ((1 << p.tw_shift) < p.TW) ++p.tw_shift;
  p.BN = a->bn > 0 ? a->bn : pick_bn(a->Cout);
  if (p.BN % 16 || p.BN > 256 || p.BN < 16) {
    vpb_set_error("conv: bad BN %d", p.BN);
    return VPB_ERR_ARG;
  }
  p.tiles_h = (a->H + p.TH - 1) / p.TH;
  p.tiles_w = (a->W + p.TW - 1) / p.TW;
  p.tiles_n = (a->Cout + p.BN - 1) / p.BN;
  p.total_tiles = p.tiles_h * p.tiles_w * p.tiles_n * p.phases;
  p.kchunks = (a->Cin + 63) / 64;
  const size_t stage_bytes = kATileBytes + static_cast<size_t>(p.BN) * 128;
  int stages = static_cast<int>((kMaxDynSmem - 1024) / stage_bytes);
  p.stages = std::max(2, std::min(stages, kMaxStages));
  p.act = a->act; p.mode = a->mode; p.final_kind = a->final_kind;
  p.bias = a->bias; p.out = a->out; p.ldo = a->ldo; p.res = a->res; p.ldr = a->ldr;
  p.out_f32 = a->out_f32; p.out_cls = a->out_cls;
  plan->dtype = a->dtype;
  plan->smem_bytes = p.stages * stage_bytes + 1024;
  plan->grid = std::min(p.total_tiles, device_sm_count());
  plan->flops = 2.0 * a->H * a->W * static_cast<double>(a->Cout) * a->Cin * a->taps * a->phases;

  const CUtensorMapDataType dt =
      a->dtype == VPB_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  {
    cuuint64_t dims[4] = {static_cast<cuuint64_t>(a->Cin), static_cast<cuuint64_t>(a->W),
                          static_cast<cuuint64_t>(a->H), 1};
    cuuint64_t strides[3] = {static_cast<cuuint64_t>(a->ldi) * 2,
                             static_cast<cuuint64_t>(a->ldi) * 2 * a->W,
                             static_cast<cuuint64_t>(a->ldi) * 2 * a->W * a->H};
    cuuint32_t box[4] = {64, static_cast<cuuint32_t>(p.TW), static_cast<cuuint32_t>(p.TH), 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(&plan->mapA, dt, 4, const_cast<void*>(a->in), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      vpb_set_error("conv: cuTensorMapEncodeTiled(A) failed: %d", static_cast<int>(r));
      return VPB_ERR_CUDA;
    }
  }
  {
    cuuint64_t dims[3] = {static_cast<cuuint64_t>(a->Cin), static_cast<cuuint64_t>(a->Cout),
                          static_cast<cuuint64_t>(a->taps * a->phases)};
    cuuint64_t strides[2] = {static_cast<cuuint64_t>(a->Cin) * 2,
                             static_cast<cuuint64_t>(a->Cin) * 2 * a->Cout};
    cuuint32_t box[3] = {64, static_cast<cuuint32_t>(p.BN), 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&plan->mapB, dt, 3, const_cast<void*>(a->w), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      vpb_set_error("conv: cuTensorMapEncodeTiled(B) failed: %d", static_cast<int>(r));
      return VPB_ERR_CUDA;
    }
  }
  return VPB_OK;
}

int conv_plan_launch(const ConvPlan* plan, cudaStream_t stream) {
  static bool attr_set[2] = {false, false};
  const int di = plan->dtype == VPB_BF16 ? 1 : 0;
  if (!attr_set[di]) {
    if (di == 0)
      VPB_CUDA_OK(cudaFuncSetAttribute(conv_gemm_kernel<F16>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
    else
      VPB_CUDA_OK(cudaFuncSetAttribute(conv_gemm_kernel<BF16>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
    attr_set[di] = true;
  }
  if (di == 0)
    conv_gemm_kernel<F16><<<plan->grid, kThreads, plan->smem_bytes, stream>>>(plan->mapA, plan->mapB,
                                                                               plan->p);
  else
    conv_gemm_kernel<BF16><<<plan->grid, kThreads, plan->smem_bytes, stream>>>(plan->mapA,
                                                                                plan->mapB, plan->p);
  VPB_CUDA_OK(cudaGetLastError());
  return VPB_OK;
}

}  // namespace vpb

extern "C" int vpb_conv_gemm(const vpb_conv_args* a, void* stream) {
  vpb::ConvPlan plan;
  int rc = vpb::conv_plan_build(a, &plan);
  if (rc != VPB_OK) return rc;
  return vpb::conv_plan_launch(&plan, static_cast<cudaStream_t>(stream));
}
