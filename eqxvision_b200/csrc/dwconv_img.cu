// Depthwise KxK convolution + folded BatchNorm + activation with the SqueezeExcitation squeeze fused in
// (K3 + K11 of SURVEY.md 2.1): reference `_MBConv` (efficientnet.py:138-160: depthwise ConvNormActivation, then
// SqueezeExcitation whose first step is `avgpool_1x1(x)`, layers/squeeze.py:52) and `_InvertedResidual`
// (mobilenetv3.py:88-112).
//
// Why a second depthwise kernel: the register-strip kernel in depthwise.cu is ISSUE bound, not HBM bound. Per output
// element a 3x3 filter needs 9 fp32 FMAs and a 5x5 filter 25, against 4 bytes of traffic: at 6.4 TB/s that is
// 14 / 40 TFMA/s of scalar FFMA on a machine that issues ~18 TFMA/s of them, before a single load, unpack or address
// instruction. This kernel
//   * issues the multiply-adds as packed `fma.rn.f32x2` (two channels per instruction: half the FMA issue slots),
//   * computes TW x TH outputs per thread so that a loaded + unpacked input vector feeds up to K*TH... taps,
//   * keeps the block's filter slab in shared memory (one LDS.128 per 4 channels of a tap, no global re-reads),
//   * evaluates SiLU with ONE special-function op (x * (0.5 + 0.5 * tanh(x/2))) instead of exp + reciprocal, and
//   * assigns a block to ONE image and a slice of its rows, so the per-(image, channel) sums the SE block needs fall
//     out of the registers: block partial sums go to a small workspace, the last block of an image (integer ticket)
//     adds them in a fixed order and writes the pooled mean. No floating-point atomics: the result is bitwise
//     reproducible and independent of batch composition, and the separate global-average-pool pass over the widest
//     tensor of every MBConv block disappears.
#include <algorithm>
#include <cstdlib>

#include "common.h"
#include "ptx.cuh"

namespace eqxv {

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ unsigned long long pack_f2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack_f2(unsigned long long v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
// two bf16 in one 32-bit word -> packed fp32 pair (exact)
__device__ __forceinline__ unsigned long long bf2_to_f2(uint32_t u) {
  return pack_f2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}

template <int ACT>
__device__ __forceinline__ float act_ct(float v) {
  if (ACT == EQXV_ACT_RELU) return fmaxf(v, 0.f);
  if (ACT == EQXV_ACT_SILU) {
    const float h = 0.5f * v;
    float th;
    asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(h));
    return fmaf(h, th, h);
  }
  if (ACT == EQXV_ACT_HARDSWISH) return v * fminf(fmaxf(v + 3.f, 0.f), 6.f) * (1.f / 6.f);
  return v;
}

struct DwImgParams {
  const __nv_bfloat16* x;
  const float* wgt;    // [K*K][wp]
  const float* bias;   // [>= c]
  __nv_bfloat16* y;
  float* partial;      // [n][rchunks][c]   (POOL, rchunks > 1)
  int* tickets;        // [n][gchunks]      (POOL, rchunks > 1), zero on entry, zero on exit
  __nv_bfloat16* pooled;   // [n][pool_pitch]
  int h, w, c, pad, ho, wo, xp, yp, wp, pool_pitch;
  int gc;              // channel groups (of 8) per block
  int tile_rows;       // output-row tiles (of TH rows) per block
  int rchunks;
  float inv_area;
};

template <int NW>
struct Words {
  uint32_t w[NW];
};
template <int NW>
__device__ __forceinline__ Words<NW> ldg_words(const __nv_bfloat16* p) {
  Words<NW> r;
  if (NW == 4) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
    r.w[0] = v.x, r.w[1] = v.y, r.w[2 % NW] = v.z, r.w[3 % NW] = v.w;
  } else {
    const uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
    r.w[0] = v.x, r.w[1] = v.y;
  }
  return r;
}

// CV = channels per thread (8: 16-byte vectors, 4: 8-byte vectors - half the registers, twice the resident warps)
template <int K, int S, int TW, int TH, int CV, int ACT, bool POOL>
__global__ void __launch_bounds__(kThreads, (CV == 4 ? 3 : (TH == 1 ? 2 : 1))) dwconv_img_kernel(const DwImgParams p) {
  constexpr int NC = (TW - 1) * S + K;   // input columns one thread reads per input row
  constexpr int NR = (TH - 1) * S + K;   // input rows one thread walks over
  constexpr int NW = CV / 2;             // 32-bit words (bf16 pairs) per vector
  extern __shared__ float4 dwi_smem4[];
  float* s_w = reinterpret_cast<float*>(dwi_smem4);   // [K*K][gc*CV] filter slab of this block
  const int gc = p.gc, gcw = gc * CV;
  const int groups = p.c / CV;
  const int g0 = blockIdx.x * gc;
  const int lane_g = threadIdx.x % gc;
  const int sl = threadIdx.x / gc;
  const int nsl = kThreads / gc;
  const int g = g0 + lane_g;
  const bool live = sl < nsl && g < groups;
  const int img = blockIdx.z;

  // filter slab + bias: constants, loaded before the PDL wait (overlaps the predecessor's tail)
  for (int i = threadIdx.x; i < K * K * gc * (CV / 4); i += kThreads) {
    const int q4 = i % (CV / 4), lg = (i / (CV / 4)) % gc, tap = i / ((CV / 4) * gc);
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    if (g0 + lg < groups) a = __ldg(reinterpret_cast<const float4*>(p.wgt + (long long)tap * p.wp + (g0 + lg) * CV + q4 * 4));
    *reinterpret_cast<float4*>(s_w + tap * gcw + lg * CV + q4 * 4) = a;
  }
  unsigned long long bias2[NW];
#pragma unroll
  for (int e = 0; e < NW; ++e) bias2[e] = live ? pack_f2(__ldg(p.bias + g * CV + 2 * e), __ldg(p.bias + g * CV + 2 * e + 1)) : 0ull;
  __syncthreads();
  griddep_wait();   // PDL: the predecessor kernel has completed (ptx.cuh)
  griddep_launch();

  const int strips = (p.wo + TW - 1) / TW;
  const int trow0 = blockIdx.y * p.tile_rows;
  const int trow1 = min(trow0 + p.tile_rows, (p.ho + TH - 1) / TH);
  const int ntiles = (trow1 - trow0) * strips;
  float psum[CV];
#pragma unroll
  for (int e = 0; e < CV; ++e) psum[e] = 0.f;
  const __nv_bfloat16* ximg = p.x + (long long)img * p.h * p.w * p.xp + g * CV;
  __nv_bfloat16* yimg = p.y + (long long)img * p.ho * p.wo * p.yp + g * CV;
  const float* wl = s_w + lane_g * CV;

  if (live) {
    for (int t = sl; t < ntiles; t += nsl) {
      const int tr = trow0 + t / strips;
      const int ow0 = (t % strips) * TW;
      const int oh0 = tr * TH;
      unsigned long long acc[TH][TW][NW];
#pragma unroll
      for (int a = 0; a < TH; ++a)
#pragma unroll
        for (int o = 0; o < TW; ++o)
#pragma unroll
          for (int e = 0; e < NW; ++e) acc[a][o][e] = bias2[e];
      const int iw0 = ow0 * S - p.pad;
      // every column of the window in range? (interior strips: no per-load predicate at all)
      const bool cols_in = iw0 >= 0 && iw0 + NC <= p.w;
#pragma unroll
      for (int ir = 0; ir < NR; ++ir) {
        const int ih = oh0 * S - p.pad + ir;
        if (ih < 0 || ih >= p.h) continue;
        const __nv_bfloat16* row = ximg + (ih * p.w + iw0) * p.xp;   // 32-bit element offsets (host-checked, signed)
        Words<NW> raw[NC];
        if (cols_in) {
#pragma unroll
          for (int j = 0; j < NC; ++j) raw[j] = ldg_words<NW>(row + j * p.xp);
        } else {
#pragma unroll
          for (int j = 0; j < NC; ++j) {
#pragma unroll
            for (int e = 0; e < NW; ++e) raw[j].w[e] = 0u;
            if ((unsigned)(iw0 + j) < (unsigned)p.w) raw[j] = ldg_words<NW>(row + j * p.xp);
          }
        }
#pragma unroll
        for (int a = 0; a < TH; ++a) {
          const int r = ir - a * S;          // filter row this input row meets in output row a (compile time)
          if (r < 0 || r >= K) continue;
#pragma unroll
          for (int q = 0; q < K; ++q) {
            unsigned long long wq[NW];
#pragma unroll
            for (int e = 0; e < NW; e += 2) {
              const float4 w4 = *reinterpret_cast<const float4*>(wl + (r * K + q) * gcw + 2 * e);
              wq[e] = pack_f2(w4.x, w4.y), wq[e + 1] = pack_f2(w4.z, w4.w);
            }
#pragma unroll
            for (int o = 0; o < TW; ++o)
#pragma unroll
              for (int e = 0; e < NW; ++e) acc[a][o][e] = fma2(bf2_to_f2(raw[o * S + q].w[e]), wq[e], acc[a][o][e]);
          }
        }
      }
#pragma unroll
      for (int a = 0; a < TH; ++a) {
        const int oh = oh0 + a;
        if (oh >= p.ho) continue;
#pragma unroll
        for (int o = 0; o < TW; ++o) {
          if (ow0 + o >= p.wo) continue;
          uint32_t packed[NW];
#pragma unroll
          for (int e = 0; e < NW; ++e) {
            float lo, hi;
            unpack_f2(acc[a][o][e], lo, hi);
            const __nv_bfloat162 b2 = __floats2bfloat162_rn(act_ct<ACT>(lo), act_ct<ACT>(hi));
            packed[e] = *reinterpret_cast<const uint32_t*>(&b2);
            if (POOL) {   // the squeeze averages what the next layer will read: the bf16-rounded activation
              psum[2 * e] += __uint_as_float(packed[e] << 16);
              psum[2 * e + 1] += __uint_as_float(packed[e] & 0xffff0000u);
            }
          }
          __nv_bfloat16* dst = yimg + (oh * p.wo + ow0 + o) * p.yp;
          if (NW == 4)
            *reinterpret_cast<uint4*>(dst) = make_uint4(packed[0], packed[1], packed[2 % NW], packed[3 % NW]);
          else
            *reinterpret_cast<uint2*>(dst) = make_uint2(packed[0], packed[1]);
        }
      }
    }
  }
  if (!POOL) return;

  // ---- block reduction over the strip lanes (fixed order), then the image-level reduction by the last block ----
  __syncthreads();                       // the filter slab is dead: reuse shared memory
  float* red = reinterpret_cast<float*>(dwi_smem4);     // [nsl][gcw]
  if (sl < nsl) {
#pragma unroll
    for (int e = 0; e < CV; ++e) red[sl * gcw + lane_g * CV + e] = live ? psum[e] : 0.f;
  }
  __syncthreads();
  __shared__ int s_last;
  const int cvalid = min(gcw, p.c - g0 * CV);
  for (int ch = threadIdx.x; ch < cvalid; ch += kThreads) {
    float s = 0.f;
    for (int i = 0; i < nsl; ++i) s += red[i * gcw + ch];
    if (p.rchunks == 1) {
      p.pooled[(long long)img * p.pool_pitch + g0 * CV + ch] = __float2bfloat16_rn(s * p.inv_area);
    } else {
      p.partial[((long long)img * p.rchunks + blockIdx.y) * p.c + g0 * CV + ch] = s;
    }
  }
  if (p.rchunks == 1) return;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    int* ticket = p.tickets + (long long)img * gridDim.x + blockIdx.x;
    const int t = atomicAdd(ticket, 1);
    s_last = (t == p.rchunks - 1);
    if (s_last) *ticket = 0;             // self-cleaning: the workspace is ready for the next launch
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    for (int ch = threadIdx.x; ch < cvalid; ch += kThreads) {
      float s = 0.f;
      for (int r = 0; r < p.rchunks; ++r) s += __ldcg(p.partial + ((long long)img * p.rchunks + r) * p.c + g0 * CV + ch);
      p.pooled[(long long)img * p.pool_pitch + g0 * CV + ch] = __float2bfloat16_rn(s * p.inv_area);
    }
  }
}

using DwImgFn = void (*)(const DwImgParams);

template <int K, int S, int TW, int TH, int CV>
DwImgFn pick_act(int act, bool pool) {
  switch (act) {
    case EQXV_ACT_NONE:
      return pool ? dwconv_img_kernel<K, S, TW, TH, CV, EQXV_ACT_NONE, true> : dwconv_img_kernel<K, S, TW, TH, CV, EQXV_ACT_NONE, false>;
    case EQXV_ACT_RELU:
      return pool ? dwconv_img_kernel<K, S, TW, TH, CV, EQXV_ACT_RELU, true> : dwconv_img_kernel<K, S, TW, TH, CV, EQXV_ACT_RELU, false>;
    case EQXV_ACT_SILU:
      return pool ? dwconv_img_kernel<K, S, TW, TH, CV, EQXV_ACT_SILU, true> : dwconv_img_kernel<K, S, TW, TH, CV, EQXV_ACT_SILU, false>;
    case EQXV_ACT_HARDSWISH:
      return pool ? dwconv_img_kernel<K, S, TW, TH, CV, EQXV_ACT_HARDSWISH, true>
                  : dwconv_img_kernel<K, S, TW, TH, CV, EQXV_ACT_HARDSWISH, false>;
    default:
      return nullptr;
  }
}

struct Geometry {
  int gc, gchunks, th, cv, tile_rows, rchunks;
};

// channel groups per block: whole 128-byte lines per pixel when the layer is wide enough, and a divisor-friendly
// split otherwise (48 channels = 6 groups -> 6 per block)
int env_th() {
  static const int v = getenv("EQXV_DW_TH") ? atoi(getenv("EQXV_DW_TH")) : 0;
  return v;
}
int env_cv() {
  static const int v = getenv("EQXV_DW_CV") ? atoi(getenv("EQXV_DW_CV")) : 0;
  return v;
}

Geometry geometry(int n, int c, int ho, int k, int stride) {
  Geometry gm;
  // measured on the EfficientNet-B4 layers (tools/bench_dw.py, profiles/r02_bench_dw_*.txt): the kernel is issue bound;
  // 8 channels per thread x one output row (128 registers, 16 warps/SM) wins everywhere except the 5x5 stride-1 layers on
  // maps >= 28x28, where the 4-channel variant (80 registers, 24 warps/SM) is ~15 % faster
  gm.cv = env_cv() ? (env_cv() == 8 ? 8 : 4) : ((k == 5 && stride == 1 && ho >= 28) ? 4 : 8);
  const int groups = c / gm.cv;
  gm.gc = groups >= 32 ? 32 : (groups >= 16 ? 16 : groups);
  while (kThreads % gm.gc != 0 && gm.gc > 1) --gm.gc;
  gm.gchunks = (groups + gm.gc - 1) / gm.gc;
  gm.th = (stride == 1 && env_th() == 2) ? 2 : 1;
  const int trows = (ho + gm.th - 1) / gm.th;
  // enough blocks for ~4 per SM, but at least two row tiles per block
  const long long want = 4ll * device_sm_count();
  long long rch = (want + (long long)n * gm.gchunks - 1) / ((long long)n * gm.gchunks);
  rch = std::max(1ll, std::min<long long>(rch, std::max(1, trows / 2)));
  gm.tile_rows = (int)((trows + rch - 1) / rch);
  gm.rchunks = (trows + gm.tile_rows - 1) / gm.tile_rows;
  (void)k;
  return gm;
}

}  // namespace

bool dwconv_img_supported(int k, int stride, int dil, int act, int c, int x_pitch, int y_pitch, int w_pitch) {
  return (k == 3 || k == 5) && (stride == 1 || stride == 2) && dil == 1 && c % 8 == 0 && x_pitch % 8 == 0 &&
         y_pitch % 8 == 0 && w_pitch % 4 == 0 &&
         (act == EQXV_ACT_NONE || act == EQXV_ACT_RELU || act == EQXV_ACT_SILU || act == EQXV_ACT_HARDSWISH);
}

int dwconv_img_launch(const void* x, const float* wgt, const float* bias, void* y, void* pooled, void* workspace,
                      long long workspace_bytes, int n, int h, int w, int c, int k, int stride, int pad, int x_pitch,
                      int y_pitch, int w_pitch, int pool_pitch, int act, cudaStream_t stream) {
  const int ho = (h + 2 * pad - k) / stride + 1, wo = (w + 2 * pad - k) / stride + 1;
  EQXV_CHECK_ARG(ho > 0 && wo > 0 && n <= 65535, "dwconv: empty output / batch too large");
  const Geometry gm = geometry(n, c, ho, k, stride);
  DwImgParams p{};
  p.x = static_cast<const __nv_bfloat16*>(x), p.wgt = wgt, p.bias = bias, p.y = static_cast<__nv_bfloat16*>(y);
  p.pooled = static_cast<__nv_bfloat16*>(pooled);
  p.h = h, p.w = w, p.c = c, p.pad = pad, p.ho = ho, p.wo = wo, p.xp = x_pitch, p.yp = y_pitch, p.wp = w_pitch;
  p.pool_pitch = pool_pitch;
  p.gc = gm.gc, p.tile_rows = gm.tile_rows, p.rchunks = gm.rchunks;
  p.inv_area = 1.f / (float)((long long)ho * wo);
  const bool pool = pooled != nullptr;
  if (pool && gm.rchunks > 1) {
    const long long need_partial = (long long)n * gm.rchunks * c * 4;
    const long long need = ((need_partial + 255) / 256) * 256 + (long long)n * gm.gchunks * 4;
    EQXV_CHECK_ARG(workspace && workspace_bytes >= need, "dwconv+pool: workspace of %lld bytes needed, %lld given", need,
                   workspace_bytes);
    p.partial = static_cast<float*>(workspace);
    p.tickets = reinterpret_cast<int*>(static_cast<uint8_t*>(workspace) + ((need_partial + 255) / 256) * 256);
  }
  DwImgFn fn = nullptr;
  EQXV_CHECK_ARG((long long)h * w * x_pitch < (1ll << 31) && (long long)ho * wo * y_pitch < (1ll << 31),
                 "dwconv: one image must stay below 2^31 elements");
#define EQXV_DWI(K_, S_)                                                                                     \
  if (k == K_ && stride == S_) {                                                                             \
    if (gm.cv == 8)                                                                                          \
      fn = (gm.th == 2 && S_ == 1) ? pick_act<K_, S_, 4, (S_ == 1 ? 2 : 1), 8>(act, pool) : pick_act<K_, S_, 4, 1, 8>(act, pool); \
    else                                                                                                     \
      fn = (gm.th == 2 && S_ == 1) ? pick_act<K_, S_, 4, (S_ == 1 ? 2 : 1), 4>(act, pool) : pick_act<K_, S_, 4, 1, 4>(act, pool); \
  }
  EQXV_DWI(3, 1)
  EQXV_DWI(3, 2)
  EQXV_DWI(5, 1)
  EQXV_DWI(5, 2)
#undef EQXV_DWI
  if (!fn) {
    set_error("dwconv_img: k=%d stride=%d act=%d is not built", k, stride, act);
    return EQXV_ERR_UNSUPPORTED;
  }
  const int nsl = kThreads / gm.gc;
  const size_t smem = (size_t)std::max(k * k * gm.gc * gm.cv, pool ? nsl * gm.gc * gm.cv : 0) * 4;
  dim3 grid((unsigned)gm.gchunks, (unsigned)gm.rchunks, (unsigned)n);
  EQXV_CUDA(launch_kernel(fn, grid, dim3(kThreads), smem, stream, p));
  EQXV_CUDA(cudaGetLastError());
  return EQXV_OK;
}

long long dwconv_img_workspace_bytes(int n, int h, int c, int k, int stride, int pad) {
  const int ho = (h + 2 * pad - k) / stride + 1;
  const Geometry gm = geometry(n, c, std::max(ho, 1), k, stride);
  if (gm.rchunks <= 1) return 0;
  const long long need_partial = (long long)n * gm.rchunks * c * 4;
  return ((need_partial + 255) / 256) * 256 + (long long)n * gm.gchunks * 4;
}

}  // namespace eqxv

using namespace eqxv;

extern "C" int eqxv_dwconv_pool_workspace_bytes(int32_t n, int32_t h, int32_t w, int32_t c, int32_t k, int32_t stride,
                                                int32_t pad, int64_t* bytes) {
  EQXV_CHECK_ARG(bytes && n > 0 && h > 0 && w > 0 && c > 0 && k > 0 && stride > 0, "dwconv_pool_workspace_bytes: bad arguments");
  *bytes = dwconv_img_workspace_bytes(n, h, c, k, stride, pad);
  return EQXV_OK;
}

extern "C" int eqxv_dwconv_bn_act_pool_bf16(const void* x, const float* wgt, const float* bias, void* y, void* pooled,
                                            void* workspace, int64_t workspace_bytes, int32_t n, int32_t h, int32_t w,
                                            int32_t c, int32_t k, int32_t stride, int32_t pad, int32_t x_pitch,
                                            int32_t y_pitch, int32_t w_pitch, int32_t pool_pitch, int32_t act,
                                            void* stream) {
  EQXV_CHECK_ARG(x && wgt && bias && y && pooled && n > 0 && h > 0 && w > 0 && c > 0, "dwconv+pool: bad arguments");
  EQXV_CHECK_ARG(x_pitch >= c && y_pitch >= c && w_pitch >= c && pool_pitch >= c, "dwconv+pool: pitches smaller than c");
  EQXV_CHECK_ARG((((uintptr_t)x | (uintptr_t)y) & 15) == 0 && (((uintptr_t)wgt | (uintptr_t)bias) & 15) == 0,
                 "dwconv+pool: operands must be 16-byte aligned");
  if (!dwconv_img_supported(k, stride, 1, act, c, x_pitch, y_pitch, w_pitch)) {
    set_error("dwconv+pool: k=%d stride=%d act=%d c=%d is outside the fused kernel (k in {3,5}, stride 1/2, "
              "none/relu/silu/hard-swish, channels and pitches multiples of 8)", k, stride, act, c);
    return EQXV_ERR_UNSUPPORTED;
  }
  return dwconv_img_launch(x, wgt, bias, y, pooled, workspace, workspace_bytes, n, h, w, c, k, stride, pad, x_pitch,
                           y_pitch, w_pitch, pool_pitch, act, (cudaStream_t)stream);
}
