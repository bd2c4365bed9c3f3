// Depthwise convolution and the elementwise passes that cannot ride in a GEMM epilogue.
// HBM-bound, integer-free data paths: channels-last, 16-byte vectors (8 bf16 channels) per thread,
// no tensor cores (K3/K8/K11 of SURVEY.md §2.1).
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "common.h"
#include "ptx.cuh"

namespace eqxv {

// dwconv_img.cu
bool dwconv_img_supported(int k, int stride, int dil, int act, int c, int x_pitch, int y_pitch, int w_pitch);
int dwconv_img_launch(const void* x, const float* wgt, const float* bias, void* y, void* pooled, void* workspace,
                      long long workspace_bytes, int n, int h, int w, int c, int k, int stride, int pad, int x_pitch,
                      int y_pitch, int w_pitch, int pool_pitch, int act, cudaStream_t stream);

constexpr int kDwThreads = 256;

// Grid-stride loop over `total` work items with a 32-bit index whenever it fits: the index decomposition
// (i % groups, / wo, % ho ...) costs 3-4 divisions per item, and 64-bit integer division is a ~100-instruction
// software routine -- as much issue time as the arithmetic of a depthwise strip.
#define EQXV_GRID_STRIDE(total, body)                                                                        \
  do {                                                                                                       \
    if ((total) <= 0x7fffffffLL) {                                                                           \
      for (unsigned i_ = blockIdx.x * blockDim.x + threadIdx.x; i_ < (unsigned)(total);                      \
           i_ += gridDim.x * blockDim.x)                                                                     \
        body(i_);                                                                                            \
    } else {                                                                                                 \
      for (long long i_ = blockIdx.x * (long long)blockDim.x + threadIdx.x; i_ < (total);                    \
           i_ += (long long)gridDim.x * blockDim.x)                                                          \
        body(i_);                                                                                            \
    }                                                                                                        \
  } while (0)

struct alignas(16) bf16x8 {
  __nv_bfloat162 v[4];
};
__device__ __forceinline__ void unpack8(const uint4& a, float* f) {
  const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&a);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(p[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  __nv_bfloat162 r[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) r[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return *reinterpret_cast<const uint4*>(r);
}

__device__ __forceinline__ float act_rt(float v, int act) {
  switch (act) {
    case EQXV_ACT_RELU: return fmaxf(v, 0.f);
    case EQXV_ACT_SILU: {   // h + h * tanh(h), h = x / 2: one special-function op (see igemm.cu apply_act)
      const float h = 0.5f * v;
      float th;
      asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(h));
      return fmaf(h, th, h);
    }
    case EQXV_ACT_GELU_TANH: {
      const float u = 0.7978845608028654f * (v + 0.044715f * v * v * v);
      float th;
      asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(u));
      return 0.5f * v * (1.f + th);
    }
    case EQXV_ACT_HARDSWISH: return v * fminf(fmaxf(v + 3.f, 0.f), 6.f) * (1.f / 6.f);
    case EQXV_ACT_SIGMOID: {
      float th;
      asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(0.5f * v));
      return fmaf(0.5f, th, 0.5f);
    }
    case EQXV_ACT_HARDSIGMOID: return fminf(fmaxf(v + 3.f, 0.f), 6.f) * (1.f / 6.f);
    case EQXV_ACT_RELU6: return fminf(fmaxf(v, 0.f), 6.f);
    default: return v;
  }
}

static inline int grid_for(long long work, int threads = kDwThreads) {
  long long b = (work + threads - 1) / threads;
  const long long cap = (long long)device_sm_count() * 32;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

// ---------------------------------------------------------------------------------------------
// depthwise KxK (groups == channels) + folded BN + activation.
// One thread = 8 channels of one output pixel; consecutive threads = consecutive channel groups, so
// a warp reads whole 128-byte lines of every tap and neighbouring pixels of a block share taps in
// L1. Filter: fp32 [K*K][c_pad] (BN scale folded), bias fp32 [c_pad].
// ---------------------------------------------------------------------------------------------
template <int K, int S>
__global__ void dwconv_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ wgt,
                              const float* __restrict__ bias, __nv_bfloat16* __restrict__ y, int n, int h,
                              int w, int c, int pad, int dil, int ho, int wo, int xp, int yp, int wp,
                              int act) {
  griddep_wait();   // PDL: the predecessor kernel has completed (ptx.cuh)
  griddep_launch();
  const int groups = c / 8;
  const long long total = (long long)n * ho * wo * groups;
  auto body = [&](auto i) {
    const int g = (int)(i % groups);
    auto t = i / groups;
    const int ow = (int)(t % wo);
    t /= wo;
    const int oh = (int)(t % ho);
    const int img = (int)(t / ho);
    float acc[8];
    {
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + g * 8));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + g * 8 + 4));
      acc[0] = b0.x, acc[1] = b0.y, acc[2] = b0.z, acc[3] = b0.w;
      acc[4] = b1.x, acc[5] = b1.y, acc[6] = b1.z, acc[7] = b1.w;
    }
#pragma unroll
    for (int r = 0; r < K; ++r) {
      const int ih = oh * S - pad + r * dil;
      const bool hok = ih >= 0 && ih < h;
      const int ihc = min(max(ih, 0), h - 1);
      uint4 raw[K];
      bool ok[K];
#pragma unroll
      for (int q = 0; q < K; ++q) {
        const int iw = ow * S - pad + q * dil;
        ok[q] = hok && iw >= 0 && iw < w;
        const int iwc = min(max(iw, 0), w - 1);
        raw[q] = __ldg(reinterpret_cast<const uint4*>(x + (((long long)img * h + ihc) * w + iwc) * xp + g * 8));
      }
#pragma unroll
      for (int q = 0; q < K; ++q) {
        float f[8];
        unpack8(raw[q], f);
        const float* wr = wgt + (long long)(r * K + q) * wp + g * 8;
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(wr));
        const float4 w1 = __ldg(reinterpret_cast<const float4*>(wr + 4));
        if (ok[q]) {  // zero padding: the tap contributes nothing
          acc[0] = fmaf(f[0], w0.x, acc[0]);
          acc[1] = fmaf(f[1], w0.y, acc[1]);
          acc[2] = fmaf(f[2], w0.z, acc[2]);
          acc[3] = fmaf(f[3], w0.w, acc[3]);
          acc[4] = fmaf(f[4], w1.x, acc[4]);
          acc[5] = fmaf(f[5], w1.y, acc[5]);
          acc[6] = fmaf(f[6], w1.z, acc[6]);
          acc[7] = fmaf(f[7], w1.w, acc[7]);
        }
      }
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[q] = act_rt(acc[q], act);
    *reinterpret_cast<uint4*>(y + (((long long)img * ho + oh) * wo + ow) * yp + g * 8) = pack8(acc);
  };
  EQXV_GRID_STRIDE(total, body);
}

// Strip variant (dilation 1): one thread produces TW horizontally adjacent outputs of 8 channels, so
// every loaded input column feeds up to K/S outputs and every filter row is loaded once per strip:
// K*((TW-1)*S+K) 16-byte loads for TW outputs instead of TW*K*K, and all loads of a filter row are in
// flight together.
template <int K, int S, int TW>
__global__ void __launch_bounds__(kDwThreads) dwconv_strip_kernel(
    const __nv_bfloat16* __restrict__ x, const float* __restrict__ wgt, const float* __restrict__ bias,
    __nv_bfloat16* __restrict__ y, int n, int h, int w, int c, int pad, int ho, int wo, int xp, int yp, int wp,
    int act) {
  griddep_wait();   // PDL: the predecessor kernel has completed (ptx.cuh)
  griddep_launch();
  constexpr int NC = (TW - 1) * S + K;
  const int groups = c / 8;
  const int strips = (wo + TW - 1) / TW;
  const long long total = (long long)n * ho * strips * groups;
  auto body = [&](auto i) {
    const int g = (int)(i % groups);
    auto t = i / groups;
    const int ow0 = (int)(t % strips) * TW;
    t /= strips;
    const int oh = (int)(t % ho);
    const int img = (int)(t / ho);
    float acc[TW][8];
    {
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + g * 8));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + g * 8 + 4));
#pragma unroll
      for (int o = 0; o < TW; ++o) {
        acc[o][0] = b0.x, acc[o][1] = b0.y, acc[o][2] = b0.z, acc[o][3] = b0.w;
        acc[o][4] = b1.x, acc[o][5] = b1.y, acc[o][6] = b1.z, acc[o][7] = b1.w;
      }
    }
#pragma unroll
    for (int r = 0; r < K; ++r) {
      const int ih = oh * S - pad + r;
      if (ih < 0 || ih >= h) continue;  // warp-uniform for most warps (same output row)
      const __nv_bfloat16* row = x + ((long long)img * h + ih) * w * xp + g * 8;
      uint4 raw[NC];
#pragma unroll
      for (int j = 0; j < NC; ++j) {
        const int iw = ow0 * S - pad + j;
        const int iwc = min(max(iw, 0), w - 1);
        raw[j] = __ldg(reinterpret_cast<const uint4*>(row + (long long)iwc * xp));
      }
      float wr[K][8];
#pragma unroll
      for (int q = 0; q < K; ++q) {
        const float* wq = wgt + (long long)(r * K + q) * wp + g * 8;
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(wq));
        const float4 w1 = __ldg(reinterpret_cast<const float4*>(wq + 4));
        wr[q][0] = w0.x, wr[q][1] = w0.y, wr[q][2] = w0.z, wr[q][3] = w0.w;
        wr[q][4] = w1.x, wr[q][5] = w1.y, wr[q][6] = w1.z, wr[q][7] = w1.w;
      }
#pragma unroll
      for (int j = 0; j < NC; ++j) {
        const int iw = ow0 * S - pad + j;
        float f[8];
        unpack8(raw[j], f);
        if (iw < 0 || iw >= w) {
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = 0.f;
        }
#pragma unroll
        for (int o = 0; o < TW; ++o) {
          const int q = j - o * S;  // compile-time after unrolling
          if (q >= 0 && q < K) {
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[o][e] = fmaf(f[e], wr[q][e], acc[o][e]);
          }
        }
      }
    }
#pragma unroll
    for (int o = 0; o < TW; ++o) {
      if (ow0 + o < wo) {
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[o][e] = act_rt(acc[o][e], act);
        *reinterpret_cast<uint4*>(y + (((long long)img * ho + oh) * wo + ow0 + o) * yp + g * 8) = pack8(acc[o]);
      }
    }
  };
  EQXV_GRID_STRIDE(total, body);
}

// ---------------------------------------------------------------------------------------------
// y = act(x * scale[c] + shift[c] + other) * gate[img, c]      (every operand but x optional)
//   standalone BatchNorm(+ReLU) of DenseNet (densenet.py:64-65), residual adds that could not be
//   folded, the SE channel gate x * sigmoid(...) (squeeze.py:61).
// ---------------------------------------------------------------------------------------------
__global__ void eltwise_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ scale,
                               const float* __restrict__ shift, const __nv_bfloat16* __restrict__ other,
                               const __nv_bfloat16* __restrict__ gate, __nv_bfloat16* __restrict__ y,
                               long long rows, int c, int xp, int op, int gp, int yp, int rows_per_image,
                               int act) {
  griddep_wait();   // PDL: the predecessor kernel has completed (ptx.cuh)
  griddep_launch();
  const int groups = c / 8;
  const long long total = rows * groups;
  auto body = [&](auto i) {
    const int g = (int)(i % groups);
    const auto row = i / groups;
    float f[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(x + (long long)row * xp + g * 8)), f);
    if (scale != nullptr) {
      const float4 s0 = __ldg(reinterpret_cast<const float4*>(scale + g * 8));
      const float4 s1 = __ldg(reinterpret_cast<const float4*>(scale + g * 8 + 4));
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(shift + g * 8));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(shift + g * 8 + 4));
      f[0] = fmaf(f[0], s0.x, b0.x), f[1] = fmaf(f[1], s0.y, b0.y);
      f[2] = fmaf(f[2], s0.z, b0.z), f[3] = fmaf(f[3], s0.w, b0.w);
      f[4] = fmaf(f[4], s1.x, b1.x), f[5] = fmaf(f[5], s1.y, b1.y);
      f[6] = fmaf(f[6], s1.z, b1.z), f[7] = fmaf(f[7], s1.w, b1.w);
    }
    if (other != nullptr) {
      float o[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(other + (long long)row * op + g * 8)), o);
#pragma unroll
      for (int q = 0; q < 8; ++q) f[q] += o[q];
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) f[q] = act_rt(f[q], act);
    if (gate != nullptr) {
      float s[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(gate + (long long)(row / rows_per_image) * gp + g * 8)), s);
#pragma unroll
      for (int q = 0; q < 8; ++q) f[q] *= s[q];
    }
    *reinterpret_cast<uint4*>(y + (long long)row * yp + g * 8) = pack8(f);
  };
  EQXV_GRID_STRIDE(total, body);
}

// ---------------------------------------------------------------------------------------------
// Shared-memory stencil variant (the production path for K in {3,5}, stride 1/2, dilation 1).
//
// The register-strip kernels above re-read every input row K/S times through L2 (vertical neighbours
// live in different blocks): EfficientNet-B4's depthwise layers sat at 0.7-1.5 TB/s of DRAM traffic,
// L2-bandwidth bound. Here a persistent CTA stages the input halo tile of a
// (64-channel block) x (THo x TWo outputs) tile ONCE with a single TMA box -- out-of-bounds zero fill
// is the convolution padding -- together with the block's K*K x 64 fp32 filter slab, double-buffered so
// that the next tile is in flight while the current one is computed.
// Thread = 8 channels (one 16-byte vector) of one output column and RH consecutive output rows: each
// staged input row is read once per thread (K vectors) and feeds up to K output rows held in
// registers. A warp covers 4 adjacent pixels x 64 channels = 512 contiguous bytes per shared-memory
// read (conflict free) and per global store (full 128-byte lines).
// Reference: equinox.nn.Conv2d(groups=C) + BatchNorm + activation of _MBConv (efficientnet.py:138-150)
// and _InvertedResidual (mobilenetv3.py:88-101).
// ---------------------------------------------------------------------------------------------
struct alignas(64) DwTileParams {
  CUtensorMap tmX, tmW;
  const float* bias;
  __nv_bfloat16* y;
  int h, w, c, ho, wo, yp, pad, act;
  int two, rg, rh;                   // output columns per tile, row groups (two * rg == 32), output rows per thread
  int tiles_x, tiles_y, cblocks, num_tiles;
  int iw, ih;                        // staged input tile (pixels)
  int in_bytes, w_bytes, buf_bytes;  // per buffer: input tile, filter slab, total (128-byte multiples)
};

template <int K, int S>
__global__ void __launch_bounds__(256, 1) dwconv_tile_kernel(const __grid_constant__ DwTileParams p) {
  constexpr int W = (K - 1) / S + 1;          // output rows in flight per thread
  const int RH = p.rh;                        // output rows per thread (<= 8 for stride 1, <= 4 for stride 2)
  const int IHT = (RH - 1) * S + K;           // input rows one thread walks over
  extern __shared__ uint8_t dw_smem_raw[];
  const uint32_t raw = smem_u32(dw_smem_raw);
  const uint32_t base = (raw + 127u) & ~127u;
  uint8_t* gbase = dw_smem_raw + (base - raw);
  const uint32_t bars = base + 2u * (uint32_t)p.buf_bytes;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.tmX);
    tma_prefetch_desc(&p.tmW);
    mbar_init(bars, 1);
    mbar_init(bars + 8, 1);
    mbar_fence_init();
  }
  const int cg = threadIdx.x & 7;
  const int slot = threadIdx.x >> 3;          // 0..31
  const int col = slot % p.two, rgi = slot / p.two;
  const int tho = p.rg * RH;

  auto issue = [&](int tile, uint32_t b) {    // one thread
    int t = tile;
    const int cb = t % p.cblocks;
    t /= p.cblocks;
    const int tx = t % p.tiles_x;
    t /= p.tiles_x;
    const int ty = t % p.tiles_y;
    const int img = t / p.tiles_y;
    const uint32_t dst = base + b * (uint32_t)p.buf_bytes;
    mbar_expect_tx(bars + 8u * b, (uint32_t)(p.in_bytes + p.w_bytes));
    tma_load_4d(dst, &p.tmX, bars + 8u * b, cb * 64, tx * p.two * S - p.pad, ty * tho * S - p.pad, img);
    tma_load_2d(dst + (uint32_t)p.in_bytes, &p.tmW, bars + 8u * b, cb * 64, 0);
  };

  griddep_wait();   // PDL: the producer of x has completed
  __syncthreads();
  griddep_launch();
  if (threadIdx.x == 0 && (int)blockIdx.x < p.num_tiles) issue((int)blockIdx.x, 0u);

  int it = 0;
  for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
    const uint32_t b = (uint32_t)it & 1u;
    if (threadIdx.x == 0 && tile + (int)gridDim.x < p.num_tiles) issue(tile + (int)gridDim.x, b ^ 1u);
    int t = tile;
    const int cb = t % p.cblocks;
    t /= p.cblocks;
    const int tx = t % p.tiles_x;
    t /= p.tiles_x;
    const int ty = t % p.tiles_y;
    const int img = t / p.tiles_y;
    const int ch = cb * 64 + cg * 8;
    float bia[8];
    {
      float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
      if (ch < p.c) {
        b0 = __ldg(reinterpret_cast<const float4*>(p.bias + ch));
        b1 = __ldg(reinterpret_cast<const float4*>(p.bias + ch + 4));
      }
      bia[0] = b0.x, bia[1] = b0.y, bia[2] = b0.z, bia[3] = b0.w;
      bia[4] = b1.x, bia[5] = b1.y, bia[6] = b1.z, bia[7] = b1.w;
    }
    // Sliding window over the input rows: win[d] accumulates output row (j - d) of this thread's column;
    // step j consumes input rows j*S .. j*S+S-1, after which output row j-(W-1) is complete and leaves
    // through global memory. Registers: W*8 accumulators whatever RH is.
    float win[W][8];
#pragma unroll
    for (int d = 0; d < W; ++d)
#pragma unroll
      for (int e = 0; e < 8; ++e) win[d][e] = bia[e];
    mbar_wait(bars + 8u * b, (uint32_t)(it >> 1) & 1u);
    const uint8_t* in = gbase + b * p.buf_bytes + ((size_t)(rgi * RH * S) * p.iw + col * S) * 128 + cg * 16;
    const float* wsm = reinterpret_cast<const float*>(gbase + b * p.buf_bytes + p.in_bytes) + cg * 8;
    const int ow = tx * p.two + col;
    const int oh0 = ty * tho + rgi * RH;
    const bool live = ow < p.wo && ch < p.c;
    __nv_bfloat16* yout = p.y + (((long long)img * p.ho + oh0) * p.wo + ow) * p.yp + ch;
    // The block's filter moves from shared memory to REGISTERS once per tile: read per use, the 8 lanes of a
    // quarter-warp fetch 8 different 32-byte rows (2 wavefronts x 4 quarters per load) and the filter reads
    // were 90 % of the shared-memory traffic (ncu r01s8: 60 % of all wavefronts were bank conflicts, issue
    // slots 24 % busy). K = 3: fp32 (72 registers); K = 5: bf16 pairs (100 registers, one rounding of the
    // BN-folded filter, the same rounding the dense convolutions apply to theirs).
    constexpr int WREG = K == 3 ? 8 : 4;
    uint32_t wr[K * K][WREG];
#pragma unroll
    for (int tp = 0; tp < K * K; ++tp) {
      const float4 w0 = *reinterpret_cast<const float4*>(wsm + tp * 64);
      const float4 w1 = *reinterpret_cast<const float4*>(wsm + tp * 64 + 4);
      if constexpr (K == 3) {
        wr[tp][0] = __float_as_uint(w0.x), wr[tp][1] = __float_as_uint(w0.y), wr[tp][2] = __float_as_uint(w0.z);
        wr[tp][3] = __float_as_uint(w0.w), wr[tp][4] = __float_as_uint(w1.x), wr[tp][5] = __float_as_uint(w1.y);
        wr[tp][6] = __float_as_uint(w1.z), wr[tp][7] = __float_as_uint(w1.w);
      } else {
        const __nv_bfloat162 a = __floats2bfloat162_rn(w0.x, w0.y), bq = __floats2bfloat162_rn(w0.z, w0.w);
        const __nv_bfloat162 cq = __floats2bfloat162_rn(w1.x, w1.y), dq = __floats2bfloat162_rn(w1.z, w1.w);
        wr[tp][0] = *reinterpret_cast<const uint32_t*>(&a), wr[tp][1] = *reinterpret_cast<const uint32_t*>(&bq);
        wr[tp][2] = *reinterpret_cast<const uint32_t*>(&cq), wr[tp][3] = *reinterpret_cast<const uint32_t*>(&dq);
      }
    }
#pragma unroll 1
    for (int j = 0; j < RH + W - 1; ++j) {
#pragma unroll
      for (int sr = 0; sr < S; ++sr) {
        const int i = j * S + sr;
        if (i < IHT) {
          const uint8_t* row = in + (size_t)i * p.iw * 128;
#pragma unroll
          for (int q = 0; q < K; ++q) {
            float f[8];
            unpack8(*reinterpret_cast<const uint4*>(row + q * 128), f);
#pragma unroll
            for (int r = sr; r < K; r += S) {
              const int d = (r - sr) / S;
              float w8[8];
              if constexpr (K == 3) {
#pragma unroll
                for (int e = 0; e < 8; ++e) w8[e] = __uint_as_float(wr[r * K + q][e]);
              } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  w8[2 * e] = __uint_as_float(wr[r * K + q][e] << 16);
                  w8[2 * e + 1] = __uint_as_float(wr[r * K + q][e] & 0xffff0000u);
                }
              }
#pragma unroll
              for (int e = 0; e < 8; ++e) win[d][e] = fmaf(f[e], w8[e], win[d][e]);
            }
          }
        }
      }
      const int o = j - (W - 1);
      if (o >= 0 && live && oh0 + o < p.ho) {
        float r8[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) r8[e] = act_rt(win[W - 1][e], p.act);
        *reinterpret_cast<uint4*>(yout + (long long)o * p.wo * p.yp) = pack8(r8);
      }
#pragma unroll
      for (int d = W - 1; d > 0; --d)
#pragma unroll
        for (int e = 0; e < 8; ++e) win[d][e] = win[d - 1][e];
#pragma unroll
      for (int e = 0; e < 8; ++e) win[0][e] = bia[e];
    }
    __syncthreads();   // every thread is done with buffer b before the TMA of iteration it+1 refills it
  }
}

template <int K, int S>
static int launch_dw_tile(const void* x, const float* wgt, const float* bias, void* y, int n, int h, int w, int c,
                          int pad, int ho, int wo, int xp, int yp, int wp, int act, cudaStream_t st) {
  constexpr int RH_MAX = S == 1 ? 8 : 4;
  DwTileParams p;
  memset(&p, 0, sizeof(p));
  p.bias = bias;
  p.y = reinterpret_cast<__nv_bfloat16*>(y);
  p.h = h, p.w = w, p.c = c, p.ho = ho, p.wo = wo, p.yp = yp, p.pad = pad, p.act = act;
  p.two = wo >= 24 ? 32 : (wo >= 12 ? 16 : 8);
  p.rg = 32 / p.two;
  // rows per thread: as many as fit RH_MAX while wasting as few tile rows as possible (7x7 maps: 4 row groups x 2)
  {
    const int tiles_y = ceil_div(ho, p.rg * RH_MAX);
    p.rh = std::max(1, ceil_div(ceil_div(ho, tiles_y), p.rg));
  }
  const int tho = p.rg * p.rh;
  p.tiles_x = ceil_div(wo, p.two);
  p.tiles_y = ceil_div(ho, tho);
  p.cblocks = ceil_div(c, 64);
  const long long nt = (long long)n * p.tiles_x * p.tiles_y * p.cblocks;
  EQXV_CHECK_ARG(nt > 0 && nt < (1ll << 30), "dwconv: bad tile count");
  p.num_tiles = (int)nt;
  p.iw = (p.two - 1) * S + K;
  p.ih = (tho - 1) * S + K;
  p.in_bytes = p.iw * p.ih * 128;
  p.w_bytes = K * K * 64 * 4;
  p.buf_bytes = ceil_div(p.in_bytes + p.w_bytes, 128) * 128;
  const int smem = 2 * p.buf_bytes + 16 + 128;
  EQXV_CHECK_ARG(smem <= 232448 && p.iw <= 256 && p.ih <= 256, "dwconv: tile does not fit in shared memory");
  TmapSpec a{};
  a.base = const_cast<void*>(x);
  a.dtype = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  a.rank = 4;
  a.swizzle = CU_TENSOR_MAP_SWIZZLE_NONE;
  a.dims[0] = (uint64_t)c, a.dims[1] = (uint64_t)w, a.dims[2] = (uint64_t)h, a.dims[3] = (uint64_t)n;
  a.strides_bytes[0] = (uint64_t)xp * 2;
  a.strides_bytes[1] = a.strides_bytes[0] * (uint64_t)w;
  a.strides_bytes[2] = a.strides_bytes[1] * (uint64_t)h;
  a.box[0] = 64, a.box[1] = (uint32_t)p.iw, a.box[2] = (uint32_t)p.ih, a.box[3] = 1;
  a.estride[0] = a.estride[1] = a.estride[2] = a.estride[3] = 1;
  int rc = encode_tmap(&p.tmX, a);
  if (rc) return rc;
  TmapSpec wm{};
  wm.base = const_cast<float*>(wgt);
  wm.dtype = CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  wm.rank = 2;
  wm.swizzle = CU_TENSOR_MAP_SWIZZLE_NONE;
  wm.dims[0] = (uint64_t)wp, wm.dims[1] = (uint64_t)(K * K);
  wm.strides_bytes[0] = (uint64_t)wp * 4;
  wm.box[0] = 64, wm.box[1] = (uint32_t)(K * K);
  wm.estride[0] = wm.estride[1] = 1;
  rc = encode_tmap(&p.tmW, wm);
  if (rc) return rc;
  static bool attr_done = false;
  if (!attr_done) {
    EQXV_CUDA(cudaFuncSetAttribute(dwconv_tile_kernel<K, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    attr_done = true;
  }
  const int grid = std::min(p.num_tiles, device_sm_count());   // one persistent CTA per SM (filter in registers)
  EQXV_CUDA(launch_kernel(dwconv_tile_kernel<K, S>, dim3(grid), dim3(256), (size_t)smem, st, p));
  return EQXV_OK;
}

}  // namespace eqxv

using namespace eqxv;

static int dwconv_impl(const void* x, const float* wgt, const float* bias, void* y, int32_t n, int32_t h, int32_t w,
                       int32_t c, int32_t k, int32_t stride, int32_t pad, int32_t dil, int32_t x_pitch,
                       int32_t y_pitch, int32_t w_pitch, int32_t act, void* stream, bool force_tile) {
  EQXV_CHECK_ARG(x && wgt && bias && y && n > 0 && h > 0 && w > 0 && c > 0, "dwconv: bad arguments");
  EQXV_CHECK_ARG(c % 8 == 0 && x_pitch % 8 == 0 && y_pitch % 8 == 0 && w_pitch % 4 == 0 && x_pitch >= c &&
                     y_pitch >= c && w_pitch >= c,
                 "dwconv: channels/pitches must be multiples of 8");
  EQXV_CHECK_ARG(stride >= 1 && stride <= 2 && dil >= 1 && pad >= 0, "dwconv: bad geometry");
  const int ho = (h + 2 * pad - dil * (k - 1) - 1) / stride + 1;
  const int wo = (w + 2 * pad - dil * (k - 1) - 1) / stride + 1;
  EQXV_CHECK_ARG(ho > 0 && wo > 0, "dwconv: empty output");
  const long long total = (long long)n * ho * wo * (c / 8);
  const int grid = grid_for(total);
  cudaStream_t st = (cudaStream_t)stream;
  const __nv_bfloat16* xi = (const __nv_bfloat16*)x;
  __nv_bfloat16* yo = (__nv_bfloat16*)y;
  // Shared-memory stencil path (TMA-staged halo tiles): opt-in (EQXV_DWTILE=1 or eqxv_dwconv_tile_bf16)
  // until it beats the register strips on every EfficientNet/MobileNet shape (tools/bench_dw.py).
  static const bool use_tile = getenv("EQXV_DWTILE") != nullptr;
  if ((use_tile || force_tile) && dil == 1 && (k == 3 || k == 5) && ((uintptr_t)x & 15) == 0 && ((uintptr_t)wgt & 15) == 0 &&
      w_pitch % 4 == 0 && 2 * pad <= k) {
    if (k == 3 && stride == 1) return launch_dw_tile<3, 1>(x, wgt, bias, y, n, h, w, c, pad, ho, wo, x_pitch, y_pitch, w_pitch, act, st);
    if (k == 3 && stride == 2) return launch_dw_tile<3, 2>(x, wgt, bias, y, n, h, w, c, pad, ho, wo, x_pitch, y_pitch, w_pitch, act, st);
    if (k == 5 && stride == 1) return launch_dw_tile<5, 1>(x, wgt, bias, y, n, h, w, c, pad, ho, wo, x_pitch, y_pitch, w_pitch, act, st);
    if (k == 5 && stride == 2) return launch_dw_tile<5, 2>(x, wgt, bias, y, n, h, w, c, pad, ho, wo, x_pitch, y_pitch, w_pitch, act, st);
  }
  if (force_tile) {
    set_error("dwconv: the shared-memory stencil kernel needs k in {3,5}, dilation 1, 16-byte aligned operands");
    return EQXV_ERR_UNSUPPORTED;
  }
  {
    // production path: packed-FMA per-image kernel (dwconv_img.cu); EQXV_DWIMG=0 falls back to the strip kernels (A/B)
    static const bool use_img = !(getenv("EQXV_DWIMG") && getenv("EQXV_DWIMG")[0] == '0');
    const bool aligned = ((((uintptr_t)x | (uintptr_t)y | (uintptr_t)wgt | (uintptr_t)bias) & 15) == 0) && n <= 65535;
    if (use_img && aligned && dwconv_img_supported(k, stride, dil, act, c, x_pitch, y_pitch, w_pitch))
      return dwconv_img_launch(x, wgt, bias, y, nullptr, nullptr, 0, n, h, w, c, k, stride, pad, x_pitch, y_pitch,
                               w_pitch, 0, act, st);
  }
#define EQXV_DWS(K, S, TW)                                                                          \
  EQXV_CUDA(launch_kernel(dwconv_strip_kernel<K, S, TW>,                                            \
                          dim3(grid_for((long long)n * ho * ((wo + TW - 1) / TW) * (c / 8))),       \
                          dim3(kDwThreads), (size_t)0, st, xi, wgt, bias, yo, n, h, w, c, pad, ho,   \
                          wo, x_pitch, y_pitch, w_pitch, act))
  if (dil == 1 && wo >= 4) {
    bool done = true;
    if (k == 3 && stride == 1) {
      EQXV_DWS(3, 1, 4);
    } else if (k == 3 && stride == 2) {
      EQXV_DWS(3, 2, 4);
    } else if (k == 5 && stride == 1) {
      EQXV_DWS(5, 1, 4);
    } else if (k == 5 && stride == 2) {
      EQXV_DWS(5, 2, 4);
    } else {
      done = false;
    }
    if (done) {
      EQXV_CUDA(cudaGetLastError());
      return EQXV_OK;
    }
  }
#undef EQXV_DWS
#define EQXV_DW(K, S)                                                                               \
  EQXV_CUDA(launch_kernel(dwconv_kernel<K, S>, dim3(grid), dim3(kDwThreads), (size_t)0, st, xi, wgt, \
                          bias, yo, n, h, w, c, pad, dil, ho, wo, x_pitch, y_pitch, w_pitch, act))
  if (k == 3 && stride == 1) {
    EQXV_DW(3, 1);
  } else if (k == 3 && stride == 2) {
    EQXV_DW(3, 2);
  } else if (k == 5 && stride == 1) {
    EQXV_DW(5, 1);
  } else if (k == 5 && stride == 2) {
    EQXV_DW(5, 2);
  } else if (k == 7 && stride == 1) {
    EQXV_DW(7, 1);
  } else {
    set_error("dwconv: kernel %dx%d stride %d is not built", k, k, stride);
    return EQXV_ERR_UNSUPPORTED;
  }
#undef EQXV_DW
  EQXV_CUDA(cudaGetLastError());
  return EQXV_OK;
}

extern "C" int eqxv_dwconv_bn_act_bf16(const void* x, const float* wgt, const float* bias, void* y,
                                       int32_t n, int32_t h, int32_t w, int32_t c, int32_t k,
                                       int32_t stride, int32_t pad, int32_t dil, int32_t x_pitch,
                                       int32_t y_pitch, int32_t w_pitch, int32_t act, void* stream) {
  return dwconv_impl(x, wgt, bias, y, n, h, w, c, k, stride, pad, dil, x_pitch, y_pitch, w_pitch, act, stream, false);
}

extern "C" int eqxv_dwconv_tile_bf16(const void* x, const float* wgt, const float* bias, void* y, int32_t n,
                                     int32_t h, int32_t w, int32_t c, int32_t k, int32_t stride, int32_t pad,
                                     int32_t dil, int32_t x_pitch, int32_t y_pitch, int32_t w_pitch, int32_t act,
                                     void* stream) {
  return dwconv_impl(x, wgt, bias, y, n, h, w, c, k, stride, pad, dil, x_pitch, y_pitch, w_pitch, act, stream, true);
}

extern "C" int eqxv_eltwise_bf16(const void* x, const float* scale, const float* shift, const void* other,
                                 const void* gate, void* y, int64_t rows, int32_t c, int32_t x_pitch,
                                 int32_t other_pitch, int32_t gate_pitch, int32_t y_pitch,
                                 int32_t rows_per_image, int32_t act, void* stream) {
  EQXV_CHECK_ARG(x && y && rows > 0 && c > 0, "eltwise: bad arguments");
  EQXV_CHECK_ARG((scale == nullptr) == (shift == nullptr), "eltwise: scale and shift go together");
  EQXV_CHECK_ARG(c % 8 == 0 && x_pitch % 8 == 0 && y_pitch % 8 == 0 && x_pitch >= c && y_pitch >= c,
                 "eltwise: channels/pitches must be multiples of 8");
  if (other) EQXV_CHECK_ARG(other_pitch % 8 == 0 && other_pitch >= c, "eltwise: bad other pitch");
  if (gate) EQXV_CHECK_ARG(gate_pitch % 8 == 0 && gate_pitch >= c && rows_per_image > 0, "eltwise: bad gate");
  const long long total = rows * (c / 8);
  EQXV_CUDA(launch_kernel(eltwise_kernel, dim3(grid_for(total)), dim3(kDwThreads), (size_t)(0), (cudaStream_t)stream, 
      (const __nv_bfloat16*)x, scale, shift, (const __nv_bfloat16*)other, (const __nv_bfloat16*)gate,
      (__nv_bfloat16*)y, rows, c, x_pitch, other_pitch, gate_pitch, y_pitch, rows_per_image > 0 ? rows_per_image : 1,
      act));
  EQXV_CUDA(cudaGetLastError());
  return EQXV_OK;
}

extern "C" int eqxv_copy2d_async(void* dst, int64_t dst_pitch_bytes, const void* src, int64_t src_pitch_bytes,
                                 int64_t width_bytes, int64_t rows, void* stream) {
  EQXV_CHECK_ARG(dst && src && width_bytes > 0 && rows > 0, "copy2d: bad arguments");
  EQXV_CUDA(cudaMemcpy2DAsync(dst, (size_t)dst_pitch_bytes, src, (size_t)src_pitch_bytes, (size_t)width_bytes,
                              (size_t)rows, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return EQXV_OK;
}
