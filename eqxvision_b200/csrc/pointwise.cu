// HBM-bound helper kernels: layout changes at the boundary, pooling, LayerNorm, ViT token glue.
// All of them move 16-byte vectors (8 bf16 channels) per thread over channels-last data so that a
// warp touches whole 128-byte lines; none of them uses shared memory (no reuse to exploit) except
// the NHWC->NCHW transpose.
#include <cstdlib>

#include <cooperative_groups.h>

#include "common.h"
#include "ptx.cuh"

namespace eqxv {

constexpr int kPwThreads = 256;

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

static inline int grid_for(long long work, int threads = kPwThreads) {
  long long b = (work + threads - 1) / threads;
  const long long cap = (long long)device_sm_count() * 32;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

struct alignas(16) bf16x8 {
  __nv_bfloat162 v[4];
};

__device__ __forceinline__ void unpack8(const bf16x8& a, float* f) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(a.v[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ bf16x8 pack8(const float* f) {
  bf16x8 r;
#pragma unroll
  for (int i = 0; i < 4; ++i) r.v[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return r;
}

// ---------------------------------------------------------------------------------------------
// stem input: fp32 NCHW [n,c<=8,h,w] -> bf16 [n, h+2*pad, w+8, 8], image at (pad,pad), zeros elsewhere
// ---------------------------------------------------------------------------------------------
constexpr int kPackRows = 8;
__global__ void pack_stem_kernel(const float* __restrict__ x, bf16x8* __restrict__ y, int n, int c, int h,
                                 int w, int pad) {
  griddep_wait();   // PDL: the predecessor kernel has completed (ptx.cuh)
  griddep_launch();
  // grid = (column blocks, padded rows, images): no index division at all (with a linear index the
  // kernel was ALU bound on the div/mod decomposition: issue slots 80 % busy, profiles/r01_ncu_stem_trio_v15.txt)
  // Every thread walks kPackRows consecutive padded rows of its column: 58 880 one-row blocks of the ResNet-50 batch spent
  // their time being scheduled (ncu: issue slots 83 % busy at 47 % of the DRAM throughput, ~180 warp instructions per
  // 32 pixels); the loads of all rows are issued before the first conversion.
  const int wp = w + 8, hp = h + 2 * pad;
  const int pw = blockIdx.x * blockDim.x + threadIdx.x;
  const int ph0 = blockIdx.y * kPackRows, img = blockIdx.z;
  if (pw >= wp) return;
  const int sw = pw - pad;
  const bool col_ok = sw >= 0 && sw < w;
  const long long plane = (long long)h * w;
  const float* src0 = x + (long long)img * c * plane + (col_ok ? sw : 0);
  float f[kPackRows][3];
#pragma unroll
  for (int r = 0; r < kPackRows; ++r) {
    const int sh = ph0 + r - pad;
    const bool ok = col_ok && sh >= 0 && sh < h;
    const float* src = src0 + (long long)(ok ? sh : 0) * w;
#pragma unroll
    for (int q = 0; q < 3; ++q) f[r][q] = (ok && q < c) ? __ldg(src + q * plane) : 0.f;
  }
  bf16x8* dst = y + ((long long)img * hp + ph0) * wp + pw;
  if (c <= 3) {
#pragma unroll
    for (int r = 0; r < kPackRows; ++r) {
      if (ph0 + r < hp) {
        const float g[8] = {f[r][0], f[r][1], f[r][2], 0.f, 0.f, 0.f, 0.f, 0.f};
        dst[(long long)r * wp] = pack8(g);
      }
    }
    return;
  }
  for (int r = 0; r < kPackRows; ++r) {   // 4..8 input channels (not an image; kept general)
    if (ph0 + r >= hp) break;
    float g[8] = {f[r][0], f[r][1], f[r][2], 0.f, 0.f, 0.f, 0.f, 0.f};
    const int sh = ph0 + r - pad;
    if (col_ok && sh >= 0 && sh < h)
      for (int q = 3; q < 8; ++q)
        if (q < c) g[q] = __ldg(src0 + (long long)sh * w + q * plane);
    dst[(long long)r * wp] = pack8(g);
  }
}

// Pixel-pair layout for stride-2 first layers with <= 4 input channels (eqxv_conv_stem_c4_bf16): fp32 NCHW [n,c<=4,h,w] ->
// bf16 [n, h+2*pad, (w+8)/2, 8]: one 16-byte unit = padded columns (2u, 2u+1) x 4 channels, image at (pad, pad), zeros
// elsewhere. Half the bytes of the 8-channel layout, and the stride-2 convolution walks the units with stride 1.
__global__ void pack_stem_c4_kernel(const float* __restrict__ x, bf16x8* __restrict__ y, int n, int c, int h, int w,
                                    int pad) {
  griddep_wait();   // PDL: the predecessor kernel has completed (ptx.cuh)
  griddep_launch();
  const int wu = (w + 8) / 2, hp = h + 2 * pad;
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  const int ph0 = blockIdx.y * kPackRows, img = blockIdx.z;
  if (u >= wu) return;
  const long long plane = (long long)h * w;
  const float* src0 = x + (long long)img * c * plane;
  float f[kPackRows][8];
#pragma unroll
  for (int r = 0; r < kPackRows; ++r) {
    const int sh = ph0 + r - pad;
    const bool row_ok = sh >= 0 && sh < h;
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int sw = 2 * u + e - pad;
      const bool ok = row_ok && sw >= 0 && sw < w;
      const float* src = src0 + (long long)(ok ? sh : 0) * w + (ok ? sw : 0);
#pragma unroll
      for (int q = 0; q < 4; ++q) f[r][4 * e + q] = (ok && q < c) ? __ldg(src + q * plane) : 0.f;
    }
  }
  bf16x8* dst = y + ((long long)img * hp + ph0) * wu + u;
#pragma unroll
  for (int r = 0; r < kPackRows; ++r)
    if (ph0 + r < hp) dst[(long long)r * wu] = pack8(f[r]);
}

// fp32 NCHW -> bf16 NHWC (channels padded with zeros to c_pad)
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, bf16x8* __restrict__ y, int n, int c,
                                    int h, int w, int c_pad) {
  griddep_wait();   // PDL: the predecessor kernel has completed (ptx.cuh)
  griddep_launch();
  const int groups = c_pad / 8;
  const long long plane = (long long)h * w;
  const long long total = (long long)n * plane * groups;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    // pixel index fastest so that a warp reads 32 consecutive floats of one plane
    const long long pix = i % plane;
    const int g = (int)((i / plane) % groups);
    const int img = (int)(i / (plane * groups));
    float f[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int ch = g * 8 + q;
      f[q] = ch < c ? __ldg(x + ((long long)img * c + ch) * plane + pix) : 0.f;
    }
    y[((long long)img * plane + pix) * groups + g] = pack8(f);
  }
}

// bf16 NHWC (pitch) -> fp32 NCHW through a 32x33 shared tile (pixels x channels)
__global__ void nhwc_to_nchw_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ y, int c,
                                    long long plane, int pitch) {
  griddep_wait();   // PDL: the predecessor kernel has completed (ptx.cuh)
  griddep_launch();
  __shared__ float tile[32][33];
  const int img = blockIdx.z;
  const long long p0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    const long long pix = p0 + r;
    const int ch = c0 + tx;
    float v = 0.f;
    if (pix < plane && ch < c) v = __bfloat162float(x[((long long)img * plane + pix) * pitch + ch]);
    tile[r][tx] = v;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int ch = c0 + r;
    const long long pix = p0 + tx;
    if (pix < plane && ch < c) y[((long long)img * c + ch) * plane + pix] = tile[tx][r];
  }
}

// ---------------------------------------------------------------------------------------------
// pooling (channels-last, 8 channels per thread)
// ---------------------------------------------------------------------------------------------
template <bool kMax>
__global__ void pool2d_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                              int n, int h, int w, int c, int kh, int kw, int sh, int sw, int pad,
                              int ho, int wo, int xp, int yp) {
  griddep_wait();   // PDL: the predecessor kernel has completed (ptx.cuh)
  griddep_launch();
  const int groups = c / 8;
  const long long total = (long long)n * ho * wo * groups;
  const float inv = 1.f / (float)(kh * kw);
  auto body = [&](auto i) {
    const int g = (int)(i % groups);
    auto t = i / groups;
    const int ow = (int)(t % wo);
    t /= wo;
    const int oh = (int)(t % ho);
    const int img = (int)(t / ho);
    float acc[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[q] = kMax ? -INFINITY : 0.f;
    for (int r = 0; r < kh; ++r) {
      const int ih = oh * sh - pad + r;
      if (ih < 0 || ih >= h) continue;
      for (int s = 0; s < kw; ++s) {
        const int iw = ow * sw - pad + s;
        if (iw < 0 || iw >= w) continue;
        const bf16x8 v = *reinterpret_cast<const bf16x8*>(
            x + (((long long)img * h + ih) * w + iw) * xp + g * 8);
        float f[8];
        unpack8(v, f);
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[q] = kMax ? fmaxf(acc[q], f[q]) : acc[q] + f[q];
      }
    }
    if (!kMax) {
#pragma unroll
      for (int q = 0; q < 8; ++q) acc[q] *= inv;
    }
    *reinterpret_cast<bf16x8*>(y + (((long long)img * ho + oh) * wo + ow) * yp + g * 8) = pack8(acc);
  };
  EQXV_GRID_STRIDE(total, body);
}

// Compile-time window/stride variant: all K*K 16-byte loads of a thread are issued before any is
// consumed (memory-level parallelism), out-of-range taps are predicated instead of branched.
// The linear grid-stride order (channel vector, column, row fastest-to-slowest, <= 32 CTAs per SM) keeps
// vertically overlapping windows in the same wave, i.e. in L2: a (column block, row, image) 3-D grid without
// any index division measured 163 us against 121 us on the ResNet-50 max-pool (B200, ncu).
template <bool kMax, int K, int S>
__global__ void pool2d_fixed_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                    int n, int h, int w, int c, int pad, int ho, int wo, int xp, int yp) {
  griddep_wait();   // PDL: the predecessor kernel has completed (ptx.cuh)
  griddep_launch();
  const int groups = c / 8;
  const long long total = (long long)n * ho * wo * groups;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % groups);
    long long t = i / groups;
    const int ow = (int)(t % wo);
    t /= wo;
    const int oh = (int)(t % ho);
    const int img = (int)(t / ho);
    uint4 raw[K * K];
    bool ok[K * K];
#pragma unroll
    for (int r = 0; r < K; ++r) {
#pragma unroll
      for (int q = 0; q < K; ++q) {
        const int ih = oh * S - pad + r, iw = ow * S - pad + q;
        ok[r * K + q] = ih >= 0 && ih < h && iw >= 0 && iw < w;
        const int ihc = min(max(ih, 0), h - 1), iwc = min(max(iw, 0), w - 1);
        raw[r * K + q] = __ldg(reinterpret_cast<const uint4*>(
            x + (((long long)img * h + ihc) * w + iwc) * xp + g * 8));
      }
    }
    if constexpr (kMax) {
      // Out-of-range taps were fetched from the clamped coordinate, which lies inside the same window (2 pad <= k), so
      // they repeat an in-range tap and cannot change the maximum: no predicate, and the maximum itself on packed
      // bf16 pairs (HMNMX2, exact) - 32 instructions per output vector instead of ~150 unpack + compare + select +
      // repack (ncu: the kernel was ISSUE bound, 71 % of the issue slots busy at 50 % of the DRAM throughput).
      uint4 m = raw[0];
#pragma unroll
      for (int k = 1; k < K * K; ++k) {
        const __nv_bfloat162* a2 = reinterpret_cast<const __nv_bfloat162*>(&m);
        const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&raw[k]);
        uint4 o;
        __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
        for (int q = 0; q < 4; ++q) o2[q] = __hmax2(a2[q], b2[q]);
        m = o;
      }
      *reinterpret_cast<uint4*>(y + (((long long)img * ho + oh) * wo + ow) * yp + g * 8) = m;
      continue;
    }
    float acc[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[q] = kMax ? -INFINITY : 0.f;
#pragma unroll
    for (int k = 0; k < K * K; ++k) {
      float f[8];
      unpack8(*reinterpret_cast<const bf16x8*>(&raw[k]), f);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        if (kMax) {
          acc[q] = ok[k] ? fmaxf(acc[q], f[q]) : acc[q];
        } else {
          acc[q] += ok[k] ? f[q] : 0.f;
        }
      }
    }
    if (!kMax) {
#pragma unroll
      for (int q = 0; q < 8; ++q) acc[q] *= 1.f / (float)(K * K);
    }
    *reinterpret_cast<bf16x8*>(y + (((long long)img * ho + oh) * wo + ow) * yp + g * 8) = pack8(acc);
  }
}

// Global average pool (adaptive pool to 1x1: the SE squeeze, the classifier pool, the ASPP pooling
// branch). One block per (image, group of <= 8 channel vectors): the 256 threads are laid out as
// (pixel lane, channel vector) so that each pixel contributes one contiguous <=128-byte read, every
// thread strides over the pixels, and the lanes are combined through shared memory.
__global__ void __launch_bounds__(256) global_avgpool_kernel(const __nv_bfloat16* __restrict__ x,
                                                             __nv_bfloat16* __restrict__ y, int hw, int c,
                                                             int xp, int yp) {
  griddep_wait();   // PDL: the predecessor kernel has completed (ptx.cuh)
  griddep_launch();
  __shared__ float red[256][9];
  const int groups = c / 8;
  const int g0 = blockIdx.x * 8;
  const int gc = min(8, groups - g0);          // channel vectors handled by this block
  const int lanes = 256 / gc;                  // pixel lanes
  const int img = blockIdx.y;
  const int gi = threadIdx.x % gc, lane = threadIdx.x / gc;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (lane < lanes) {
    const __nv_bfloat16* base = x + (long long)img * hw * xp + (g0 + gi) * 8;
    // four independent 16-byte loads in flight per thread: with one, the SE squeeze of EfficientNet-B4
    // streamed at 1.8 TB/s (latency-bound, profiles/r01_launch_metrics_efficientnet_b4.txt)
    int p = lane;
    for (; p + 3 * lanes < hw; p += 4 * lanes) {
      bf16x8 r[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) r[u] = *reinterpret_cast<const bf16x8*>(base + (long long)(p + u * lanes) * xp);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float f[8];
        unpack8(r[u], f);
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[q] += f[q];
      }
    }
    for (; p < hw; p += lanes) {
      float f[8];
      unpack8(*reinterpret_cast<const bf16x8*>(base + (long long)p * xp), f);
#pragma unroll
      for (int q = 0; q < 8; ++q) acc[q] += f[q];
    }
  }
#pragma unroll
  for (int q = 0; q < 8; ++q) red[threadIdx.x][q] = acc[q];
  __syncthreads();
  if (threadIdx.x < gc) {
    float tot[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int l = 0; l < lanes; ++l) {
#pragma unroll
      for (int q = 0; q < 8; ++q) tot[q] += red[l * gc + threadIdx.x][q];
    }
    const float inv = 1.f / (float)hw;
#pragma unroll
    for (int q = 0; q < 8; ++q) tot[q] *= inv;
    *reinterpret_cast<bf16x8*>(y + (long long)img * yp + (g0 + threadIdx.x) * 8) = pack8(tot);
  }
}

// Small maps (the 7x7 classifier pool of ResNet / EfficientNet heads): one thread per (image, channel vector),
// consecutive threads = consecutive channel vectors, so every pixel is one coalesced row read and there is no
// cross-thread reduction at all (the block kernel above left most of its 256 threads idle on 49 pixels).
__global__ void __launch_bounds__(256) global_avgpool_small_kernel(const __nv_bfloat16* __restrict__ x,
                                                                   __nv_bfloat16* __restrict__ y, int hw, int c,
                                                                   int xp, int yp) {
  griddep_wait();
  griddep_launch();
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  const int img = blockIdx.y;
  if (g >= c / 8) return;
  const __nv_bfloat16* base = x + (long long)img * hw * xp + g * 8;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  int p = 0;
  for (; p + 6 < hw; p += 7) {   // seven 16-byte loads in flight per thread (a 7x7 map is seven rounds)
    bf16x8 r[7];
#pragma unroll
    for (int u = 0; u < 7; ++u) r[u] = *reinterpret_cast<const bf16x8*>(base + (long long)(p + u) * xp);
#pragma unroll
    for (int u = 0; u < 7; ++u) {
      float f[8];
      unpack8(r[u], f);
#pragma unroll
      for (int q = 0; q < 8; ++q) acc[q] += f[q];
    }
  }
  for (; p < hw; ++p) {
    float f[8];
    unpack8(*reinterpret_cast<const bf16x8*>(base + (long long)p * xp), f);
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[q] += f[q];
  }
  const float inv = 1.f / (float)hw;
#pragma unroll
  for (int q = 0; q < 8; ++q) acc[q] *= inv;
  *reinterpret_cast<bf16x8*>(y + (long long)img * yp + g * 8) = pack8(acc);
}

// Cluster variant for large maps: the pixels of one (image, 64-channel slab) are split over the 8 CTAs of a
// thread-block cluster; every CTA reduces its slice as above, the partial sums meet in the leader through
// distributed shared memory in rank order (deterministic, no atomics, no scratch buffer). One CTA per slab
// left the SE squeeze of EfficientNet-B4's 112x112 / 56x56 maps latency bound at ~1 TB/s.
constexpr int kPoolCluster = 8;
__global__ void __cluster_dims__(kPoolCluster, 1, 1) __launch_bounds__(256)
    global_avgpool_cluster_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int hw, int c,
                                  int xp, int yp) {
  griddep_wait();
  griddep_launch();
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ float red[256][9];
  __shared__ float part[8][8];                 // this CTA's partial sums: [channel vector][8 channels]
  const int rank = (int)cluster.block_rank();
  const int groups = c / 8;
  const int g0 = (blockIdx.x / kPoolCluster) * 8;
  const int gc = min(8, groups - g0);
  const int lanes = 256 / gc;
  const int img = blockIdx.y;
  const int gi = threadIdx.x % gc, lane = threadIdx.x / gc;
  const int per = (hw + kPoolCluster - 1) / kPoolCluster;
  const int p_lo = rank * per, p_hi = min(hw, p_lo + per);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (lane < lanes) {
    const __nv_bfloat16* base = x + (long long)img * hw * xp + (g0 + gi) * 8;
    int p = p_lo + lane;
    for (; p + 3 * lanes < p_hi; p += 4 * lanes) {
      bf16x8 r[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) r[u] = *reinterpret_cast<const bf16x8*>(base + (long long)(p + u * lanes) * xp);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float f[8];
        unpack8(r[u], f);
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[q] += f[q];
      }
    }
    for (; p < p_hi; p += lanes) {
      float f[8];
      unpack8(*reinterpret_cast<const bf16x8*>(base + (long long)p * xp), f);
#pragma unroll
      for (int q = 0; q < 8; ++q) acc[q] += f[q];
    }
  }
#pragma unroll
  for (int q = 0; q < 8; ++q) red[threadIdx.x][q] = acc[q];
  __syncthreads();
  if (threadIdx.x < gc) {
    float tot[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int l = 0; l < lanes; ++l) {
#pragma unroll
      for (int q = 0; q < 8; ++q) tot[q] += red[l * gc + threadIdx.x][q];
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) part[threadIdx.x][q] = tot[q];
  }
  cluster.sync();                              // every CTA's partial sums are visible cluster-wide
  if (rank == 0 && threadIdx.x < gc) {
    float tot[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int r = 0; r < kPoolCluster; ++r) {   // fixed order: bitwise reproducible
      const float* rp = cluster.map_shared_rank(&part[0][0], r) + threadIdx.x * 8;
#pragma unroll
      for (int q = 0; q < 8; ++q) tot[q] += rp[q];
    }
    const float inv = 1.f / (float)hw;
#pragma unroll
    for (int q = 0; q < 8; ++q) tot[q] *= inv;
    *reinterpret_cast<bf16x8*>(y + (long long)img * yp + (g0 + threadIdx.x) * 8) = pack8(tot);
  }
  cluster.sync();                              // nobody leaves while the leader still reads its shared memory
}

// ---------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, the row is kept in registers (d <= 2048), fp32 statistics,
// two-pass (mean, then centred variance) as equinox.nn.LayerNorm does.
// ---------------------------------------------------------------------------------------------
constexpr int kLnMaxVec = 8;  // 8 vectors x 8 elements x 32 lanes = 2048

__global__ void layernorm_kernel(const __nv_bfloat16* __restrict__ x, long long ldx,
                                 const float* __restrict__ gamma, const float* __restrict__ beta,
                                 __nv_bfloat16* __restrict__ y, long long ldy, long long rows, int d,
                                 float eps) {
  griddep_wait();   // PDL: the predecessor kernel has completed (ptx.cuh)
  griddep_launch();
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const int nvec = d / 8;
  for (long long row = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); row < rows;
       row += (long long)gridDim.x * wpb) {
    const __nv_bfloat16* xr = x + row * ldx;
    float f[kLnMaxVec][8];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < kLnMaxVec; ++i) {
      const int v = lane + i * 32;
      if (v < nvec) {
        unpack8(*reinterpret_cast<const bf16x8*>(xr + v * 8), f[i]);
#pragma unroll
        for (int q = 0; q < 8; ++q) sum += f[i][q];
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum / (float)d;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < kLnMaxVec; ++i) {
      const int v = lane + i * 32;
      if (v < nvec) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float dlt = f[i][q] - mean;
          sq += dlt * dlt;
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq / (float)d + eps);
    __nv_bfloat16* yr = y + row * ldy;
#pragma unroll
    for (int i = 0; i < kLnMaxVec; ++i) {
      const int v = lane + i * 32;
      if (v < nvec) {
        float o[8];
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + v * 8));
        const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + v * 8 + 4));
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + v * 8));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + v * 8 + 4));
        const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int q = 0; q < 8; ++q) o[q] = (f[i][q] - mean) * rstd * gg[q] + bb[q];
        *reinterpret_cast<bf16x8*>(yr + v * 8) = pack8(o);
      }
    }
  }
}


// d == NV * 256: every lane owns exactly NV 16-byte vectors, the row lives in NV*8 registers and the NEXT
// row of this warp is already in flight while the current one is reduced (one warp handles ~1.3 rows of
// a ViT-B/16 token matrix: without the prefetch the kernel is bound by two dependent load latencies).
template <int NV>
__global__ void __launch_bounds__(256) layernorm_fixed_kernel(const __nv_bfloat16* __restrict__ x, long long ldx,
                                                              const float* __restrict__ gamma,
                                                              const float* __restrict__ beta,
                                                              __nv_bfloat16* __restrict__ y, long long ldy,
                                                              long long rows, float eps) {
  constexpr int D = NV * 256;
  const int lane = threadIdx.x & 31;
  const long long wstride = (long long)gridDim.x * (blockDim.x >> 5);
  long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  // affine parameters of this lane's columns (same for every row; constants, so they are fetched
  // BEFORE the PDL wait and overlap the tail of the kernel that produces x)
  float gg[NV][8], bb[NV][8];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (lane + i * 32) * 8;
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + c + 4));
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c)), b1 = __ldg(reinterpret_cast<const float4*>(beta + c + 4));
    gg[i][0] = g0.x, gg[i][1] = g0.y, gg[i][2] = g0.z, gg[i][3] = g0.w, gg[i][4] = g1.x, gg[i][5] = g1.y, gg[i][6] = g1.z, gg[i][7] = g1.w;
    bb[i][0] = b0.x, bb[i][1] = b0.y, bb[i][2] = b0.z, bb[i][3] = b0.w, bb[i][4] = b1.x, bb[i][5] = b1.y, bb[i][6] = b1.z, bb[i][7] = b1.w;
  }
  griddep_wait();   // PDL: the predecessor kernel has completed (ptx.cuh)
  griddep_launch();
  if (row >= rows) return;
  bf16x8 cur[NV], nxt[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) cur[i] = *reinterpret_cast<const bf16x8*>(x + row * ldx + (lane + i * 32) * 8);
  for (; row < rows; row += wstride) {
    const long long nrow = row + wstride;
    if (nrow < rows) {
#pragma unroll
      for (int i = 0; i < NV; ++i) nxt[i] = *reinterpret_cast<const bf16x8*>(x + nrow * ldx + (lane + i * 32) * 8);
    }
    float f[NV][8];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      unpack8(cur[i], f[i]);
#pragma unroll
      for (int q = 0; q < 8; ++q) sum += f[i][q];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum * (1.f / (float)D);
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        f[i][q] -= mean;
        sq = fmaf(f[i][q], f[i][q], sq);
      }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq * (1.f / (float)D) + eps);
    __nv_bfloat16* yr = y + row * ldy;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float o[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) o[q] = fmaf(f[i][q] * rstd, gg[i][q], bb[i][q]);
      *reinterpret_cast<bf16x8*>(yr + (lane + i * 32) * 8) = pack8(o);
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) cur[i] = nxt[i];
  }
}

// ---------------------------------------------------------------------------------------------
// ViT glue
// ---------------------------------------------------------------------------------------------
// rows[(img*gh + gy)*gw + gx, (ch*p + py)*p + px] = x[img, ch, gy*p+py, gx*p+px]
__global__ void patchify_kernel(const float* __restrict__ x, bf16x8* __restrict__ rows, int n, int c,
                                int h, int w, int p) {
  griddep_wait();   // PDL: the predecessor kernel has completed (ptx.cuh)
  griddep_launch();
  const int gh = h / p, gw = w / p;
  const int kvec = c * p * p / 8;
  const int pv = p / 8;  // vectors per patch row
  const long long total = (long long)n * gh * gw * kvec;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int kv = (int)(i % kvec);
    const long long prow = i / kvec;
    const int gx = (int)(prow % gw);
    const int gy = (int)((prow / gw) % gh);
    const int img = (int)(prow / ((long long)gw * gh));
    const int px0 = (kv % pv) * 8;
    const int py = (kv / pv) % p;
    const int ch = kv / (pv * p);
    const float* src = x + (((long long)img * c + ch) * h + gy * p + py) * w + gx * p + px0;
    const float4 a = __ldg(reinterpret_cast<const float4*>(src));
    const float4 b = __ldg(reinterpret_cast<const float4*>(src + 4));
    const float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    rows[i] = pack8(f);
  }
}

__global__ void assemble_tokens_kernel(const bf16x8* __restrict__ patches, const float* __restrict__ cls,
                                       const float* __restrict__ pos, bf16x8* __restrict__ out, int n,
                                       int np, int d) {
  griddep_wait();   // PDL: the predecessor kernel has completed (ptx.cuh)
  griddep_launch();
  const int dv = d / 8;
  const long long total = (long long)n * (np + 1) * dv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % dv);
    const long long r = i / dv;
    const int t = (int)(r % (np + 1));
    const int img = (int)(r / (np + 1));
    float f[8];
    if (t == 0) {
#pragma unroll
      for (int q = 0; q < 8; ++q) f[q] = __ldg(cls + v * 8 + q);
    } else {
      unpack8(patches[((long long)img * np + (t - 1)) * dv + v], f);
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) f[q] += __ldg(pos + (long long)t * d + v * 8 + q);
    out[i] = pack8(f);
  }
}

__global__ void gather_rows_kernel(const __nv_bfloat16* __restrict__ x, long long ldx,
                                   __nv_bfloat16* __restrict__ y, long long ldy, int n, int tokens,
                                   int row, int d) {
  griddep_wait();   // PDL: the predecessor kernel has completed (ptx.cuh)
  griddep_launch();
  const int dv = d / 8;
  const long long total = (long long)n * dv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % dv);
    const int img = (int)(i / dv);
    *reinterpret_cast<bf16x8*>(y + (long long)img * ldy + v * 8) =
        *reinterpret_cast<const bf16x8*>(x + ((long long)img * tokens + row) * ldx + v * 8);
  }
}

}  // namespace eqxv

using namespace eqxv;

#define EQXV_LAUNCH_CHECK() EQXV_CUDA(cudaGetLastError())

extern "C" int eqxv_pack_stem_input(const float* x, void* xpad, int32_t n, int32_t c, int32_t h, int32_t w,
                                    int32_t pad, void* stream) {
  EQXV_CHECK_ARG(x && xpad && n > 0 && h > 0 && w > 0 && c >= 1 && c <= 8 && pad >= 0 && pad <= 4,
                 "pack_stem_input: bad arguments");
  EQXV_CHECK_ARG(h + 2 * pad <= 65535 && n <= 65535, "pack_stem_input: image too tall / batch too large");
  EQXV_CUDA(launch_kernel(pack_stem_kernel,
                          dim3((unsigned)((w + 8 + 127) / 128), (unsigned)((h + 2 * pad + kPackRows - 1) / kPackRows), (unsigned)n),
                          dim3(128), (size_t)0, (cudaStream_t)stream, x, reinterpret_cast<bf16x8*>(xpad), n, c, h, w,
                          pad));
  EQXV_LAUNCH_CHECK();
  return EQXV_OK;
}

extern "C" int eqxv_pack_stem_input_c4(const float* x, void* xpad4, int32_t n, int32_t c, int32_t h, int32_t w, int32_t pad,
                                       void* stream) {
  EQXV_CHECK_ARG(x && xpad4 && n > 0 && h > 0 && w > 0 && w % 2 == 0 && c >= 1 && c <= 4 && pad >= 0 && pad <= 4,
                 "pack_stem_input_c4: bad arguments (<= 4 channels, even width)");
  EQXV_CHECK_ARG(h + 2 * pad <= 65535 * kPackRows && n <= 65535, "pack_stem_input_c4: image too tall / batch too large");
  EQXV_CUDA(launch_kernel(pack_stem_c4_kernel,
                          dim3((unsigned)(((w + 8) / 2 + 127) / 128), (unsigned)((h + 2 * pad + kPackRows - 1) / kPackRows),
                               (unsigned)n),
                          dim3(128), (size_t)0, (cudaStream_t)stream, x, reinterpret_cast<bf16x8*>(xpad4), n, c, h, w, pad));
  EQXV_LAUNCH_CHECK();
  return EQXV_OK;
}

extern "C" int eqxv_nchw_f32_to_nhwc_bf16(const float* x, void* y, int32_t n, int32_t c, int32_t h,
                                          int32_t w, int32_t c_pad, void* stream) {
  EQXV_CHECK_ARG(x && y && n > 0 && c > 0 && h > 0 && w > 0, "nchw_to_nhwc: bad arguments");
  EQXV_CHECK_ARG(c_pad >= c && c_pad % 8 == 0, "nchw_to_nhwc: c_pad must be a multiple of 8 >= c");
  const long long total = (long long)n * h * w * (c_pad / 8);
  EQXV_CUDA(launch_kernel(nchw_to_nhwc_kernel, dim3(grid_for(total)), dim3(kPwThreads), (size_t)(0), (cudaStream_t)stream, 
      x, reinterpret_cast<bf16x8*>(y), n, c, h, w, c_pad));
  EQXV_LAUNCH_CHECK();
  return EQXV_OK;
}

extern "C" int eqxv_nhwc_bf16_to_nchw_f32(const void* x, float* y, int32_t n, int32_t c, int32_t h,
                                          int32_t w, int32_t x_pitch, void* stream) {
  EQXV_CHECK_ARG(x && y && n > 0 && c > 0 && h > 0 && w > 0 && x_pitch >= c,
                 "nhwc_to_nchw: bad arguments");
  const long long plane = (long long)h * w;
  dim3 grid((unsigned)((plane + 31) / 32), (unsigned)((c + 31) / 32), (unsigned)n);
  EQXV_CUDA(launch_kernel(nhwc_to_nchw_kernel, dim3(grid), dim3(dim3(32, 8)), (size_t)(0), (cudaStream_t)stream, 
      reinterpret_cast<const __nv_bfloat16*>(x), y, c, plane, x_pitch));
  EQXV_LAUNCH_CHECK();
  return EQXV_OK;
}

static int pool_common(bool is_max, const void* x, void* y, int n, int h, int w, int c, int kh, int kw,
                       int sh, int sw, int pad, int ho, int wo, int xp, int yp, cudaStream_t stream) {
  EQXV_CHECK_ARG(x && y && n > 0 && h > 0 && w > 0 && c > 0, "pool: bad arguments");
  EQXV_CHECK_ARG(c % 8 == 0 && xp % 8 == 0 && yp % 8 == 0 && xp >= c && yp >= c,
                 "pool: channels and pitches must be multiples of 8");
  EQXV_CHECK_ARG(ho > 0 && wo > 0, "pool: empty output");
  const long long total = (long long)n * ho * wo * (c / 8);
  if (kh == kw && sh == sw && kh == 3 && sh == 2) {
    if (is_max)
      EQXV_CUDA(launch_kernel(pool2d_fixed_kernel<true, 3, 2>, dim3(grid_for(total)), dim3(kPwThreads), (size_t)0, stream,
                              (const __nv_bfloat16*)x, (__nv_bfloat16*)y, n, h, w, c, pad, ho, wo, xp, yp));
    else
      EQXV_CUDA(launch_kernel(pool2d_fixed_kernel<false, 3, 2>, dim3(grid_for(total)), dim3(kPwThreads), (size_t)0, stream,
                              (const __nv_bfloat16*)x, (__nv_bfloat16*)y, n, h, w, c, pad, ho, wo, xp, yp));
    EQXV_LAUNCH_CHECK();
    return EQXV_OK;
  }
  if (kh == kw && sh == sw && kh == 2 && sh == 2) {
    if (is_max)
      EQXV_CUDA(launch_kernel(pool2d_fixed_kernel<true, 2, 2>, dim3(grid_for(total)), dim3(kPwThreads), (size_t)0, stream,
                              (const __nv_bfloat16*)x, (__nv_bfloat16*)y, n, h, w, c, pad, ho, wo, xp, yp));
    else
      EQXV_CUDA(launch_kernel(pool2d_fixed_kernel<false, 2, 2>, dim3(grid_for(total)), dim3(kPwThreads), (size_t)0, stream,
                              (const __nv_bfloat16*)x, (__nv_bfloat16*)y, n, h, w, c, pad, ho, wo, xp, yp));
    EQXV_LAUNCH_CHECK();
    return EQXV_OK;
  }
  if (is_max) {
    EQXV_CUDA(launch_kernel(pool2d_kernel<true>, dim3(grid_for(total)), dim3(kPwThreads), (size_t)(0), stream, 
        (const __nv_bfloat16*)x, (__nv_bfloat16*)y, n, h, w, c, kh, kw, sh, sw, pad, ho, wo, xp, yp));
  } else {
    EQXV_CUDA(launch_kernel(pool2d_kernel<false>, dim3(grid_for(total)), dim3(kPwThreads), (size_t)(0), stream, 
        (const __nv_bfloat16*)x, (__nv_bfloat16*)y, n, h, w, c, kh, kw, sh, sw, pad, ho, wo, xp, yp));
  }
  EQXV_LAUNCH_CHECK();
  return EQXV_OK;
}

extern "C" int eqxv_maxpool2d_nhwc_bf16(const void* x, void* y, int32_t n, int32_t h, int32_t w,
                                        int32_t c, int32_t k, int32_t stride, int32_t pad,
                                        int32_t x_pitch, int32_t y_pitch, void* stream) {
  EQXV_CHECK_ARG(k >= 1 && stride >= 1 && pad >= 0 && 2 * pad <= k, "maxpool: bad window");
  const int ho = (h + 2 * pad - k) / stride + 1;
  const int wo = (w + 2 * pad - k) / stride + 1;
  return pool_common(true, x, y, n, h, w, c, k, k, stride, stride, pad, ho, wo, x_pitch, y_pitch,
                     (cudaStream_t)stream);
}

// ceil-mode output size with torch's rule (the last window must start inside the input or its left padding);
// that is the arithmetic the reference's own SqueezeNet / GoogLeNet tests pin against torchvision
static int pool_out_ceil(int size, int k, int stride, int pad) {
  int o = (size + 2 * pad - k + stride - 1) / stride + 1;
  if ((o - 1) * stride >= size + pad) --o;
  return o;
}

extern "C" int eqxv_maxpool2d_ceil_nhwc_bf16(const void* x, void* y, int32_t n, int32_t h, int32_t w,
                                             int32_t c, int32_t k, int32_t stride, int32_t pad,
                                             int32_t x_pitch, int32_t y_pitch, void* stream) {
  EQXV_CHECK_ARG(k >= 1 && stride >= 1 && pad >= 0 && 2 * pad <= k, "maxpool(ceil): bad window");
  EQXV_CHECK_ARG(h + 2 * pad >= k && w + 2 * pad >= k, "maxpool(ceil): window larger than the padded input");
  // windows are clipped to the input by the kernels, so the partial windows at the bottom / right edge need no
  // extra padding: only the output extent changes
  return pool_common(true, x, y, n, h, w, c, k, k, stride, stride, pad, pool_out_ceil(h, k, stride, pad),
                     pool_out_ceil(w, k, stride, pad), x_pitch, y_pitch, (cudaStream_t)stream);
}

extern "C" int eqxv_avgpool2d_nhwc_bf16(const void* x, void* y, int32_t n, int32_t h, int32_t w,
                                        int32_t c, int32_t k, int32_t stride, int32_t x_pitch,
                                        int32_t y_pitch, void* stream) {
  EQXV_CHECK_ARG(k >= 1 && stride >= 1, "avgpool: bad window");
  const int ho = (h - k) / stride + 1;
  const int wo = (w - k) / stride + 1;
  return pool_common(false, x, y, n, h, w, c, k, k, stride, stride, 0, ho, wo, x_pitch, y_pitch,
                     (cudaStream_t)stream);
}

// equinox.nn.AdaptiveAvgPool2d on an axis that does not divide evenly (GoogLeNet's auxiliary heads pool 14x14 -> 4x4,
// googlenet.py:265-268; AlexNet / VGG away from 224 px): Equinox splits the axis into `t` consecutive blocks, the first
// dim % t of them one element longer (dim // t + 1), the rest dim // t - NOT torch's overlapping windows.
__device__ __forceinline__ void eqx_block(int i, int dim, int t, int& start, int& len) {
  const int head = dim % t, block = dim / t;
  if (i < head) {
    start = i * (block + 1), len = block + 1;
  } else {
    start = head * (block + 1) + (i - head) * block, len = block;
  }
}
__global__ void adaptive_avgpool_uneven_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int n,
                                               int h, int w, int c, int oh, int ow, int xp, int yp) {
  griddep_wait();
  griddep_launch();
  const int groups = c / 8;
  const long long total = (long long)n * oh * ow * groups;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % groups);
    long long t = i / groups;
    const int ox = (int)(t % ow);
    t /= ow;
    const int oy = (int)(t % oh);
    const int img = (int)(t / oh);
    int y0, hy, x0, wx;
    eqx_block(oy, h, oh, y0, hy);
    eqx_block(ox, w, ow, x0, wx);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int r = y0; r < y0 + hy; ++r)
      for (int q = x0; q < x0 + wx; ++q) {
        float f[8];
        unpack8(*reinterpret_cast<const bf16x8*>(x + (((long long)img * h + r) * w + q) * xp + g * 8), f);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += f[e];
      }
    const float inv = 1.f / (float)(hy * wx);
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] *= inv;
    *reinterpret_cast<bf16x8*>(y + (((long long)img * oh + oy) * ow + ox) * yp + g * 8) = pack8(acc);
  }
}

extern "C" int eqxv_adaptive_avgpool_nhwc_bf16(const void* x, void* y, int32_t n, int32_t h, int32_t w,
                                               int32_t c, int32_t oh, int32_t ow, int32_t x_pitch,
                                               int32_t y_pitch, void* stream) {
  EQXV_CHECK_ARG(oh >= 1 && ow >= 1, "adaptive_avgpool: bad output size");
  if (h % oh != 0 || w % ow != 0) {
    EQXV_CHECK_ARG(x && y && n > 0 && c > 0 && c % 8 == 0 && x_pitch % 8 == 0 && y_pitch % 8 == 0 && x_pitch >= c &&
                       y_pitch >= c && oh <= h && ow <= w,
                   "adaptive_avgpool: bad arguments (uneven split needs oh <= h, ow <= w, channels a multiple of 8)");
    const long long total = (long long)n * oh * ow * (c / 8);
    EQXV_CUDA(launch_kernel(adaptive_avgpool_uneven_kernel, dim3(grid_for(total)), dim3(kPwThreads), (size_t)0,
                            (cudaStream_t)stream, (const __nv_bfloat16*)x, (__nv_bfloat16*)y, n, h, w, c, oh, ow, x_pitch,
                            y_pitch));
    EQXV_LAUNCH_CHECK();
    return EQXV_OK;
  }
  if (oh == 1 && ow == 1 && h * w >= 32) {
    EQXV_CHECK_ARG(x && y && n > 0 && c > 0 && c % 8 == 0 && x_pitch % 8 == 0 && y_pitch % 8 == 0 &&
                       x_pitch >= c && y_pitch >= c && n <= 65535,
                   "global_avgpool: bad arguments");
    static const bool no_cluster_pool = getenv("EQXV_NO_POOL_CLUSTER") != nullptr;
    if (h * w >= 1024 && !no_cluster_pool) {   // big maps: split the pixels over an 8-CTA cluster
      dim3 cgrid((unsigned)(((c / 8 + 7) / 8) * kPoolCluster), (unsigned)n);
      EQXV_CUDA(launch_kernel(global_avgpool_cluster_kernel, cgrid, dim3(256), (size_t)0, (cudaStream_t)stream,
                              (const __nv_bfloat16*)x, (__nv_bfloat16*)y, h * w, c, x_pitch, y_pitch));
      EQXV_LAUNCH_CHECK();
      return EQXV_OK;
    }
    // small map and enough (image, channel vector) pairs to fill the machine with one thread each (ResNet-50 head,
    // 256 x 7x7 x 2048: 24.8 us against 31-33 us for the block kernel). With fewer pairs or longer pixel loops the
    // serial per-thread sum loses: EfficientNet-B4's 33 SE squeezes (128 images, 14x14 x 672..7x7 x 2688) measured
    // 1.12 ms with this kernel against 0.74 ms with the block kernel (profiles/r01_bench_v24.json vs v26).
    if (h * w <= 64 && (long long)n * (c / 8) >= 65536) {
      // 64-thread blocks: 1024 blocks instead of 256 for the ResNet-50 head (the kernel was latency bound on 256 CTAs)
      dim3 sgrid((unsigned)((c / 8 + 63) / 64), (unsigned)n);
      EQXV_CUDA(launch_kernel(global_avgpool_small_kernel, sgrid, dim3(64), (size_t)0, (cudaStream_t)stream,
                              (const __nv_bfloat16*)x, (__nv_bfloat16*)y, h * w, c, x_pitch, y_pitch));
      EQXV_LAUNCH_CHECK();
      return EQXV_OK;
    }
    dim3 grid((unsigned)((c / 8 + 7) / 8), (unsigned)n);
    EQXV_CUDA(launch_kernel(global_avgpool_kernel, dim3(grid), dim3(256), (size_t)(0), (cudaStream_t)stream, (const __nv_bfloat16*)x, (__nv_bfloat16*)y,
                                                                  h * w, c, x_pitch, y_pitch));
    EQXV_LAUNCH_CHECK();
    return EQXV_OK;
  }
  return pool_common(false, x, y, n, h, w, c, h / oh, w / ow, h / oh, w / ow, 0, oh, ow, x_pitch,
                     y_pitch, (cudaStream_t)stream);
}

extern "C" int eqxv_layernorm_bf16(const void* x, int64_t ldx, const float* gamma, const float* beta,
                                   void* y, int64_t ldy, int64_t rows, int32_t d, float eps,
                                   void* stream) {
  EQXV_CHECK_ARG(x && y && gamma && beta && rows > 0 && d > 0, "layernorm: bad arguments");
  EQXV_CHECK_ARG(d % 8 == 0 && d <= kLnMaxVec * 256, "layernorm: d must be a multiple of 8, <= 2048");
  EQXV_CHECK_ARG(ldx % 8 == 0 && ldy % 8 == 0 && ldx >= d && ldy >= d, "layernorm: bad row pitch");
  const int wpb = 8;
  long long blocks = (rows + wpb - 1) / wpb;
  const long long cap = (long long)device_sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (d % 256 == 0 && d <= 1024 && ldx % 8 == 0 && ldy % 8 == 0) {
    // fixed-width fast path: every warp pipelines several rows (next-row prefetch) and amortises its 6 KiB of
    // gamma/beta loads over them (EQXV_LN_RPW rows per warp, default 8: 21.9k -> 22.1k img/s on ViT-B/16 against 2)
    static const int rpw_env = getenv("EQXV_LN_RPW") ? atoi(getenv("EQXV_LN_RPW")) : 0;
    const int rpw = rpw_env >= 1 && rpw_env <= 64 ? rpw_env : 8;
    long long fb = std::min<long long>((rows + (long long)rpw * wpb - 1) / ((long long)rpw * wpb), cap);
    if (fb < 1) fb = 1;
    const __nv_bfloat16* xb = (const __nv_bfloat16*)x;
    __nv_bfloat16* yb = (__nv_bfloat16*)y;
    switch (d / 256) {
      case 1: EQXV_CUDA(launch_kernel(layernorm_fixed_kernel<1>, dim3((int)fb), dim3(wpb * 32), (size_t)(0), (cudaStream_t)stream, xb, ldx, gamma, beta, yb, ldy, rows, eps)); break;
      case 2: EQXV_CUDA(launch_kernel(layernorm_fixed_kernel<2>, dim3((int)fb), dim3(wpb * 32), (size_t)(0), (cudaStream_t)stream, xb, ldx, gamma, beta, yb, ldy, rows, eps)); break;
      case 3: EQXV_CUDA(launch_kernel(layernorm_fixed_kernel<3>, dim3((int)fb), dim3(wpb * 32), (size_t)(0), (cudaStream_t)stream, xb, ldx, gamma, beta, yb, ldy, rows, eps)); break;
      default: EQXV_CUDA(launch_kernel(layernorm_fixed_kernel<4>, dim3((int)fb), dim3(wpb * 32), (size_t)(0), (cudaStream_t)stream, xb, ldx, gamma, beta, yb, ldy, rows, eps)); break;
    }
    EQXV_LAUNCH_CHECK();
    return EQXV_OK;
  }
  EQXV_CUDA(launch_kernel(layernorm_kernel, dim3((int)blocks), dim3(wpb * 32), (size_t)(0), (cudaStream_t)stream, 
      (const __nv_bfloat16*)x, ldx, gamma, beta, (__nv_bfloat16*)y, ldy, rows, d, eps));
  EQXV_LAUNCH_CHECK();
  return EQXV_OK;
}

extern "C" int eqxv_patchify_nchw_f32_bf16(const float* x, void* rows, int32_t n, int32_t c, int32_t h,
                                           int32_t w, int32_t p, void* stream) {
  EQXV_CHECK_ARG(x && rows && n > 0 && c > 0 && h > 0 && w > 0, "patchify: bad arguments");
  EQXV_CHECK_ARG(p % 8 == 0 && h % p == 0 && w % p == 0 && w % 4 == 0,
                 "patchify: patch size must be a multiple of 8 dividing h and w");
  const long long total = (long long)n * (h / p) * (w / p) * (c * p * p / 8);
  EQXV_CUDA(launch_kernel(patchify_kernel, dim3(grid_for(total)), dim3(kPwThreads), (size_t)(0), (cudaStream_t)stream, 
      x, reinterpret_cast<bf16x8*>(rows), n, c, h, w, p));
  EQXV_LAUNCH_CHECK();
  return EQXV_OK;
}

extern "C" int eqxv_vit_assemble_tokens_bf16(const void* patches, const float* cls, const float* pos,
                                             void* out, int32_t n, int32_t np, int32_t d,
                                             void* stream) {
  EQXV_CHECK_ARG(patches && cls && pos && out && n > 0 && np > 0 && d > 0 && d % 8 == 0,
                 "assemble_tokens: bad arguments");
  const long long total = (long long)n * (np + 1) * (d / 8);
  EQXV_CUDA(launch_kernel(assemble_tokens_kernel, dim3(grid_for(total)), dim3(kPwThreads), (size_t)(0), (cudaStream_t)stream, 
      reinterpret_cast<const bf16x8*>(patches), cls, pos, reinterpret_cast<bf16x8*>(out), n, np, d));
  EQXV_LAUNCH_CHECK();
  return EQXV_OK;
}

extern "C" int eqxv_gather_rows_bf16(const void* x, int64_t ldx, void* y, int64_t ldy, int32_t n,
                                     int32_t tokens, int32_t row, int32_t d, void* stream) {
  EQXV_CHECK_ARG(x && y && n > 0 && tokens > 0 && row >= 0 && row < tokens && d > 0 && d % 8 == 0 &&
                     ldx % 8 == 0 && ldy % 8 == 0,
                 "gather_rows: bad arguments");
  const long long total = (long long)n * (d / 8);
  EQXV_CUDA(launch_kernel(gather_rows_kernel, dim3(grid_for(total)), dim3(kPwThreads), (size_t)(0), (cudaStream_t)stream, 
      (const __nv_bfloat16*)x, ldx, (__nv_bfloat16*)y, ldy, n, tokens, row, d));
  EQXV_LAUNCH_CHECK();
  return EQXV_OK;
}
