// Depthwise convolution and the elementwise passes that cannot ride in a GEMM epilogue.
// HBM-bound, integer-free data paths: channels-last, 16-byte vectors (8 bf16 channels) per thread,
// no tensor cores (K3/K8/K11 of SURVEY.md §2.1).
#include "common.h"
#include "ptx.cuh"

namespace eqxv {

constexpr int kDwThreads = 256;

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
    case EQXV_ACT_SILU: return __fdividef(v, 1.f + __expf(-v));
    case EQXV_ACT_GELU_TANH: {
      const float u = 0.7978845608028654f * (v + 0.044715f * v * v * v);
      float th;
      asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(u));
      return 0.5f * v * (1.f + th);
    }
    case EQXV_ACT_HARDSWISH: return v * fminf(fmaxf(v + 3.f, 0.f), 6.f) * (1.f / 6.f);
    case EQXV_ACT_SIGMOID: return __fdividef(1.f, 1.f + __expf(-v));
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
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % groups);
    long long t = i / groups;
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
  }
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
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % groups);
    long long t = i / groups;
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
  }
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
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % groups);
    const long long row = i / groups;
    float f[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(x + row * xp + g * 8)), f);
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
      unpack8(__ldg(reinterpret_cast<const uint4*>(other + row * op + g * 8)), o);
#pragma unroll
      for (int q = 0; q < 8; ++q) f[q] += o[q];
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) f[q] = act_rt(f[q], act);
    if (gate != nullptr) {
      float s[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(gate + (row / rows_per_image) * gp + g * 8)), s);
#pragma unroll
      for (int q = 0; q < 8; ++q) f[q] *= s[q];
    }
    *reinterpret_cast<uint4*>(y + row * yp + g * 8) = pack8(f);
  }
}

}  // namespace eqxv

using namespace eqxv;

extern "C" int eqxv_dwconv_bn_act_bf16(const void* x, const float* wgt, const float* bias, void* y,
                                       int32_t n, int32_t h, int32_t w, int32_t c, int32_t k,
                                       int32_t stride, int32_t pad, int32_t dil, int32_t x_pitch,
                                       int32_t y_pitch, int32_t w_pitch, int32_t act, void* stream) {
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
