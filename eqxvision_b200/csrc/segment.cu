// Bilinear resize for the segmentation heads (K14): jax.image.resize(method="bilinear") when
// upsampling == half-pixel centres with edge clamping (== torch align_corners=False), reference
// call sites models/segmentation/_utils.py:52,57 and deeplabv3.py:74.
#include "common.h"
#include "ptx.cuh"

namespace eqxv {

__device__ __forceinline__ void src_index(int dst, float scale, int in, int& i0, int& i1, float& t) {
  float s = ((float)dst + 0.5f) * scale - 0.5f;
  s = fmaxf(s, 0.f);
  i0 = min((int)s, in - 1);
  i1 = min(i0 + 1, in - 1);
  t = s - (float)i0;
}

// NHWC bf16 -> NCHW fp32 (the model output handed back to the caller). One thread per output
// element, consecutive threads along the output row: stores are fully coalesced (the output is
// 64x larger than the input for the x8 DeepLab upsample, so the stores are what matters).
__global__ void resize_to_nchw_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ y, int n,
                                      int c, int h, int w, int oh, int ow, int xp, float sh, float sw) {
  griddep_wait();   // PDL: the predecessor kernel has completed (ptx.cuh)
  griddep_launch();
  const long long total = (long long)n * c * oh * ow;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(i % ow);
    long long t = i / ow;
    const int oy = (int)(t % oh);
    t /= oh;
    const int ch = (int)(t % c);
    const int img = (int)(t / c);
    int y0, y1, x0, x1;
    float ty, tx;
    src_index(oy, sh, h, y0, y1, ty);
    src_index(ox, sw, w, x0, x1, tx);
    const __nv_bfloat16* base = x + (long long)img * h * w * xp + ch;
    const float v00 = __bfloat162float(base[((long long)y0 * w + x0) * xp]);
    const float v01 = __bfloat162float(base[((long long)y0 * w + x1) * xp]);
    const float v10 = __bfloat162float(base[((long long)y1 * w + x0) * xp]);
    const float v11 = __bfloat162float(base[((long long)y1 * w + x1) * xp]);
    const float top = v00 + (v01 - v00) * tx;
    const float bot = v10 + (v11 - v10) * tx;
    y[i] = top + (bot - top) * ty;
  }
}

// NHWC bf16 -> NHWC bf16 (ASPP pooling branch: a 1x1 map broadcast to HxW), 8 channels per thread
__global__ void resize_nhwc_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int n,
                                   int c, int h, int w, int oh, int ow, int xp, int yp, float sh, float sw) {
  griddep_wait();   // PDL: the predecessor kernel has completed (ptx.cuh)
  griddep_launch();
  const int groups = c / 8;
  const long long total = (long long)n * oh * ow * groups;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % groups);
    long long t = i / groups;
    const int ox = (int)(t % ow);
    t /= ow;
    const int oy = (int)(t % oh);
    const int img = (int)(t / oh);
    int y0, y1, x0, x1;
    float ty, tx;
    src_index(oy, sh, h, y0, y1, ty);
    src_index(ox, sw, w, x0, x1, tx);
    const __nv_bfloat16* base = x + (long long)img * h * w * xp + g * 8;
    const uint4 r00 = __ldg(reinterpret_cast<const uint4*>(base + ((long long)y0 * w + x0) * xp));
    const uint4 r01 = __ldg(reinterpret_cast<const uint4*>(base + ((long long)y0 * w + x1) * xp));
    const uint4 r10 = __ldg(reinterpret_cast<const uint4*>(base + ((long long)y1 * w + x0) * xp));
    const uint4 r11 = __ldg(reinterpret_cast<const uint4*>(base + ((long long)y1 * w + x1) * xp));
    const __nv_bfloat162* p00 = reinterpret_cast<const __nv_bfloat162*>(&r00);
    const __nv_bfloat162* p01 = reinterpret_cast<const __nv_bfloat162*>(&r01);
    const __nv_bfloat162* p10 = reinterpret_cast<const __nv_bfloat162*>(&r10);
    const __nv_bfloat162* p11 = reinterpret_cast<const __nv_bfloat162*>(&r11);
    __nv_bfloat162 out[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float2 a = __bfloat1622float2(p00[q]), b = __bfloat1622float2(p01[q]);
      const float2 cc = __bfloat1622float2(p10[q]), d = __bfloat1622float2(p11[q]);
      const float tx0 = a.x + (b.x - a.x) * tx, bx0 = cc.x + (d.x - cc.x) * tx;
      const float tx1 = a.y + (b.y - a.y) * tx, bx1 = cc.y + (d.y - cc.y) * tx;
      out[q] = __floats2bfloat162_rn(tx0 + (bx0 - tx0) * ty, tx1 + (bx1 - tx1) * ty);
    }
    *reinterpret_cast<uint4*>(y + (((long long)img * oh + oy) * ow + ox) * yp + g * 8) =
        *reinterpret_cast<const uint4*>(out);
  }
}

static inline int grid_for2(long long work, int threads) {
  long long b = (work + threads - 1) / threads;
  const long long cap = (long long)device_sm_count() * 32;
  return (int)(b > cap ? cap : (b < 1 ? 1 : b));
}

}  // namespace eqxv

using namespace eqxv;

extern "C" int eqxv_resize_bilinear_nhwc_bf16_to_nchw_f32(const void* x, float* y, int32_t n, int32_t c,
                                                          int32_t h, int32_t w, int32_t oh, int32_t ow,
                                                          int32_t x_pitch, void* stream) {
  EQXV_CHECK_ARG(x && y && n > 0 && c > 0 && h > 0 && w > 0 && oh > 0 && ow > 0 && x_pitch >= c,
                 "resize: bad arguments");
  EQXV_CHECK_ARG(oh >= h && ow >= w, "resize: only upsampling matches jax.image.resize here");
  const long long total = (long long)n * c * oh * ow;
  EQXV_CUDA(launch_kernel(resize_to_nchw_kernel, dim3(grid_for2(total, 256)), dim3(256), (size_t)(0), (cudaStream_t)stream, 
      (const __nv_bfloat16*)x, y, n, c, h, w, oh, ow, x_pitch, (float)h / (float)oh, (float)w / (float)ow));
  EQXV_CUDA(cudaGetLastError());
  return EQXV_OK;
}

extern "C" int eqxv_resize_bilinear_nhwc_bf16(const void* x, void* y, int32_t n, int32_t c, int32_t h,
                                              int32_t w, int32_t oh, int32_t ow, int32_t x_pitch,
                                              int32_t y_pitch, void* stream) {
  EQXV_CHECK_ARG(x && y && n > 0 && c > 0 && h > 0 && w > 0 && oh > 0 && ow > 0, "resize: bad arguments");
  EQXV_CHECK_ARG(c % 8 == 0 && x_pitch % 8 == 0 && y_pitch % 8 == 0 && x_pitch >= c && y_pitch >= c,
                 "resize: channels/pitches must be multiples of 8");
  EQXV_CHECK_ARG(oh >= h && ow >= w, "resize: only upsampling matches jax.image.resize here");
  const long long total = (long long)n * oh * ow * (c / 8);
  EQXV_CUDA(launch_kernel(resize_nhwc_kernel, dim3(grid_for2(total, 256)), dim3(256), (size_t)(0), (cudaStream_t)stream, 
      (const __nv_bfloat16*)x, (__nv_bfloat16*)y, n, c, h, w, oh, ow, x_pitch, y_pitch, (float)h / (float)oh,
      (float)w / (float)ow));
  EQXV_CUDA(cudaGetLastError());
  return EQXV_OK;
}
