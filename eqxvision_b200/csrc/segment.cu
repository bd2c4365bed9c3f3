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

// NHWC bf16 -> NCHW fp32 (the model output handed back to the caller). The output is 64x larger than the
// input for the x8 DeepLab upsample, so the stores are what matters: grid = (column blocks, output rows,
// image x channel planes), one thread = 4 consecutive output columns = one 16-byte store, no integer
// division anywhere (the one-element-per-thread version with a linear index spent ~150 instructions per
// output on 64-bit div/mod: 140 us per 88 MB DeepLabV3 output against 15 us of store traffic).
__global__ void __launch_bounds__(128) resize_to_nchw_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ y,
                                                             int c, int h, int w, int oh, int ow, int xp, float sh,
                                                             float sw) {
  griddep_wait();   // PDL: the predecessor kernel has completed (ptx.cuh)
  griddep_launch();
  const int ox0 = (blockIdx.x * 128 + threadIdx.x) * 4;
  if (ox0 >= ow) return;
  const int oy = blockIdx.y;
  const int plane = blockIdx.z;                 // img * c + ch
  const int img = plane / c, ch = plane - img * c;
  int y0, y1;
  float ty;
  src_index(oy, sh, h, y0, y1, ty);
  const __nv_bfloat16* r0 = x + ((long long)img * h + y0) * w * xp + ch;
  const __nv_bfloat16* r1 = x + ((long long)img * h + y1) * w * xp + ch;
  float out[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int x0, x1;
    float tx;
    src_index(min(ox0 + j, ow - 1), sw, w, x0, x1, tx);
    const float v00 = __bfloat162float(r0[(long long)x0 * xp]), v01 = __bfloat162float(r0[(long long)x1 * xp]);
    const float v10 = __bfloat162float(r1[(long long)x0 * xp]), v11 = __bfloat162float(r1[(long long)x1 * xp]);
    const float top = v00 + (v01 - v00) * tx;
    const float bot = v10 + (v11 - v10) * tx;
    out[j] = top + (bot - top) * ty;
  }
  float* dst = y + ((long long)plane * oh + oy) * ow + ox0;
  if (ox0 + 3 < ow && (ow & 3) == 0) {
    *reinterpret_cast<float4*>(dst) = make_float4(out[0], out[1], out[2], out[3]);
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (ox0 + j < ow) dst[j] = out[j];
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
  EQXV_CHECK_ARG(oh <= 65535 && (long long)n * c <= 65535, "resize: output too tall / too many planes");
  EQXV_CHECK_ARG(((uintptr_t)y & 15) == 0, "resize: output must be 16-byte aligned");
  EQXV_CUDA(launch_kernel(resize_to_nchw_kernel, dim3((unsigned)((ow + 511) / 512), (unsigned)oh, (unsigned)(n * c)),
                          dim3(128), (size_t)0, (cudaStream_t)stream, (const __nv_bfloat16*)x, y, c, h, w, oh, ow,
                          x_pitch, (float)h / (float)oh, (float)w / (float)ow));
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
