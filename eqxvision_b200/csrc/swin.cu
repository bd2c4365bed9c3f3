// Swin glue (K16): shifted-window attention and patch merging.
// Reference: models/classification/swin.py:90-255 (_shifted_window_attention) and :23-43
// (_patch_merging_pad). The cyclic roll, the window partition / reverse and the shift mask are pure
// index arithmetic on the channels-last token matrix: nothing is permuted in memory.
#include "common.h"
#include "ptx.cuh"

namespace eqxv {

// One CTA = one (window, image, head group): blockIdx.z picks a contiguous slice of the heads (one head per CTA when the
// grid would otherwise be small: the last Swin stage has ONE window per image and 24 heads - looping over them inside
// 64 CTAs left most of the 148 SMs idle: 660 us per launch at batch 64). qkv rows are in SPATIAL order
// (row = (img*H + y)*W + x), columns ordered (3, heads, head_dim) as produced by reshape(..,3,heads,d)
// (swin.py:166-171). A window token (i) of window (wr, wc) sits at rolled position
// (wr*ws + i/ws, wc*ws + i%ws), i.e. at source pixel ((r + shift) % H, (c + shift) % W) (jnp.roll by
// -shift, swin.py:122-123); the output goes back to the same source pixel (reverse roll, :249-250).
template <int HD>
__global__ void __launch_bounds__(128) window_attention_kernel(
    const __nv_bfloat16* __restrict__ qkv, const float* __restrict__ bias, __nv_bfloat16* __restrict__ out,
    int H, int W, int heads, int ws, int shift_h, int shift_w, float scale) {
  griddep_wait();   // PDL: the predecessor kernel has completed (ptx.cuh)
  griddep_launch();
  extern __shared__ float sm[];
  const int T = ws * ws;                 // tokens per window (<= 64)
  float* q = sm;                         // [T][HD+1]
  float* k = q + T * (HD + 1);
  float* v = k + T * (HD + 1);
  float* S = v + T * (HD + 1);           // [T][T+1]
  int* src = reinterpret_cast<int*>(S + T * (T + 1));   // [T] source pixel index
  int* lab = src + T;                    // [T] mask region label
  const int nwc = W / ws;
  const int wr = blockIdx.x / nwc, wc = blockIdx.x % nwc;
  const int img = blockIdx.y;
  const int C = heads * HD;
  const long long ld = 3ll * C;
  const int tid = threadIdx.x;
  for (int i = tid; i < T; i += blockDim.x) {
    const int r = wr * ws + i / ws, c = wc * ws + i % ws;      // rolled coordinates
    const int sy = (r + shift_h) % H, sx = (c + shift_w) % W;  // source pixel
    src[i] = sy * W + sx;
    // region labels of the shift mask (swin.py:185-229): 3 bands per axis
    const int lh = (shift_h == 0) ? 0 : (r < H - ws ? 0 : (r < H - shift_h ? 1 : 2));
    const int lw = (shift_w == 0) ? 0 : (c < W - ws ? 0 : (c < W - shift_w ? 1 : 2));
    lab[i] = lh * 3 + lw;
  }
  __syncthreads();
  const long long img_row0 = (long long)img * H * W;
  const int hpb = (heads + gridDim.z - 1) / gridDim.z;       // heads per CTA
  const int h_begin = blockIdx.z * hpb, h_end = min(heads, h_begin + hpb);
  for (int h = h_begin; h < h_end; ++h) {
    // ---- load q, k, v of this head (bf16 -> fp32 smem) ----
    for (int e = tid; e < T * (HD / 2); e += blockDim.x) {
      const int i = e / (HD / 2), d2 = e % (HD / 2);
      const __nv_bfloat16* row = qkv + (img_row0 + src[i]) * ld + h * HD + d2 * 2;
      const float2 fq = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(row));
      const float2 fk = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(row + C));
      const float2 fv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(row + 2 * C));
      q[i * (HD + 1) + d2 * 2] = fq.x * scale;       // q * d^-1/2 before the product (swin.py:180)
      q[i * (HD + 1) + d2 * 2 + 1] = fq.y * scale;
      k[i * (HD + 1) + d2 * 2] = fk.x;
      k[i * (HD + 1) + d2 * 2 + 1] = fk.y;
      v[i * (HD + 1) + d2 * 2] = fv.x;
      v[i * (HD + 1) + d2 * 2 + 1] = fv.y;
    }
    __syncthreads();
    // ---- S = q k^T + relative position bias + shift mask ----
    const float* bh = bias + (long long)h * T * T;
    for (int e = tid; e < T * T; e += blockDim.x) {
      const int i = e / T, j = e % T;
      float acc = 0.f;
#pragma unroll
      for (int d = 0; d < HD; ++d) acc = fmaf(q[i * (HD + 1) + d], k[j * (HD + 1) + d], acc);
      acc += __ldg(bh + e);
      if (lab[i] != lab[j]) acc += -100.f;
      S[i * (T + 1) + j] = acc;
    }
    __syncthreads();
    // ---- row softmax: one warp per row ----
    for (int i = tid >> 5; i < T; i += blockDim.x >> 5) {
      const int lane = tid & 31;
      float m = -INFINITY;
      for (int j = lane; j < T; j += 32) m = fmaxf(m, S[i * (T + 1) + j]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      float sum = 0.f;
      for (int j = lane; j < T; j += 32) {
        const float p = __expf(S[i * (T + 1) + j] - m);
        S[i * (T + 1) + j] = p;
        sum += p;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      const float inv = 1.f / sum;
      for (int j = lane; j < T; j += 32) S[i * (T + 1) + j] *= inv;
    }
    __syncthreads();
    // ---- O = P V, written back to the source pixel, column block of this head ----
    for (int e = tid; e < T * (HD / 2); e += blockDim.x) {
      const int i = e / (HD / 2), d2 = e % (HD / 2);
      float a0 = 0.f, a1 = 0.f;
      for (int j = 0; j < T; ++j) {
        const float p = S[i * (T + 1) + j];
        a0 = fmaf(p, v[j * (HD + 1) + d2 * 2], a0);
        a1 = fmaf(p, v[j * (HD + 1) + d2 * 2 + 1], a1);
      }
      *reinterpret_cast<__nv_bfloat162*>(out + (img_row0 + src[i]) * C + h * HD + d2 * 2) =
          __floats2bfloat162_rn(a0, a1);
    }
    __syncthreads();
  }
}

// Swin-V2 cosine attention as the REFERENCE computes it (swin.py:158-166): q / ||q|| and k / ||k|| with the L2 norm taken
// over axis 0 of the (num_windows, heads, tokens, d) arrays, i.e. over the WINDOWS of one image for every (head, window
// token, channel) - not over the channel axis as torchvision does. q is then multiplied by
// exp(min(logit_scale, log 100)) per head (swin.py:164-166; `scale_q`, evaluated on the host). In place on the q and k
// column ranges of the spatial-order qkv matrix; one CTA = one (window token, image), one thread = 8 channels.
__global__ void __launch_bounds__(256) swin_v2_qk_normalize_kernel(__nv_bfloat16* __restrict__ qkv,
                                                                   const float* __restrict__ scale_q, int H, int W, int C,
                                                                   int head_dim, int ws, int shift_h, int shift_w) {
  griddep_wait();
  griddep_launch();
  const int ty = blockIdx.x / ws, tx = blockIdx.x % ws;
  const int img = blockIdx.y;
  const int nwr = H / ws, nwc = W / ws;
  const long long ld = 3ll * C;
  __nv_bfloat16* base = qkv + (long long)img * H * W * ld;
  for (int v = threadIdx.x; v < 2 * C / 8; v += blockDim.x) {      // q columns [0,C), k columns [C,2C)
    float ss[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int wr = 0; wr < nwr; ++wr)
      for (int wc = 0; wc < nwc; ++wc) {
        const int sy = (wr * ws + ty + shift_h) % H, sx = (wc * ws + tx + shift_w) % W;
        const uint4 raw = *reinterpret_cast<const uint4*>(base + ((long long)sy * W + sx) * ld + v * 8);
        const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 f = __bfloat1622float2(h2[i]);
          ss[2 * i] = fmaf(f.x, f.x, ss[2 * i]);
          ss[2 * i + 1] = fmaf(f.y, f.y, ss[2 * i + 1]);
        }
      }
    const float sq = (v * 8 < C) ? __ldg(scale_q + (v * 8) / head_dim) : 1.f;
    float inv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) inv[i] = sq / sqrtf(ss[i]);
    for (int wr = 0; wr < nwr; ++wr)
      for (int wc = 0; wc < nwc; ++wc) {
        const int sy = (wr * ws + ty + shift_h) % H, sx = (wc * ws + tx + shift_w) % W;
        uint4* ptr = reinterpret_cast<uint4*>(base + ((long long)sy * W + sx) * ld + v * 8);
        uint4 raw = *ptr;
        __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&raw);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 f = __bfloat1622float2(h2[i]);
          h2[i] = __floats2bfloat162_rn(f.x * inv[2 * i], f.y * inv[2 * i + 1]);
        }
        *ptr = raw;
      }
  }
}

// out[n, y2, x2, k*C + c] = x[n, 2*y2 + dy_k, 2*x2 + dx_k, c],  (dy,dx)_k = (0,0),(1,0),(0,1),(1,1)
// (x0,x1,x2,x3 of swin.py:26-31); H, W even.
__global__ void patch_merge_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int n, int h,
                                   int w, int c, int xp, int yp) {
  griddep_wait();   // PDL: the predecessor kernel has completed (ptx.cuh)
  griddep_launch();
  const int groups = c / 8;
  const int h2 = h / 2, w2 = w / 2;
  const long long total = (long long)n * h2 * w2 * 4 * groups;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % groups);
    long long t = i / groups;
    const int kq = (int)(t % 4);
    t /= 4;
    const int x2 = (int)(t % w2);
    t /= w2;
    const int y2 = (int)(t % h2);
    const int img = (int)(t / h2);
    const int dy = kq & 1, dx = kq >> 1;
    const uint4 val = __ldg(reinterpret_cast<const uint4*>(
        x + (((long long)img * h + 2 * y2 + dy) * w + 2 * x2 + dx) * xp + g * 8));
    *reinterpret_cast<uint4*>(y + (((long long)img * h2 + y2) * w2 + x2) * yp + kq * c + g * 8) = val;
  }
}

}  // namespace eqxv

using namespace eqxv;

extern "C" int eqxv_window_attention_bf16(const void* qkv, const float* bias, void* out, int32_t n, int32_t h,
                                          int32_t w, int32_t heads, int32_t head_dim, int32_t window,
                                          int32_t shift_h, int32_t shift_w, float scale, void* stream) {
  EQXV_CHECK_ARG(qkv && bias && out && n > 0 && h > 0 && w > 0 && heads > 0, "window_attention: bad arguments");
  EQXV_CHECK_ARG(window >= 1 && window <= 8 && h % window == 0 && w % window == 0,
                 "window_attention: the map (%dx%d) must be a multiple of the window (%d <= 8)", h, w, window);
  EQXV_CHECK_ARG(shift_h >= 0 && shift_h < window && shift_w >= 0 && shift_w < window && n <= 65535,
                 "window_attention: bad shift");
  if (head_dim != 32) {
    set_error("window_attention: head_dim %d unsupported (only 32)", head_dim);
    return EQXV_ERR_UNSUPPORTED;
  }
  const int T = window * window;
  const size_t smem = (size_t)(3 * T * 33 + T * (T + 1)) * 4 + 2 * T * 4;
  static bool attr = false;
  if (!attr) {
    EQXV_CUDA(cudaFuncSetAttribute(window_attention_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    attr = true;
  }
  // split the heads over blockIdx.z until the grid fills the machine a few times over
  const long long wn = (long long)(h / window) * (w / window) * n;
  int zsplit = 1;
  while (zsplit < heads && wn * zsplit < 8ll * device_sm_count()) ++zsplit;
  while (heads % zsplit != 0) ++zsplit;
  dim3 grid((unsigned)((h / window) * (w / window)), (unsigned)n, (unsigned)zsplit);
  EQXV_CUDA(launch_kernel(window_attention_kernel<32>, dim3(grid), dim3(128), (size_t)(smem), (cudaStream_t)stream, 
      (const __nv_bfloat16*)qkv, bias, (__nv_bfloat16*)out, h, w, heads, window, shift_h, shift_w, scale));
  EQXV_CUDA(cudaGetLastError());
  return EQXV_OK;
}

extern "C" int eqxv_swin_v2_qk_normalize_bf16(void* qkv, const float* scale_q, int32_t n, int32_t h, int32_t w,
                                              int32_t heads, int32_t head_dim, int32_t window, int32_t shift_h,
                                              int32_t shift_w, void* stream) {
  EQXV_CHECK_ARG(qkv && scale_q && n > 0 && h > 0 && w > 0 && heads > 0 && head_dim > 0 && head_dim % 8 == 0,
                 "swin_v2_qk_normalize: bad arguments");
  EQXV_CHECK_ARG(window >= 1 && h % window == 0 && w % window == 0 && shift_h >= 0 && shift_h < window && shift_w >= 0 &&
                     shift_w < window && n <= 65535,
                 "swin_v2_qk_normalize: the map (%dx%d) must be a multiple of the window (%d)", h, w, window);
  dim3 grid((unsigned)(window * window), (unsigned)n);
  EQXV_CUDA(launch_kernel(swin_v2_qk_normalize_kernel, grid, dim3(256), (size_t)0, (cudaStream_t)stream,
                          (__nv_bfloat16*)qkv, scale_q, h, w, heads * head_dim, head_dim, window, shift_h, shift_w));
  EQXV_CUDA(cudaGetLastError());
  return EQXV_OK;
}

extern "C" int eqxv_patch_merge_bf16(const void* x, void* y, int32_t n, int32_t h, int32_t w, int32_t c,
                                     int32_t x_pitch, int32_t y_pitch, void* stream) {
  EQXV_CHECK_ARG(x && y && n > 0 && h > 0 && w > 0 && c > 0, "patch_merge: bad arguments");
  EQXV_CHECK_ARG(h % 2 == 0 && w % 2 == 0, "patch_merge: odd feature maps need padding (unsupported)");
  EQXV_CHECK_ARG(c % 8 == 0 && x_pitch % 8 == 0 && y_pitch % 8 == 0 && x_pitch >= c && y_pitch >= 4 * c,
                 "patch_merge: channels/pitches must be multiples of 8");
  const long long total = (long long)n * (h / 2) * (w / 2) * 4 * (c / 8);
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)device_sm_count() * 32;
  if (blocks > cap) blocks = cap;
  EQXV_CUDA(launch_kernel(patch_merge_kernel, dim3((int)blocks), dim3(256), (size_t)(0), (cudaStream_t)stream, (const __nv_bfloat16*)x, (__nv_bfloat16*)y, n, h,
                                                                    w, c, x_pitch, y_pitch));
  EQXV_CUDA(cudaGetLastError());
  return EQXV_OK;
}
