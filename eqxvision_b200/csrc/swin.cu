// Swin glue (K16): shifted-window attention and patch merging.
// Reference: models/classification/swin.py:90-255 (_shifted_window_attention) and :23-43
// (_patch_merging_pad). The cyclic roll, the window partition / reverse and the shift mask are pure
// index arithmetic on the channels-last token matrix: nothing is permuted in memory.
#include <algorithm>
#include <cstdlib>

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

// ---------------------------------------------------------------------------------------------------------------------
// Tensor-core window attention (head_dim 32, windows of <= 64 tokens): tcgen05.mma for q k^T and p v.
// One work item = TWO windows x one head = one 128-row MMA tile: rows 0..T-1 are the tokens of window A, rows 64..64+T-1
// those of window B (rows T..63 of each half are padding). Every operand is a K-major, 128-byte-swizzled shared-memory
// tile written by the CTA's own threads (the window's tokens are a gather over the rolled map: 7 row segments per
// window, wrapped at the border - no TMA box describes them):
//   Q, K   [128 rows x 64]  : 32 real channels + 32 zero columns        S = Q K^T        M=128, N=128, K=32
//   P      2 x [128 x 64]   : keys 0..63 (window A) | 64..127 (window B) O = P V          M=128, N=32,  K=128
//   V^T    2 x [32 x 64]    : V transposed on the way into shared memory (K-major B operand)
// P of a row is non-zero only in its OWN window's key block, so the cross-window quarter of S is never read and the
// other block of P stays at the zeros it was initialised with. Softmax: one thread per row (scale, relative-position
// bias, -100 shift mask, max, exp, sum), probabilities unnormalised in bf16, 1/sum applied to O.
// Items run back to back inside a CTA (no software pipeline); two CTAs per SM overlap one item's loads / softmax with
// the other's MMAs. S lives in TMEM columns [0,128), O in [128,160).
// Reference: swin.py:117-253 (same arithmetic as window_attention_kernel above, which stays as the fallback).
struct WinAttnParams {
  const __nv_bfloat16* qkv;
  const float* bias;       // [heads][T][T]
  __nv_bfloat16* out;
  int n, H, W, heads, ws, shift_h, shift_w;
  float scale;
  int nwin, total_w, items;   // windows per image, windows in the batch, work items = ceil(total_w / 2) * heads
};

__global__ void __launch_bounds__(128) window_attention_tc_kernel(const WinAttnParams p) {
  extern __shared__ uint8_t wa_raw[];
  const uint32_t raw = smem_u32(wa_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gb = wa_raw + (base - raw);
  constexpr uint32_t kQ = 0, kK = 16384, kP = 32768, kV = 65536, kBias = 73728, kMeta = 90112, kBars = 91136;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int T = p.ws * p.ws;
  const int C = p.heads * 32;
  const long long ld = 3ll * C;
  const uint32_t bar_s = base + kBars, bar_o = base + kBars + 8;
  const uint32_t tmem_slot = base + kBars + 16;
  float* s_bias = reinterpret_cast<float*>(gb + kBias);
  int* s_src = reinterpret_cast<int*>(gb + kMeta);          // [128] absolute source row of each tile row (-1: padding)
  int* s_lab = s_src + 128;                                 // [128] shift-mask region label

  // zero every operand tile once: padding rows / columns and the cross-window halves of P are never written again
  for (uint32_t i = tid; i < 73728u / 16u; i += 128) reinterpret_cast<uint4*>(gb)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (tid == 0) {
    mbar_init(bar_s, 1);
    mbar_init(bar_o, 1);
    mbar_fence_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 256u);
    tmem_relinquish();
  }
  griddep_wait();   // PDL: the predecessor kernel has completed (ptx.cuh)
  tc_fence_before();
  __syncthreads();
  griddep_launch();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gb + kBars + 16);
  const uint32_t tmem_s = tmem_base, tmem_o = tmem_base + 128u;
  const uint32_t lane_sel = (uint32_t)(warp * 32) << 16;   // this warp's TMEM lane quadrant
  const uint32_t idesc_s = umma_idesc_bf16_m128(128u), idesc_o = umma_idesc_bf16_m128(32u);
  const uint64_t dq = umma_desc_sw128(base + kQ), dk = umma_desc_sw128(base + kK);
  const uint32_t desc_hi = (uint32_t)(dq >> 32);
  const uint32_t p_lo = (((base + kP) & 0x3FFFF) >> 4) | (1u << 16), v_lo = (((base + kV) & 0x3FFFF) >> 4) | (1u << 16);

  const int nwc = p.W / p.ws;
  const int half = tid >> 6, tok = tid & 63;    // window (A / B) and token of this thread's row
  const int wpairs = (p.total_w + 1) >> 1;
  uint32_t phase = 0;
  int bias_head = -1;
  for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
    const int head = item / wpairs, wp = item - head * wpairs;
    // ---- this row's token: window -> rolled coordinates -> source pixel, mask label ----
    const int gw = 2 * wp + half;
    int src = -1, lab = 0;
    if (tok < T && gw < p.total_w) {
      const int img = gw / p.nwin, win = gw - img * p.nwin;
      const int r = (win / nwc) * p.ws + tok / p.ws, c = (win % nwc) * p.ws + tok % p.ws;   // rolled coordinates
      const int sy = (r + p.shift_h) % p.H, sx = (c + p.shift_w) % p.W;
      src = (img * p.H + sy) * p.W + sx;
      const int lh = (p.shift_h == 0) ? 0 : (r < p.H - p.ws ? 0 : (r < p.H - p.shift_h ? 1 : 2));
      const int lw = (p.shift_w == 0) ? 0 : (c < p.W - p.ws ? 0 : (c < p.W - p.shift_w ? 1 : 2));
      lab = lh * 3 + lw;
    }
    s_src[tid] = src;
    s_lab[tid] = lab;
    if (head != bias_head) {   // relative position bias of this head (9.6 KB for 7x7 windows)
      for (int i = tid; i < T * T; i += 128) s_bias[i] = __ldg(p.bias + (long long)head * T * T + i);
      bias_head = head;
    }
    // ---- gather q, k (rows, K-major) and v (transposed) of this row's token ----
    if (src >= 0) {
      const __nv_bfloat16* row = p.qkv + (long long)src * ld + head * 32;
      uint4 q4[4], k4[4], v4[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        q4[j] = __ldg(reinterpret_cast<const uint4*>(row) + j);
        k4[j] = __ldg(reinterpret_cast<const uint4*>(row + C) + j);
        v4[j] = __ldg(reinterpret_cast<const uint4*>(row + 2 * C) + j);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        *reinterpret_cast<uint4*>(gb + kQ + sw128_off(tid, j)) = q4[j];
        *reinterpret_cast<uint4*>(gb + kK + sw128_off(tid, j)) = k4[j];
      }
      // V^T[d][key]: key = tid -> K block `half`, column tok
      uint8_t* vt = gb + kV + half * 4096;
      const uint32_t cchunk = (uint32_t)tok >> 3, cin = ((uint32_t)tok & 7u) * 2u;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint16_t* e = reinterpret_cast<const uint16_t*>(&v4[j]);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const uint32_t d = (uint32_t)(j * 8 + q);
          *reinterpret_cast<uint16_t*>(vt + d * 128u + ((cchunk ^ (d & 7u)) << 4) + cin) = e[q];
        }
      }
    }
    fence_proxy_async_smem();
    __syncthreads();
    // ---- S = Q K^T (K = 32: two 16-deep steps) ----
    if (tid == 0) {
      tc_fence_after();
      umma_bf16(tmem_s, dq, dk, idesc_s, 0u);
      umma_bf16(tmem_s, dq + 2, dk + 2, idesc_s, 1u);
      umma_commit(bar_s);
    }
    mbar_wait(bar_s, phase);
    tc_fence_after();
    // ---- softmax of this row over its own window's keys ----
    float inv = 0.f;
    {
      float sv[64];
      const uint32_t srow = tmem_s + lane_sel + (uint32_t)(half * 64);
#pragma unroll
      for (int j = 0; j < 4; ++j) tmem_ld_x16(srow + j * 16, sv + j * 16);
      tmem_ld_wait();
      uint32_t pk[32];
      if (src >= 0) {
        const float* brow = s_bias + tok * T;
        const int* labs = s_lab + half * 64;
        float m = -INFINITY;
#pragma unroll
        for (int j = 0; j < 64; ++j) {
          float x = -INFINITY;
          if (j < T) {
            x = fmaf(sv[j], p.scale, brow[j]);
            if (labs[j] != lab) x += -100.f;
          }
          sv[j] = x;
          m = fmaxf(m, x);
        }
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < 64; j += 2) {
          const float e0 = __expf(sv[j] - m), e1 = __expf(sv[j + 1] - m);   // exp(-inf) = 0 beyond T
          const __nv_bfloat162 h2 = __floats2bfloat162_rn(e0, e1);
          // the row sum uses the ROUNDED probabilities the MMA will multiply with
          sum += __low2float(h2) + __high2float(h2);
          pk[j >> 1] = *reinterpret_cast<const uint32_t*>(&h2);
        }
        inv = 1.f / sum;
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) pk[j] = 0u;
      }
      uint8_t* prow = gb + kP + half * 16384;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<uint4*>(prow + sw128_off(tid, j)) = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
    }
    tc_fence_before();
    fence_proxy_async_smem();
    __syncthreads();
    // ---- O = P V (K = 128 keys: two 64-deep blocks) ----
    if (tid == 0) {
      tc_fence_after();
      umma_bf16_kblock64_nc(tmem_o, p_lo, v_lo, desc_hi, desc_hi, idesc_o, 0u);
      umma_bf16_kblock64_nc(tmem_o, p_lo + (16384u >> 4), v_lo + (4096u >> 4), desc_hi, desc_hi, idesc_o, 1u);
      umma_commit(bar_o);
    }
    mbar_wait(bar_o, phase);
    tc_fence_after();
    {
      float o[32];
      tmem_ld_x32(tmem_o + lane_sel, o);
      tmem_ld_wait();
      if (src >= 0) {
        __nv_bfloat16* orow = p.out + (long long)src * C + head * 32;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t w4[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const __nv_bfloat162 h2 = __floats2bfloat162_rn(o[j * 8 + 2 * q] * inv, o[j * 8 + 2 * q + 1] * inv);
            w4[q] = *reinterpret_cast<const uint32_t*>(&h2);
          }
          reinterpret_cast<uint4*>(orow)[j] = make_uint4(w4[0], w4[1], w4[2], w4[3]);
        }
      }
    }
    tc_fence_before();
    __syncthreads();   // TMEM and the operand tiles are free for the next item
    phase ^= 1u;
  }
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256u);
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
  {
    // tensor-core path (tcgen05): EQXV_WATTN_TC=0 falls back to the CUDA-core kernel below (A/B, parity reference)
    static const bool use_tc = !(getenv("EQXV_WATTN_TC") && getenv("EQXV_WATTN_TC")[0] == '0');
    const long long rows = (long long)n * h * w;
    if (use_tc && rows < (1ll << 31) && (((uintptr_t)qkv | (uintptr_t)out) & 15) == 0) {
      WinAttnParams wp{};
      wp.qkv = (const __nv_bfloat16*)qkv, wp.bias = bias, wp.out = (__nv_bfloat16*)out;
      wp.n = n, wp.H = h, wp.W = w, wp.heads = heads, wp.ws = window, wp.shift_h = shift_h, wp.shift_w = shift_w;
      wp.scale = scale;
      wp.nwin = (h / window) * (w / window);
      wp.total_w = n * wp.nwin;
      wp.items = ((wp.total_w + 1) / 2) * heads;
      constexpr int kSmem = 91136 + 64 + 1024;
      static bool attr_tc = false;
      if (!attr_tc) {
        EQXV_CUDA(cudaFuncSetAttribute(window_attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
        attr_tc = true;
      }
      const int grid_tc = std::min(wp.items, 2 * device_sm_count());
      EQXV_CUDA(launch_kernel(window_attention_tc_kernel, dim3((unsigned)grid_tc), dim3(128), (size_t)kSmem,
                              (cudaStream_t)stream, wp));
      EQXV_CUDA(cudaGetLastError());
      return EQXV_OK;
    }
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
