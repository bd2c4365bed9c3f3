// Fused multi-head self-attention forward: softmax(q k^T * scale) v, head_dim 64, any token count.
// Reference: _VitAttention.__call__ vit.py:62-73 (scale applied after the product, softmax over the
// key axis, output transposed back to (tokens, heads*dim)).
//
// One CTA per (64-query tile, head, image); 4 warps x 16 query rows. Keys/values stream through
// shared memory in blocks of 64 with an online softmax (fp32 running max / sum in registers, warp
// shuffles across the 4 lanes that share a row), so S = q k^T never leaves the SM.
// v1 uses the legacy mma.sync tensor path (HMMA): attention is ~4 % of ViT-B/16 FLOPs; the tcgen05
// version is tracked in DESIGN.md.
#include "common.h"
#include "ptx.cuh"

namespace eqxv {

constexpr int kHd = 64;        // head dim
constexpr int kQT = 64;        // queries per CTA
constexpr int kKB = 64;        // keys per block
constexpr int kPitch = 72;     // smem row pitch in elements (144 B: conflict-free fragment loads)

__device__ __forceinline__ void mma_bf16_16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t* r, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

// copy `rows_valid` rows (64 bf16 each, global row stride ld) into a [64][kPitch] smem tile,
// zero-filling the rest; 128 threads, 16-byte vectors
__device__ __forceinline__ void load_tile(__nv_bfloat16* dst, const __nv_bfloat16* src, long long ld,
                                          int rows_valid) {
  for (int i = threadIdx.x; i < 64 * 8; i += 128) {
    const int r = i >> 3, v = i & 7;
    uint4 val = make_uint4(0u, 0u, 0u, 0u);
    if (r < rows_valid) val = __ldg(reinterpret_cast<const uint4*>(src + (long long)r * ld + v * 8));
    *reinterpret_cast<uint4*>(dst + r * kPitch + v * 8) = val;
  }
}

__global__ void __launch_bounds__(128) attention_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                        __nv_bfloat16* __restrict__ out, int tokens,
                                                        int heads, float scale_log2) {
  __shared__ __align__(16) __nv_bfloat16 Qs[kQT * kPitch];
  __shared__ __align__(16) __nv_bfloat16 Ks[kKB * kPitch];
  __shared__ __align__(16) __nv_bfloat16 Vs[kKB * kPitch];

  const int q0 = blockIdx.x * kQT;
  const int head = blockIdx.y;
  const int img = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const long long ld = 3ll * heads * kHd;
  const __nv_bfloat16* base = qkv + (long long)img * tokens * ld + head * kHd;
  const __nv_bfloat16* qptr = base;
  const __nv_bfloat16* kptr = base + (long long)heads * kHd;
  const __nv_bfloat16* vptr = base + 2ll * heads * kHd;

  load_tile(Qs, qptr + (long long)q0 * ld, ld, min(kQT, tokens - q0));
  __syncthreads();

  // Q fragments for this warp's 16 rows: 4 k-steps x 4 registers
  uint32_t qf[4][4];
  {
    const __nv_bfloat16* qw = Qs + (warp * 16) * kPitch;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      qf[ks][0] = *reinterpret_cast<const uint32_t*>(qw + g * kPitch + ks * 16 + 2 * t);
      qf[ks][1] = *reinterpret_cast<const uint32_t*>(qw + (g + 8) * kPitch + ks * 16 + 2 * t);
      qf[ks][2] = *reinterpret_cast<const uint32_t*>(qw + g * kPitch + ks * 16 + 2 * t + 8);
      qf[ks][3] = *reinterpret_cast<const uint32_t*>(qw + (g + 8) * kPitch + ks * 16 + 2 * t + 8);
    }
  }

  float o[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY};
  float l_run[2] = {0.f, 0.f};

  const int nkb = (tokens + kKB - 1) / kKB;
  for (int kb = 0; kb < nkb; ++kb) {
    const int valid = min(kKB, tokens - kb * kKB);
    __syncthreads();  // previous block's fragments are consumed
    load_tile(Ks, kptr + (long long)kb * kKB * ld, ld, valid);
    load_tile(Vs, vptr + (long long)kb * kKB * ld, ld, valid);
    __syncthreads();

    // ---- S = Q K^T (16 x 64 per warp) ----
    float s[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int j = 0; j < 4; ++j) s[nt][j] = 0.f;
      if (nt * 8 < valid) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const __nv_bfloat16* kr = Ks + (nt * 8 + g) * kPitch + ks * 16 + 2 * t;
          const uint32_t b0 = *reinterpret_cast<const uint32_t*>(kr);
          const uint32_t b1 = *reinterpret_cast<const uint32_t*>(kr + 8);
          mma_bf16_16816(s[nt], qf[ks], b0, b1);
        }
      }
    }
    // ---- mask, running max ----
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = nt * 8 + 2 * t + (j & 1);
        if (col >= valid) s[nt][j] = -INFINITY;
        mx[j >> 1] = fmaxf(mx[j >> 1], s[nt][j]);
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
    }
    float m_new[2], corr[2], rs[2] = {0.f, 0.f};
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      m_new[r] = fmaxf(m_run[r], mx[r]);
      corr[r] = exp2f((m_run[r] - m_new[r]) * scale_log2);
      m_run[r] = m_new[r];
    }
    // ---- P = exp(scale * (S - max)) as bf16 A fragments ----
    uint32_t pf[4][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      float pv[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        pv[j] = exp2f((s[nt][j] - m_new[j >> 1]) * scale_log2);
        rs[j >> 1] += pv[j];
      }
      pf[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16(pv[0], pv[1]);
      pf[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(pv[2], pv[3]);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      rs[r] += __shfl_xor_sync(0xffffffffu, rs[r], 1);
      rs[r] += __shfl_xor_sync(0xffffffffu, rs[r], 2);
      l_run[r] = l_run[r] * corr[r] + rs[r];
    }
#pragma unroll
    for (int dn = 0; dn < 8; ++dn) {
      o[dn][0] *= corr[0];
      o[dn][1] *= corr[0];
      o[dn][2] *= corr[1];
      o[dn][3] *= corr[1];
    }
    // ---- O += P V ----
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      if (kk * 16 < valid) {
#pragma unroll
        for (int dp = 0; dp < 4; ++dp) {  // pairs of 8-wide d tiles
          uint32_t vb[4];
          const int krow = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
          const int dcol = dp * 16 + (lane >> 4) * 8;
          ldmatrix_x4_trans(vb, smem_u32(Vs + krow * kPitch + dcol));
          mma_bf16_16816(o[dp * 2], pf[kk], vb[0], vb[1]);
          mma_bf16_16816(o[dp * 2 + 1], pf[kk], vb[2], vb[3]);
        }
      }
    }
  }

  // ---- normalise and store: out[(img*tokens + row), head*64 + d] ----
  const long long ldo = (long long)heads * kHd;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int row = q0 + warp * 16 + g + r * 8;
    if (row < tokens) {
      const float inv = 1.f / l_run[r];
      __nv_bfloat16* dst = out + ((long long)img * tokens + row) * ldo + head * kHd + 2 * t;
#pragma unroll
      for (int dn = 0; dn < 8; ++dn) {
        *reinterpret_cast<uint32_t*>(dst + dn * 8) =
            pack_bf16(o[dn][r * 2] * inv, o[dn][r * 2 + 1] * inv);
      }
    }
  }
}

int attention_init() { return EQXV_OK; }

}  // namespace eqxv

using namespace eqxv;

extern "C" int eqxv_attention_fwd_bf16(const void* qkv, void* out, float* attn_out, int32_t images,
                                       int32_t tokens, int32_t heads, int32_t head_dim, float scale,
                                       void* stream) {
  EQXV_CHECK_ARG(qkv && out && images > 0 && tokens > 0 && heads > 0, "attention: bad arguments");
  if (head_dim != kHd) {
    set_error("attention: head_dim %d unsupported (only 64)", head_dim);
    return EQXV_ERR_UNSUPPORTED;
  }
  if (attn_out != nullptr) {
    set_error("attention: returning the probability matrix is not implemented yet");
    return EQXV_ERR_UNSUPPORTED;
  }
  EQXV_CHECK_ARG(heads <= 65535 && images <= 65535, "attention: grid too large");
  dim3 grid((unsigned)((tokens + kQT - 1) / kQT), (unsigned)heads, (unsigned)images);
  attention_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)qkv, (__nv_bfloat16*)out, tokens, heads, scale * 1.4426950408889634f);
  EQXV_CUDA(cudaGetLastError());
  return EQXV_OK;
}
