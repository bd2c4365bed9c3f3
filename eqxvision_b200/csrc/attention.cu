// Fused multi-head self-attention forward: softmax(q k^T * scale) v, head_dim 64, any token count.
// Reference: _VitAttention.__call__ vit.py:62-73 (scale applied after the product, softmax over the
// key axis, output transposed back to (tokens, heads*dim)).
//
// One CTA per (64-query tile, head, image); 4 warps x 16 query rows. Keys/values stream through
// shared memory in blocks of 64 with an online softmax (fp32 running max / sum in registers, warp
// shuffles across the 4 lanes that share a row), so S = q k^T never leaves the SM.
// v1 uses the legacy mma.sync tensor path (HMMA): attention is ~4 % of ViT-B/16 FLOPs; the tcgen05
// version is tracked in DESIGN.md.
#include <algorithm>
#include <cstdlib>
#include <cstring>

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
  griddep_wait();   // PDL: the predecessor kernel has completed (ptx.cuh)
  griddep_launch();
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


// ------------------------------------------------------------------------------------------------
// tcgen05 version (tokens <= 208): persistent, one CTA per SM, a flat sequence of 128-query tiles
// g = 0, 1, ... over the (image, head) pairs of this CTA.
//
//   S_g[128 x Tk] = Q_g K^T          tcgen05.mma, A = Q tile, B = K (both K-major SW128), fp32 in TMEM;
//                                    two S regions (g & 1) so that S_{g+2} is produced while g+1 is in softmax
//   P_g = exp2((S_g - rowmax) * scale*log2e)   8 softmax warps, TWO threads per query row, each holds half
//                                    of the row's scores in registers (one TMEM pass, no online rescaling:
//                                    the whole key axis fits), bf16 P -> SW128 smem
//   O_g[128 x 64] = P_g V            A = P (K-major), B = V read MN-major straight from the TMA tile
//   out = O_g / rowsum               fp32 row sums accumulated from the unrounded probabilities
//
// Software pipeline of the softmax warps: softmax(g) | drain O_{g-1} | write P_g: the exponentials
// (MUFU-bound) of tile g overlap the P.V product of tile g-1 and the Q.K^T of tile g+1; they only ever
// wait for work that was issued a whole softmax earlier.
// warp 0: one thread issues the TMA loads (Q/K/V of the next pair are prefetched into the second smem
// buffer) and all MMAs. Tk = tokens rounded up to 16; rows/keys beyond `tokens` are zero-filled by TMA
// (3-D map: column, token, image) and masked. Reference arithmetic: vit.py:62-73.
constexpr int kAtThreads = 288;

struct alignas(64) AttnParams {
  CUtensorMap tm;        // qkv as [images][tokens][3*heads*64]
  CUtensorMap tmO;       // out as [images][tokens][heads*64], box = 32 rows x 64 columns (ping-pong kernel)
  __nv_bfloat16* out;
  int tokens, tk, heads, pairs, ntile;
  int op_bytes;          // bytes per operand buffer (tk * 128 rounded up to 1024)
  float scale_log2;
  long long* ts;   // optional phase timestamps of CTA 0 (tools/attn_timeline.py); NULL in production
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// phase timestamps: slot = tile * 16 + event, events 0..7 MMA thread, 8..15 softmax warp 1 lane 0
#define AT_TS(ev, g)                                                                          \
  do {                                                                                        \
    if (p.ts != nullptr && blockIdx.x == 0 && (g) < 16) p.ts[(g) * 16 + (ev)] = clock64();    \
  } while (0)

constexpr uint32_t kAtS = 208;   // TMEM columns per S region; O lives at [2*kAtS, 2*kAtS + 64)

template <int NCH>  // 8-key chunks held per softmax thread (2 threads per row): 13 covers Tk <= 208
__global__ void __launch_bounds__(kAtThreads, 1) attention_tc_kernel(const __grid_constant__ AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t buf_bytes = 3u * (uint32_t)p.op_bytes;
  const uint32_t p_smem = base + 2 * buf_bytes;            // P: 4 K-blocks of [128 x 64] (64 KiB)
  uint8_t* p_g = gbase + 2 * buf_bytes;
  float* red_g = reinterpret_cast<float*>(gbase + 2 * buf_bytes + 65536);   // [2 halves][128] x {max, sum}
  const uint32_t bars = p_smem + 65536 + 2048;
  auto bar_load = [&](int b) { return bars + 8u * b; };
  auto bar_s = [&](int r) { return bars + 16u + 8u * r; };
  const uint32_t bar_p = bars + 32u, bar_o = bars + 40u;
  const uint32_t tmem_slot = bars + 48u;
  volatile uint32_t* tmem_slot_g = reinterpret_cast<volatile uint32_t*>(gbase + 2 * buf_bytes + 65536 + 2048 + 48);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.tm);
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_load(i), 1);
      mbar_init(bar_s(i), 1);
    }
    mbar_init(bar_p, 256);
    mbar_init(bar_o, 1);
    mbar_fence_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  griddep_wait();     // PDL: everything above overlapped the previous kernel's tail
  tc_fence_before();
  __syncthreads();
  griddep_launch();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_g;
  const int C = p.heads * 64;
  const int my_pairs = ((int)blockIdx.x < p.pairs) ? (p.pairs - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int G = my_pairs * p.ntile;   // tiles this CTA processes

  if (warp == 0) {
    if (elect_one()) {
      const uint64_t hi = (uint64_t)(umma_desc_sw128(0) >> 32) << 32;   // SBO 1024, SW128 (K-major and MN-major)
      const uint32_t lbo = 1u << 16;
      const uint32_t idesc_s = umma_idesc_bf16_m128((uint32_t)p.tk);
      const uint32_t idesc_o = umma_idesc_bf16_m128_bmn(64);
      const int ksteps = p.tk / 16;
      const uint32_t p_lo = ((p_smem & 0x3FFFF) >> 4) | lbo;
      auto issue_load = [&](int pi) {   // pi: index of the pair inside this CTA's sequence
        const int pair = (int)blockIdx.x + pi * (int)gridDim.x;
        const int img = pair / p.heads, head = pair - img * p.heads;
        const int b = pi & 1;
        const uint32_t dst = base + b * buf_bytes;
        mbar_expect_tx(bar_load(b), 3u * (uint32_t)(p.tk * 128));
        for (int o = 0; o < 3; ++o) {
          asm volatile(
              "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
              ::"r"(dst + o * p.op_bytes), "l"(reinterpret_cast<uint64_t>(&p.tm)), "r"(bar_load(b)),
                "r"(o * C + head * 64), "r"(0), "r"(img)
              : "memory");
        }
      };
      auto issue_qk = [&](int g) {      // S_g -> region g & 1
        const int pi = g / p.ntile, t = g - pi * p.ntile;
        if (t == 0) mbar_wait(bar_load(pi & 1), (uint32_t)((pi >> 1) & 1));
        const uint32_t q_s = base + (pi & 1) * buf_bytes;
        const uint32_t q_lo = (((q_s + t * 16384) & 0x3FFFF) >> 4) | lbo;
        const uint32_t k_lo = (((q_s + p.op_bytes) & 0x3FFFF) >> 4) | lbo;
        const uint32_t d = tmem_base + (uint32_t)(g & 1) * kAtS;
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(d, hi | (q_lo + 2 * k), hi | (k_lo + 2 * k), idesc_s, (uint32_t)k);
        umma_commit(bar_s(g & 1));
      };
      if (my_pairs > 0) issue_load(0);
      if (my_pairs > 1) issue_load(1);
      if (G > 0) issue_qk(0);
      if (G > 1) issue_qk(1);
      for (int g = 0; g < G; ++g) {
        const int pi = g / p.ntile, t = g - pi * p.ntile;
        AT_TS(0, g);
        mbar_wait(bar_p, (uint32_t)(g & 1));    // P_g written; S_g fully read; O_{g-1} drained
        tc_fence_after();
        AT_TS(1, g);
        const uint32_t v_s = base + (pi & 1) * buf_bytes + 2 * p.op_bytes;
        const uint32_t v_lo = ((v_s & 0x3FFFF) >> 4) | lbo;
        const uint32_t d = tmem_base + 2 * kAtS;
        for (int j = 0; j < ksteps; ++j)
          umma_bf16(d, hi | (p_lo + (uint32_t)((j >> 2) * 1024 + (j & 3) * 2)), hi | (v_lo + (uint32_t)(j * 128)),
                    idesc_o, (uint32_t)j);
        umma_commit(bar_o);
        AT_TS(2, g);
        if (t == p.ntile - 1 && pi + 2 < my_pairs) {
          mbar_wait(bar_o, (uint32_t)(g & 1));  // last reader of this pair's buffer has finished
          issue_load(pi + 2);
        }
        AT_TS(3, g);
        if (g + 2 < G) issue_qk(g + 2);         // region g & 1 is free again
        AT_TS(4, g);
      }
    }
    __syncwarp();
  } else {
    // ============================== softmax + output: warps 1..8 ==============================
    const int sw = warp - 1;                 // 0..7
    const int half = sw >> 2;                // which half of the key axis this thread owns
    const int quad = warp & 3;               // TMEM lane quadrant
    const int row = quad * 32 + lane;        // query row inside the tile
    // column split in units of 8 keys. Every thread processes NCH chunks starting at ch0; chunks past the
    // key axis are masked to -inf / never written, so the loops below are fully unrolled with the scores
    // in registers.
    const int nch_tot = p.tk / 8;
    const int ch0 = half ? (nch_tot + 1) / 2 : 0;
    const int ch_end = half ? nch_tot : (nch_tot + 1) / 2;   // first chunk NOT owned by this thread
    const int lim = min(p.tokens, ch_end * 8);               // keys >= lim are not this thread's (or do not exist)
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16);

    // O of tile g: TMEM -> registers (before the P write of the next tile is announced: the next P.V
    // overwrites O), registers -> global AFTER the announcement so that neither the proxy fence nor the
    // MMA issuer ever waits on global stores.
    auto store_o = [&](const float* o, int g, float inv) {
      const int pi = g / p.ntile, t = g - pi * p.ntile;
      const int pair = (int)blockIdx.x + pi * (int)gridDim.x;
      const int img = pair / p.heads, head = pair - img * p.heads;
      const int tok = t * 128 + row;
      if (tok < p.tokens) {
        __nv_bfloat16* dst = p.out + ((long long)img * p.tokens + tok) * C + head * 64 + half * 32;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint32_t w[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const __nv_bfloat162 h = __floats2bfloat162_rn(o[q * 8 + 2 * e] * inv, o[q * 8 + 2 * e + 1] * inv);
            w[e] = *reinterpret_cast<const uint32_t*>(&h);
          }
          *reinterpret_cast<uint4*>(dst + q * 8) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
    };
    const uint32_t oaddr = lane_addr + 2 * kAtS + (uint32_t)(half * 32);

    float inv_prev = 0.f;
    bool live_prev = false;
    for (int g = 0; g < G; ++g) {
      const int t = g % p.ntile;
      const bool live = t * 128 + quad * 32 < p.tokens;   // warp-uniform: any valid row in this warp's slab
      float inv = 0.f;
      uint32_t pk[NCH][4];
      const bool rec = threadIdx.x == 32;
      if (rec) AT_TS(8, g);
      mbar_wait(bar_s(g & 1), (uint32_t)((g >> 1) & 1));
      tc_fence_after();
      if (rec) AT_TS(9, g);
      if (live) {
        float s[NCH][8];
        const uint32_t taddr = lane_addr + (uint32_t)(g & 1) * kAtS + (uint32_t)(ch0 * 8);
#pragma unroll
        for (int c = 0; c < NCH; ++c) tmem_ld_x8(taddr + c * 8, s[c]);
        tmem_ld_wait();
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};   // independent chains (ILP)
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          if ((ch0 + c) * 8 + 8 > lim) {   // warp-uniform: only the chunk(s) straddling the end need masking
#pragma unroll
            for (int e = 0; e < 8; ++e)
              if ((ch0 + c) * 8 + e >= lim) s[c][e] = -INFINITY;
          }
#pragma unroll
          for (int e = 0; e < 8; ++e) m4[e & 3] = fmaxf(m4[e & 3], s[c][e]);
        }
        float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
        red_g[half * 128 + row] = mx;
        if (rec) AT_TS(10, g);
        named_bar_sync(1, 256);
        mx = fmaxf(mx, red_g[(half ^ 1) * 128 + row]);
        const float mb = mx * p.scale_log2;
        float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
#pragma unroll
          for (int e = 0; e < 8; e += 2) {
            const float p0 = ex2_approx(fmaf(s[c][e], p.scale_log2, -mb));       // ex2(-inf) = 0: masked keys
            const float p1 = ex2_approx(fmaf(s[c][e + 1], p.scale_log2, -mb));
            s4[(e >> 1) & 3] += p0 + p1;
            const __nv_bfloat162 h = __floats2bfloat162_rn(p0, p1);
            pk[c][e >> 1] = *reinterpret_cast<const uint32_t*>(&h);
          }
        }
        float sum = (s4[0] + s4[1]) + (s4[2] + s4[3]);
        red_g[256 + half * 128 + row] = sum;
        if (rec) AT_TS(11, g);
        named_bar_sync(2, 256);
        sum += red_g[256 + (half ^ 1) * 128 + row];
        inv = 1.f / sum;
      } else {
        named_bar_sync(1, 256);
        named_bar_sync(2, 256);
      }
      // O_{g-1}: its P.V was issued a whole softmax ago; once it is complete the P buffer is free as well.
      if (rec) AT_TS(12, g);
      float o[32];
      if (g > 0) {
        mbar_wait(bar_o, (uint32_t)((g - 1) & 1));
        tc_fence_after();
        if (live_prev) {
          tmem_ld_x16(oaddr, o);
          tmem_ld_x16(oaddr + 16, o + 16);
          tmem_ld_wait();
        }
      }
      if (rec) AT_TS(13, g);
      if (live) {
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          const int col = (ch0 + c) * 8;
          if (col < ch_end * 8) {   // chunks past this thread's range belong to the other half
            uint8_t* dst = p_g + (col >> 6) * 16384 + sw128_off((uint32_t)row, (uint32_t)((col & 63) >> 3));
            *reinterpret_cast<uint4*>(dst) = make_uint4(pk[c][0], pk[c][1], pk[c][2], pk[c][3]);
          }
        }
      }
      tc_fence_before();           // TMEM reads (S_g, O_{g-1}) ordered before the MMAs that overwrite them
      fence_proxy_async_smem();    // P visible to the tensor core (async proxy)
      mbar_arrive(bar_p);
      if (rec) AT_TS(14, g);
      if (g > 0 && live_prev) store_o(o, g - 1, inv_prev);
      inv_prev = inv;
      live_prev = live;
    }
    if (G > 0) {
      mbar_wait(bar_o, (uint32_t)((G - 1) & 1));
      tc_fence_after();
      if (live_prev) {
        float o[32];
        tmem_ld_x16(oaddr, o);
        tmem_ld_x16(oaddr + 16, o + 16);
        tmem_ld_wait();
        store_o(o, G - 1, inv_prev);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// Attention probabilities as an OUTPUT (vit.py:70 returned through _VitBlock(return_attention=True),
// vit.py:151-152, VisionTransformer.get_last_self_attention vit.py:275-292): fp32
// softmax(q k^T * scale) [images, heads, tokens, tokens]. Runs once per call on the last block only
// and is bound by its own fp32 store (tokens^2 * 4 B per head); CUDA-core dot products from smem.
// One CTA per (32-query tile, head, image), 256 threads: thread j owns key j (+256, ...), keeps that
// key's 64 dims in registers and sweeps the 32 query rows (broadcast smem reads); then one warp per
// row normalises and streams the row out.
constexpr int kPrQ = 32;
__global__ void __launch_bounds__(256) attention_probs_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                              float* __restrict__ probs, int tokens, int heads,
                                                              float scale_log2) {
  griddep_wait();
  griddep_launch();
  extern __shared__ __align__(16) uint8_t pr_smem[];
  const int tpad = (tokens + 3) & ~3;
  float* qs = reinterpret_cast<float*>(pr_smem);             // [32][64] fp32
  float* sc = qs + kPrQ * kHd;                               // [32][tpad] scores
  const int q0 = blockIdx.x * kPrQ, head = blockIdx.y, img = blockIdx.z;
  const long long ld = 3ll * heads * kHd;
  const __nv_bfloat16* base = qkv + (long long)img * tokens * ld + head * kHd;
  const __nv_bfloat16* kptr = base + (long long)heads * kHd;
  const int nq = min(kPrQ, tokens - q0);
  for (int i = threadIdx.x; i < kPrQ * kHd; i += 256) {
    const int r = i >> 6, d = i & 63;
    qs[i] = r < nq ? __bfloat162float(base[(long long)(q0 + r) * ld + d]) : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.x; j < tokens; j += 256) {
    float kf[kHd];
    const uint4* kr = reinterpret_cast<const uint4*>(kptr + (long long)j * ld);
#pragma unroll
    for (int v = 0; v < 8; ++v) {
      const uint4 u = __ldg(kr + v);
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        kf[v * 8 + 2 * e] = __uint_as_float(w[e] << 16);
        kf[v * 8 + 2 * e + 1] = __uint_as_float(w[e] & 0xffff0000u);
      }
    }
    for (int r = 0; r < nq; ++r) {
      const float4* q4 = reinterpret_cast<const float4*>(qs + r * kHd);
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
      for (int d = 0; d < 16; ++d) {
        const float4 q = q4[d];
        a0 = fmaf(q.x, kf[4 * d], a0);
        a1 = fmaf(q.y, kf[4 * d + 1], a1);
        a2 = fmaf(q.z, kf[4 * d + 2], a2);
        a3 = fmaf(q.w, kf[4 * d + 3], a3);
      }
      sc[r * tpad + j] = (a0 + a1) + (a2 + a3);
    }
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < nq; r += 8) {
    float* row = sc + r * tpad;
    float mx = -INFINITY;
    for (int j = lane; j < tokens; j += 32) mx = fmaxf(mx, row[j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    for (int j = lane; j < tokens; j += 32) {
      const float e = exp2f((row[j] - mx) * scale_log2);   // scale applied after the product (vit.py:69)
      row[j] = e;
      sum += e;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = 1.f / sum;
    float* dst = probs + (((long long)img * heads + head) * tokens + (q0 + r)) * tokens;
    for (int j = lane; j < tokens; j += 32) dst[j] = row[j] * inv;
  }
}

// ------------------------------------------------------------------------------------------------
// Ping-pong variant of the tcgen05 kernel (same operands, same MMAs, same shared-memory layout).
//
// In attention_tc_kernel all eight softmax warps work on the SAME tile in lock-step: every phase of a tile
// (TMEM read + row max, exponentials, P write + proxy fence, O drain + global store) is a latency chain with
// two warps per SM sub-partition and nothing to overlap it (ncu r01s12: issue slots 24 % busy, flat stall
// profile; 7000 cycles per 128-row tile of which 2400 are the MUFU-bound exponentials).
// Here the softmax warps form TWO groups of four (one thread per query row, the whole key axis in two passes
// over TMEM: row max, then exp / sum / bf16 P), group g & 1 owns tile g and S region g & 1. While one group
// is in its exponentials the other waits for its P.V product, drains O and stores it. No cross-thread
// exchange is needed any more (one thread sees the whole row): no named barriers, no reduction scratch.
// P never touches shared memory: the bf16 probabilities are written back into the group's own S region
// (tcgen05.st, 2 values per 32-bit column) and the P.V product reads its A operand from tensor memory, so the
// two groups only meet at the single O accumulator (the product of tile g waits for the drain of O_{g-1}).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kAtThreads, 1) attention_pp_kernel(const __grid_constant__ AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t buf_bytes = 3u * (uint32_t)p.op_bytes;
  const uint32_t p_smem = base + 2 * buf_bytes;            // P: 4 K-blocks of [128 x 64] (64 KiB)
  const uint32_t bars = p_smem + 65536 + 2048;
  auto bar_load = [&](int b) { return bars + 8u * b; };
  auto bar_s = [&](int r) { return bars + 16u + 8u * r; };
  auto bar_p = [&](int r) { return bars + 32u + 8u * r; };
  // one O-ready barrier per group: a parity wait may lag its barrier by at most one phase, and a group only
  // ever observes the products of its own tiles
  auto bar_o = [&](int r) { return bars + 48u + 8u * r; };
  const uint32_t bar_od = bars + 64u;
  const uint32_t tmem_slot = bars + 72u;
  volatile uint32_t* tmem_slot_g = reinterpret_cast<volatile uint32_t*>(gbase + 2 * buf_bytes + 65536 + 2048 + 72);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.tm);
    tma_prefetch_desc(&p.tmO);
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_load(i), 1);
      mbar_init(bar_s(i), 1);
      mbar_init(bar_p(i), 128);
    }
    mbar_init(bar_o(0), 1);
    mbar_init(bar_o(1), 1);
    mbar_init(bar_od, 128);
    mbar_fence_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  griddep_wait();     // PDL: everything above overlapped the previous kernel's tail
  tc_fence_before();
  __syncthreads();
  griddep_launch();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_g;
  const int C = p.heads * 64;
  const int my_pairs = ((int)blockIdx.x < p.pairs) ? (p.pairs - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int G = my_pairs * p.ntile;   // tiles this CTA processes

  if (warp == 0) {
    if (elect_one()) {
      const uint64_t hi = (uint64_t)(umma_desc_sw128(0) >> 32) << 32;
      const uint32_t lbo = 1u << 16;
      const uint32_t idesc_s = umma_idesc_bf16_m128((uint32_t)p.tk);
      const uint32_t idesc_o = umma_idesc_bf16_m128_bmn(64);
      const int ksteps = p.tk / 16;
      auto issue_load = [&](int pi) {
        const int pair = (int)blockIdx.x + pi * (int)gridDim.x;
        const int img = pair / p.heads, head = pair - img * p.heads;
        const int b = pi & 1;
        const uint32_t dst = base + b * buf_bytes;
        mbar_expect_tx(bar_load(b), 3u * (uint32_t)(p.tk * 128));
        for (int o = 0; o < 3; ++o) {
          asm volatile(
              "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
              ::"r"(dst + o * p.op_bytes), "l"(reinterpret_cast<uint64_t>(&p.tm)), "r"(bar_load(b)),
                "r"(o * C + head * 64), "r"(0), "r"(img)
              : "memory");
        }
      };
      auto issue_qk = [&](int g) {      // S_g -> region g & 1
        const int pi = g / p.ntile, t = g - pi * p.ntile;
        if (t == 0) mbar_wait(bar_load(pi & 1), (uint32_t)((pi >> 1) & 1));
        const uint32_t q_s = base + (pi & 1) * buf_bytes;
        const uint32_t q_lo = (((q_s + t * 16384) & 0x3FFFF) >> 4) | lbo;
        const uint32_t k_lo = (((q_s + p.op_bytes) & 0x3FFFF) >> 4) | lbo;
        const uint32_t d = tmem_base + (uint32_t)(g & 1) * kAtS;
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(d, hi | (q_lo + 2 * k), hi | (k_lo + 2 * k), idesc_s, (uint32_t)k);
        umma_commit(bar_s(g & 1));
      };
      if (my_pairs > 0) issue_load(0);
      if (my_pairs > 1) issue_load(1);
      if (G > 0) issue_qk(0);
      if (G > 1) issue_qk(1);
      for (int g = 0; g < G; ++g) {
        const int pi = g / p.ntile, t = g - pi * p.ntile;
        AT_TS(0, g);
        mbar_wait(bar_p(g & 1), (uint32_t)((g >> 1) & 1));      // P_g written, S_g fully read
        if (g > 0) mbar_wait(bar_od, (uint32_t)((g - 1) & 1));  // O_{g-1} has left TMEM
        tc_fence_after();
        AT_TS(1, g);
        const uint32_t v_s = base + (pi & 1) * buf_bytes + 2 * p.op_bytes;
        const uint32_t v_lo = ((v_s & 0x3FFFF) >> 4) | lbo;
        const uint32_t d = tmem_base + 2 * kAtS;
        const uint32_t p_tmem = tmem_base + (uint32_t)(g & 1) * kAtS;   // P_g aliases the S region of its group
        for (int j = 0; j < ksteps; ++j)
          umma_bf16_ts(d, p_tmem + (uint32_t)(j * 8), hi | (v_lo + (uint32_t)(j * 128)), idesc_o, (uint32_t)j);
        umma_commit(bar_o(g & 1));
        AT_TS(2, g);
        if (t == p.ntile - 1 && pi + 2 < my_pairs) {
          mbar_wait(bar_o(g & 1), (uint32_t)((g >> 1) & 1));  // last reader of this pair's buffer has finished
          issue_load(pi + 2);
        }
        if (g + 2 < G) issue_qk(g + 2);         // region g & 1 is free again
        AT_TS(4, g);
      }
    }
    __syncwarp();
  } else {
    // ============================== softmax groups: warps 1..4 and 5..8 ==============================
    const int grp = (warp - 1) >> 2;         // tiles g with (g & 1) == grp
    const int quad = warp & 3;               // TMEM lane quadrant
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16);
    const uint32_t s_addr = lane_addr + (uint32_t)grp * kAtS;
    const uint32_t o_addr = lane_addr + 2 * kAtS;
    const int nch = p.tk / 16;               // 16-column chunks of the key axis
    // O leaves through the shared memory the probabilities no longer need: every warp stages its 32 rows
    // (4 KiB, 128B-swizzled) and one lane issues a TMA store (rows past `tokens` are clipped by the tensor
    // map). Direct 16-byte global stores touched 32 different lines per warp instruction and cost ~1800
    // cycles per tile on the softmax warps.
    const uint32_t o_slab = p_smem + (uint32_t)(grp * 4 + quad) * 4096u;
    uint8_t* o_slab_g = gbase + 2 * buf_bytes + (uint32_t)(grp * 4 + quad) * 4096u;
    for (int g = grp; g < G; g += 2) {
      const int pi = g / p.ntile, t = g - pi * p.ntile;
      const bool live = t * 128 + quad * 32 < p.tokens;   // warp-uniform: any valid row in this warp's slab
      const bool rec = lane == 0 && (warp == 1 || warp == 5);
      if (rec) AT_TS(8, g);
      mbar_wait(bar_s(grp), (uint32_t)((g >> 1) & 1));
      tc_fence_after();
      if (rec) AT_TS(9, g);
      float mb = 0.f, inv = 0.f;
      if (live) {
        // ---- pass 1: row maximum ----
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        auto max16 = [&](int c, float* s) {
          if (c * 16 + 16 > p.tokens) {
#pragma unroll
            for (int e = 0; e < 16; ++e)
              if (c * 16 + e >= p.tokens) s[e] = -INFINITY;
          }
#pragma unroll
          for (int e = 0; e < 16; ++e) m4[e & 3] = fmaxf(m4[e & 3], s[e]);
        };
        // software pipeline over stages of 32 columns: the TMEM loads of stage st+1 are in flight while stage
        // st is reduced (tcgen05.wait::ld waits for ALL outstanding loads, so the wait follows the processing)
        float ra[16], rb[16];
        tmem_ld_x16(s_addr, ra);
        tmem_ld_wait();
        for (int c = 0; c < nch; c += 2) {
          if (c + 1 < nch) tmem_ld_x16(s_addr + (c + 1) * 16, rb);
          max16(c, ra);
          tmem_ld_wait();
          if (c + 1 < nch) {
            if (c + 2 < nch) tmem_ld_x16(s_addr + (c + 2) * 16, ra);
            max16(c + 1, rb);
            tmem_ld_wait();
          }
        }
        mb = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])) * p.scale_log2;
      }
      if (rec) AT_TS(10, g);
      if (live) {
        // ---- pass 2: exponentials, row sum, bf16 P into the swizzled A-operand layout ----
        float s4[4] = {0.f, 0.f, 0.f, 0.f};
        auto exp16 = [&](int c, const float* s) {
          uint32_t pk[8];
          const bool tail = c * 16 + 16 > p.tokens;     // warp-uniform: only the chunk straddling the end
#pragma unroll
          for (int e = 0; e < 16; e += 2) {
            float p0 = ex2_approx(fmaf(s[e], p.scale_log2, -mb));
            float p1 = ex2_approx(fmaf(s[e + 1], p.scale_log2, -mb));
            if (tail) {                                    // keys past the end (zero-filled K rows)
              if (c * 16 + e >= p.tokens) p0 = 0.f;
              if (c * 16 + e + 1 >= p.tokens) p1 = 0.f;
            }
            s4[(e >> 1) & 3] += p0 + p1;
            const __nv_bfloat162 h = __floats2bfloat162_rn(p0, p1);
            pk[e >> 1] = *reinterpret_cast<const uint32_t*>(&h);
          }
          // P chunk c (16 bf16 = 8 packed columns) overwrites columns [8c, 8c+8) of the S region: they belong
          // to S chunk c/2 <= c, which is already in registers
          tmem_st_x8(s_addr + (uint32_t)(c * 8), pk);
        };
        float ra[16], rb[16];
        tmem_ld_x16(s_addr, ra);
        tmem_ld_wait();
        for (int c = 0; c < nch; c += 2) {
          if (c + 1 < nch) tmem_ld_x16(s_addr + (c + 1) * 16, rb);
          exp16(c, ra);
          tmem_ld_wait();
          if (c + 1 < nch) {
            if (c + 2 < nch) tmem_ld_x16(s_addr + (c + 2) * 16, ra);
            exp16(c + 1, rb);
            tmem_ld_wait();
          }
        }
        inv = 1.f / ((s4[0] + s4[1]) + (s4[2] + s4[3]));
      }
      if (rec) AT_TS(12, g);
      tmem_st_wait();              // P_g is in tensor memory
      tc_fence_before();           // ... and ordered before the MMA that reads it
      mbar_arrive(bar_p(grp));
      if (rec) AT_TS(13, g);
      // ---- O_g: wait for the P.V product, drain, release the accumulator, store ----
      mbar_wait(bar_o(grp), (uint32_t)((g >> 1) & 1));
      tc_fence_after();
      float o[64];
      if (live) {
#pragma unroll
        for (int q = 0; q < 4; ++q) tmem_ld_x16(o_addr + q * 16, o + q * 16);
        tmem_ld_wait();
      }
      tc_fence_before();
      mbar_arrive(bar_od);
      if (rec) AT_TS(14, g);
      if (live) {
        const int pair = (int)blockIdx.x + pi * (int)gridDim.x;
        const int img = pair / p.heads, head = pair - img * p.heads;
        if (lane == 0) tma_store_wait_read<0>();   // this warp's previous store has read the slab
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          uint32_t w[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const __nv_bfloat162 h = __floats2bfloat162_rn(o[q * 8 + 2 * e] * inv, o[q * 8 + 2 * e + 1] * inv);
            w[e] = *reinterpret_cast<const uint32_t*>(&h);
          }
          *reinterpret_cast<uint4*>(o_slab_g + sw128_off((uint32_t)lane, (uint32_t)q)) = make_uint4(w[0], w[1], w[2], w[3]);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_3d(&p.tmO, o_slab, head * 64, t * 128 + quad * 32, img);
          tma_store_commit();
        }
      }
    }
    if (lane == 0) tma_store_wait_all();
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

static long long* g_attn_ts = nullptr;   // set by eqxv_debug_attention_timeline

static int launch_attention_tc(const void* qkv, void* out, int images, int tokens, int heads, float scale,
                               cudaStream_t stream) {
  AttnParams p;
  memset(&p, 0, sizeof(p));
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.tokens = tokens, p.heads = heads, p.pairs = images * heads;
  p.tk = ceil_div(tokens, 16) * 16;
  p.ntile = ceil_div(tokens, 128);
  p.op_bytes = ceil_div(p.tk * 128, 1024) * 1024;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.ts = g_attn_ts;
  TmapSpec m{};
  m.base = const_cast<void*>(qkv);
  m.dtype = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  m.rank = 3;
  const uint64_t ld = 3ull * heads * 64;
  m.dims[0] = ld, m.dims[1] = (uint64_t)tokens, m.dims[2] = (uint64_t)images;
  m.strides_bytes[0] = ld * 2, m.strides_bytes[1] = ld * 2 * (uint64_t)tokens;
  m.box[0] = 64, m.box[1] = (uint32_t)p.tk, m.box[2] = 1;
  m.estride[0] = m.estride[1] = m.estride[2] = 1;
  m.swizzle = CU_TENSOR_MAP_SWIZZLE_128B;
  int rc = encode_tmap(&p.tm, m);
  if (rc) return rc;
  TmapSpec mo{};
  mo.base = out;
  mo.dtype = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  mo.rank = 3;
  const uint64_t ldo = (uint64_t)heads * 64;
  mo.dims[0] = ldo, mo.dims[1] = (uint64_t)tokens, mo.dims[2] = (uint64_t)images;
  mo.strides_bytes[0] = ldo * 2, mo.strides_bytes[1] = ldo * 2 * (uint64_t)tokens;
  mo.box[0] = 64, mo.box[1] = 32, mo.box[2] = 1;
  mo.estride[0] = mo.estride[1] = mo.estride[2] = 1;
  mo.swizzle = CU_TENSOR_MAP_SWIZZLE_128B;
  rc = encode_tmap(&p.tmO, mo);
  if (rc) return rc;
  // smem: 2 x (Q,K,V) + P (64 KiB) + reductions (2 KiB) + barriers; the Q tile of the second M block may be
  // read up to row 255 of a tk-row buffer: the bytes behind it (K, V, P) are finite garbage feeding rows
  // that are never stored.
  const int smem = 2 * 3 * p.op_bytes + 65536 + 2048 + 128 + 1024;
  const int grid = std::min(p.pairs, device_sm_count());
  // two softmax groups ping-ponging on alternate tiles (default) or the lock-step kernel (EQXV_ATTN_PP=0)
  const char* pp = getenv("EQXV_ATTN_PP");
  if (pp == nullptr || atoi(pp) != 0) {
    EQXV_CUDA(launch_kernel(attention_pp_kernel, dim3(grid), dim3(kAtThreads), (size_t)(smem), stream, p));
  } else {
    EQXV_CUDA(launch_kernel(attention_tc_kernel<13>, dim3(grid), dim3(kAtThreads), (size_t)(smem), stream, p));
  }
  EQXV_CUDA(cudaGetLastError());
  return EQXV_OK;
}

int attention_init() {
  EQXV_CUDA(cudaFuncSetAttribute(attention_tc_kernel<13>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
  EQXV_CUDA(cudaFuncSetAttribute(attention_pp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
  EQXV_CUDA(cudaFuncSetAttribute(attention_probs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  return EQXV_OK;
}

}  // namespace eqxv

using namespace eqxv;

extern "C" int eqxv_attention_fwd_bf16(const void* qkv, void* out, float* attn_out, int32_t images,
                                       int32_t tokens, int32_t heads, int32_t head_dim, float scale,
                                       void* stream) {
  EQXV_CHECK_ARG(qkv && (out || attn_out) && images > 0 && tokens > 0 && heads > 0, "attention: bad arguments");
  if (head_dim != kHd) {
    set_error("attention: head_dim %d unsupported (only 64)", head_dim);
    return EQXV_ERR_UNSUPPORTED;
  }
  if (attn_out != nullptr) {
    const int tpad = (tokens + 3) & ~3;
    const size_t smem = (size_t)(kPrQ * kHd + kPrQ * tpad) * sizeof(float);
    EQXV_CHECK_ARG(smem <= 200 * 1024 && heads <= 65535 && images <= 65535 && ((uintptr_t)qkv & 15) == 0,
                   "attention: probability output supports up to ~1500 tokens");
    EQXV_CUDA(launch_kernel(attention_probs_kernel, dim3((unsigned)((tokens + kPrQ - 1) / kPrQ), (unsigned)heads,
                                                         (unsigned)images),
                            dim3(256), smem, (cudaStream_t)stream, (const __nv_bfloat16*)qkv, attn_out, tokens,
                            heads, scale * 1.4426950408889634f));
    if (out == nullptr) return EQXV_OK;   // probabilities only (return_attention=True discards the values)
  }
  // tcgen05 path: the double-buffered Q/K/V tiles + P fit in shared memory up to 208 keys (ViT @224: 197)
  if (tokens <= 208 && ((uintptr_t)qkv & 15) == 0 && ((uintptr_t)out & 15) == 0)
    return launch_attention_tc(qkv, out, images, tokens, heads, scale, (cudaStream_t)stream);
  EQXV_CHECK_ARG(heads <= 65535 && images <= 65535, "attention: grid too large");
  dim3 grid((unsigned)((tokens + kQT - 1) / kQT), (unsigned)heads, (unsigned)images);
  EQXV_CUDA(launch_kernel(attention_kernel, dim3(grid), dim3(128), (size_t)(0), (cudaStream_t)stream, 
      (const __nv_bfloat16*)qkv, (__nv_bfloat16*)out, tokens, heads, scale * 1.4426950408889634f));
  EQXV_CUDA(cudaGetLastError());
  return EQXV_OK;
}

// Debug aid (not on the product path): when `ts` is non-NULL the tcgen05 attention kernel records clock64()
// phase timestamps of CTA 0 into ts[16 tiles][16 events]; pass NULL to switch it off again.
extern "C" int eqxv_debug_attention_timeline(long long* ts) {
  g_attn_ts = ts;
  return EQXV_OK;
}
