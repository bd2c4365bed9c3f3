// Implicit-GEMM convolution / linear layer on the 5th-generation tensor cores (tcgen05), sm_100a.
//
//   D[128 pixels x block_n channels] (fp32, TMEM)  +=  A[128 x 64] (bf16, smem)  x  B[block_n x 64]^T
//
// * persistent kernel, one CTA per SM, static round-robin tile schedule (n-tiles fastest so CTAs
//   that run together share the same activation tile through L2)
// * warp 0 (one thread): TMA producer.  A tiles are 4-D boxes (64 ch, tw, th, tn) of the NHWC
//   activation: the (kh,kw) taps are the same box shifted by (r*dil-pad, s*dil-pad); out-of-bounds
//   zero fill is the convolution padding, the box traversal stride is the convolution stride.
//   B tiles are [block_n x 64] slabs of the K-major packed filter.
// * warp 1 (one thread): tcgen05.mma issuer, 4 x (128 x block_n x 16) per 64-wide K block,
//   accumulators double-buffered in TMEM so the epilogue of tile i overlaps the mainloop of i+1.
// * warps 2..5: epilogue. tcgen05.ld -> +bias(folded BN shift) -> (+residual) -> activation ->
//   bf16 -> 128B-swizzled smem -> TMA store; the residual tile is prefetched by TMA.
//
// Replaces the XLA:CPU conv/dot thunks behind equinox.nn.Conv2d / Linear + BatchNorm(inference) +
// jax.nn activation as composed in resnet.py:144-162, conv_norm_activation.py:61-85, vit.py:64,74,
// mlps.py:61-65 (reference paths; see include/eqxv_b200.h).
#include <algorithm>
#include <cstdlib>

#include "common.h"
#include "ptx.cuh"
#include "epilogue.cuh"

namespace eqxv {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kEpiWarpsMax = 8;                 // two epilogue warps per TMEM lane quadrant
constexpr int kThreads = 64 + 32 * kEpiWarpsMax;  // warp 0 producer, warp 1 MMA issuer, warps 2..9 epilogue
constexpr int kABytes = kBlockM * kBlockK * 2;  // 16 KiB
constexpr int kStageBuf = 16384;                // one epilogue staging buffer (128 rows x 128 B)
constexpr int kMaxSmem = 232448;                // 227 KiB

struct alignas(64) IgemmParams {
  CUtensorMap tmA, tmB, tmC, tmR;
  const float* bias;
  int tiles_w, tiles_h, tiles_n, n_tiles, num_tiles;
  int tw, th, tn;
  int block_n, acc_stride, tmem_cols;
  int kh, kw, dil_h, dil_w, pad_h, pad_w, mul_h, mul_w;
  int in_h, in_w;   // extent of A dims 2 / 1 (tap skipping)
  int kchunks;      // ceil(cin / 64)
  int a_tail;       // EQXV_FLAG_K_TAIL_SHIFT: channel coordinate of the LAST K chunk's A box (cin - 64), else -1
  int cin_pack;     // K offset between consecutive taps in B
  int cout;
  int act, res_after_act, has_res;
  int stages;
  int b_resident;   // halo kernel: the whole packed filter stays in the B ring (loaded once per CTA)
  int grouped;      // 1: block-diagonal grouped conv, the A channel offset follows the n-tile (block_n == 64)
  int epi_sub;      // epilogue warps per TMEM lane quadrant (1 or 2): blockDim = 64 + 128 * epi_sub
  int epi_obufs;    // output staging slabs per epilogue warp (eight-warp epilogue: 1, or 2 where shared memory allows)
  // shared-memory carve-up (byte offsets from the 1024-aligned base)
  int off_out, off_res, off_bias, off_bars;
  // LayerNorm folded into the GEMMs on either side of it (flat GEMMs only, kLN template parameter):
  //   kLN == 2 (producer, y = a w^T + bias + residual): besides y, (sum, sum of squares) of every stored row over each
  //             64-column chunk go to ln_stats[row][chunk]
  //   kLN == 1 (consumer of LayerNorm(x) with gamma folded into w and beta into bias): y = act(rstd[row] * (acc -
  //             mean[row] * wsum[col]) + bias[col]); mean / rstd from the producer's ln_stats
  float2* ln_stats;
  const float* ln_wsum;
  int ln_slots, ln_rows, off_wsum;
  float ln_inv_d, ln_eps;
  // Direct epilogue (flat bf16 GEMMs with few K blocks and <= 32 output channels): every thread stores its row's 16-byte
  // vectors straight to global memory and reads the residual the same way - no staging slab, no proxy fence, no TMA
  // store / load per 32-row slab (see launch_igemm for where it wins and where it loses).
  int direct_out;
  __nv_bfloat16* y_ptr;
  const __nv_bfloat16* res_ptr;
  int y_pitch, res_pitch;
  // SqueezeExcitation gate applied to the A operand (kGate kernels, flat GEMMs): A[row, k] *= gate[row / gate_rpi, k]
  const __nv_bfloat16* gate;
  int gate_pitch, gate_rpi, gate_cols;
  // Narrow-K A operand (kEpi16 kernels, flat GEMMs with ONE K block and k < 64): a TMA box that reaches past the tensor's
  // inner extent is served at a fraction of the normal rate (measured: [1.6 M x 24] -> 144 columns runs at 1.3 TB/s,
  // the same rows padded to 64 columns at 5.9 TB/s), so warp 0 gathers the rows with 16-byte cp.async instead.
  const __nv_bfloat16* ga_ptr;
  const __nv_bfloat16* gb_ptr;   // the filter [cout, gb_pitch]: gathered ONCE per CTA into a resident slab at off_bres
  int ga_on, ga_pitch, ga_rows, ga_k16, gb_pitch, off_bres;
  // first-layer kernel with the 3x3 / stride 2 / pad 1 max-pool in its epilogue (stem_pool_epilogue)
  __nv_bfloat16* pool_y;
  int pool_h, pool_w, pool_pitch;
  // first-layer (halo) kernel only
  int h_stride, h_planes, h_px, h_rows, h_plane_pitch, h_stage_bytes, h_ksteps, h_off_b;
  int h_tma;        // first-layer kernel, pixel-pair layout: the halo tile is ONE dense TMA box per tile (no gather warps)
  long long* dbg;   // eqxv_debug_stem_timeline: clock64 stamps of CTA 0 (tools/stem_timeline.py), normally null
  int h_stride_y;   // vertical stride (== h_stride except for the pixel-pair layout: horizontal 1, vertical 2)
  const void* h_src;  // padded NHWC8 image [n, in_h, in_w, 8]
};

#define IG_STAMP(it, ev)                                                                        \
  do {                                                                                          \
    if (p.dbg != nullptr && blockIdx.x == 0 && (it) < 24) p.dbg[(it) * 16 + (ev)] = clock64(); \
  } while (0)

struct TileCoord {
  int ncol0, w0, h0, n0;
};

// kPair: `tile` counts pairs of m-tiles; the CTA's rank inside its cluster picks the m-tile. An m-tile
// past the end (odd m-tile count) decodes to an out-of-range image index: TMA zero-fills its loads and
// clips its stores.
template <bool kPair = false>
__device__ __forceinline__ TileCoord decode_tile(const IgemmParams& p, int tile) {
  TileCoord t;
  const int nt = tile % p.n_tiles;
  int m = tile / p.n_tiles;
  if constexpr (kPair) m = 2 * m + (int)(blockIdx.x & 1u);
  t.ncol0 = nt * p.block_n;
  t.w0 = (m % p.tiles_w) * p.tw;
  m /= p.tiles_w;
  t.h0 = (m % p.tiles_h) * p.th;
  t.n0 = (m / p.tiles_h) * p.tn;
  return t;
}

// A tap whose whole box lies in the zero padding contributes nothing (dilated ASPP convs,
// deeplabv3.py:43-53 with d=24/36 on a 64x64 map): producer and MMA issuer skip it identically.

// ============================== epilogue, four warps (epi_sub == 1) ==============================
// One warp per TMEM lane quadrant, tile-major loop, double-buffered staging slab per warp. Used where the
// K loop hides the epilogue (first layer, halo kernel, deep-K residual GEMMs): it is measurably leaner per
// chunk than the generic two-warps-per-quadrant loop below (first layer 183 vs 205 us on B200).
template <bool kOutF32, int kAct, int kRes, bool kPair = false, int kLN = 0>
__device__ __forceinline__ void epilogue_warps_x4(const IgemmParams& p, const uint32_t base, uint8_t* gbase,
                                               const uint32_t tmem_base, const int warp, const int lane) {
  const int S = p.stages;
  const int t_first = kPair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;   // persistent schedule of this CTA
  const int t_stride = kPair ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const uint32_t bars = base + p.off_bars;
  auto tfull_bar = [&](int a) { return bars + 8u * (2 * S + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (2 * S + 2 + a); };
  {
    // ============================== epilogue (warps 2..5) ==============================
    // Every warp owns the 32 accumulator rows of its TMEM lane quadrant as an independent slab:
    // its own staging buffers, its own TMA stores / residual loads (issued by an elected lane) and
    // its own mbarriers. There is no CTA-wide barrier on this path, so one warp waiting on a TMA
    // store or a residual tile never stalls the other three.
    constexpr int CH = kOutF32 ? 32 : 64;  // columns per staged chunk (128 B per row)
    const int quad = warp & 3;             // TMEM lane quadrant this warp may access
    const int cpt = (p.block_n + CH - 1) / CH;
    const float* s_bias = reinterpret_cast<const float*>(gbase + p.off_bias);
    constexpr bool has_res = kRes != 0;
    // slab origin inside the (tn, th, tw) tile: rows are ordered n, h, w (w fastest)
    const int so = quad * 32;
    const int w_off = so % p.tw, h_off = (so / p.tw) % p.th, n_off = so / (p.tw * p.th);
    constexpr uint32_t kSlab = 32 * 128;  // 4 KiB per warp per buffer
    const uint32_t out_u32 = base + p.off_out + quad * kSlab;   // + buf * kStageBuf
    const uint32_t res_u32 = base + p.off_res + quad * kSlab;
    uint8_t* out_g = gbase + p.off_out + quad * kSlab;
    const uint8_t* res_g = gbase + p.off_res + quad * kSlab;
    auto rbar = [&](uint32_t b) { return bars + 8u * (48u + (uint32_t)quad * 2u + b); };

    auto issue_res = [&](uint32_t gg) {  // called by ONE lane
      const int ti = gg / cpt, c = gg - ti * cpt;
      const long long tile = (long long)t_first + (long long)ti * t_stride;
      if (tile >= p.num_tiles) return;
      const TileCoord t = decode_tile<kPair>(p, (int)tile);
      const uint32_t b = gg & 1u;
      mbar_expect_tx(rbar(b), kSlab);
      tma_load_4d(res_u32 + b * kStageBuf, &p.tmR, rbar(b), t.ncol0 + c * CH, t.w0 + w_off, t.h0 + h_off,
                  t.n0 + n_off);
    };

    uint32_t g = 0;
    const bool tma_res = has_res && !(kLN == 0 && !kOutF32 && p.direct_out);
    if (tma_res && lane == 0) {
      issue_res(0);
      issue_res(1);
    }
    __syncwarp();
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = t_first; tile < p.num_tiles; tile += t_stride) {
      const TileCoord t = decode_tile<kPair>(p, tile);
      // LayerNorm folding (flat GEMMs: the tile is 128 consecutive rows, this thread owns row t.w0 + so + lane)
      // (the phantom m-tile of an odd pair decodes to n0 = 1, w0 = 0: its rows must not be mistaken for rows 0..127)
      const int ln_row = (t.n0 == 0 && t.h0 == 0) ? t.w0 + so + lane : p.ln_rows;
      float ln_rstd = 1.f, ln_nmr = 0.f;
      if constexpr (kLN == 1) {
        // statistics of this row, written by the producing GEMM's epilogue: loaded before the accumulator wait
        // (all loads of the row are issued back to back - one L2 latency, not one per slot: with a runtime-bound loop
        // the twelve dependent-looking loads of a ViT-B row cost ~8000 cycles at the head of every tile's epilogue)
        float sum = 0.f, sq = 0.f;
        if (ln_row < p.ln_rows) {
          const float2* st = p.ln_stats + (long long)ln_row * p.ln_slots;
          if ((p.ln_slots & 1) == 0 && p.ln_slots <= 32) {
            const float4* st4 = reinterpret_cast<const float4*>(st);   // rows of an even slot count are 16-byte aligned
            const int n4 = p.ln_slots >> 1;
            float4 a[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = i < n4 ? __ldg(st4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              sum += a[i].x + a[i].z;
              sq += a[i].y + a[i].w;
            }
          } else {
            for (int i = 0; i < p.ln_slots; ++i) {
              const float2 a = __ldg(st + i);
              sum += a.x;
              sq += a.y;
            }
          }
        }
        const float mean = sum * p.ln_inv_d;
        ln_rstd = rsqrtf(fmaxf(sq * p.ln_inv_d - mean * mean, 0.f) + p.ln_eps);
        ln_nmr = -mean * ln_rstd;
      }
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t t_acc = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * p.acc_stride);

      for (int c = 0; c < cpt; ++c, ++g) {
        const uint32_t buf = g & 1u;
        const uint32_t rphase = (g >> 1) & 1u;
        const int ncols = min(CH, p.block_n - c * CH);
        float v[CH];
#pragma unroll
        for (int j = 0; j < CH / 16; ++j) {
          if (j * 16 < ncols) {
            tmem_ld_x16(t_acc + (uint32_t)(c * CH + j * 16), &v[j * 16]);
          } else {
#pragma unroll
            for (int q = 0; q < 16; ++q) v[j * 16 + q] = 0.f;
          }
        }
        tmem_ld_wait();
        if (c == cpt - 1) {
          // all TMEM reads of this accumulator are done: hand it back to the MMA issuer
          tc_fence_before();
          if constexpr (kPair) {
            mbar_arrive_leader(tempty_bar(acc));   // the leader's issuer waits for both CTAs' epilogues
          } else {
            mbar_arrive(tempty_bar(acc));
          }
        }
        const float* bias_c = s_bias + t.ncol0 + c * CH;
        uint8_t* out_row = out_g + buf * kStageBuf;
        if constexpr (!kOutF32 && kLN == 0) {
          if (p.direct_out) {   // CTA-uniform
            const int col0 = t.ncol0 + c * CH;
            if (ln_row < p.ln_rows) {
              __nv_bfloat16* orow = p.y_ptr + (long long)ln_row * p.y_pitch + col0;
              const __nv_bfloat16* rrow = has_res ? p.res_ptr + (long long)ln_row * p.res_pitch + col0 : nullptr;
              uint4 rv[8];
#pragma unroll
              for (int j = 0; j < 8; ++j)
                rv[j] = (has_res && col0 + j * 8 < p.cout) ? __ldg(reinterpret_cast<const uint4*>(rrow + j * 8))
                                                           : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
              for (int j = 0; j < 8; ++j)
                if (col0 + j * 8 < p.cout)
                  *reinterpret_cast<uint4*>(orow + j * 8) = epilogue8<kAct, kRes>(&v[j * 8], bias_c + j * 8, rv[j]);
            }
            continue;
          }
        }
        if constexpr (!kOutF32) {
          uint4 packed[8];
          if (has_res) mbar_wait(rbar(buf), rphase);
          const uint8_t* res_row = res_g + buf * kStageBuf;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            uint4 rv = make_uint4(0u, 0u, 0u, 0u);
            if constexpr (has_res) rv = *reinterpret_cast<const uint4*>(res_row + sw128_off(lane, j));
            if constexpr (kLN == 1) {
              const float* wsum_c = reinterpret_cast<const float*>(gbase + p.off_wsum) + t.ncol0 + c * CH;
              packed[j] = epilogue8_ln<kAct>(&v[j * 8], bias_c + j * 8, wsum_c + j * 8, ln_rstd, ln_nmr);
            } else {
              packed[j] = epilogue8<kAct, kRes>(&v[j * 8], bias_c + j * 8, rv);
            }
          }
          if constexpr (kLN == 2) {
            // row statistics of what is STORED (the bf16-rounded values the LayerNorm would have read back)
            uint64_t sum2 = 0ull, sq2 = 0ull;   // packed pairs: (even columns, odd columns)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const uint32_t wv[4] = {packed[j].x, packed[j].y, packed[j].z, packed[j].w};
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const uint64_t f = bf2_to_f2(wv[q]);
                sum2 = f2_add(sum2, f);
                sq2 = f2_fma(f, f, sq2);
              }
            }
            float s_lo, s_hi, q_lo, q_hi;
            f2_unpack(sum2, s_lo, s_hi);
            f2_unpack(sq2, q_lo, q_hi);
            const float sum = s_lo + s_hi, sq = q_lo + q_hi;
            const int col0 = t.ncol0 + c * CH;
            if (ln_row < p.ln_rows && col0 < p.cout)
              p.ln_stats[(long long)ln_row * p.ln_slots + (col0 >> 6)] = make_float2(sum, sq);
          }
          if (lane == 0) tma_store_wait_read<1>();  // this warp's store that last used out[buf] is done
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<uint4*>(out_row + sw128_off(lane, j)) = packed[j];
        } else {
#pragma unroll
          for (int q = 0; q < CH; ++q) v[q] = apply_act<kAct>(v[q] + bias_c[q]);
          if (lane == 0) tma_store_wait_read<1>();
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(out_row + sw128_off(lane, j)) =
                make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_4d(&p.tmC, out_u32 + buf * kStageBuf, t.ncol0 + c * CH, t.w0 + w_off, t.h0 + h_off,
                       t.n0 + n_off);
          tma_store_commit();
          if (tma_res) issue_res(g + 2);
        }
        __syncwarp();
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
    if (lane == 0) tma_store_wait_all();
    __syncwarp();
  }
}


// ============================== epilogue (warps 2..) ==============================
// Shared by the generic implicit-GEMM kernel, the CTA-pair kernel, the halo kernel and the first-layer
// kernel. EIGHT warps (kEpiWarps): two per TMEM lane quadrant (a warp may only touch the 32 lanes
// 32*(warp%4)..), which take alternate 64-column chunks of the quadrant's 32 accumulator rows.
// With one warp per quadrant (= one per SM sub-partition) every chunk was a serial latency chain
// (tcgen05.ld -> residual wait -> math -> st.shared -> proxy fence -> TMA store) with nothing to overlap
// it: a 128 x 256 tile took ~8000 cycles of epilogue whatever the layer (ResNet c3 layers: 5.1 us per
// tile at K = 256, profiles/r01_layer_roofline_v3.txt), i.e. the epilogue, not HBM or the tensor
// pipe, bounded every shallow-K layer. Two warps per sub-partition hide each other's latencies.
// Every warp owns its slab privately: its own staging buffer, its own TMA stores / residual loads
// (issued by an elected lane) and its own mbarriers; there is no CTA-wide barrier on this path.
template <bool kOutF32, int kAct, int kRes, bool kPair = false, int kLN = 0>
__device__ __forceinline__ void epilogue_warps(const IgemmParams& p, const uint32_t base, uint8_t* gbase,
                                               const uint32_t tmem_base, const int warp, const int lane) {
  if (p.epi_sub == 1 || kLN != 0) {   // CTA-uniform (LayerNorm folding exists in the four-warp epilogue only)
    epilogue_warps_x4<kOutF32, kAct, kRes, kPair, kLN>(p, base, gbase, tmem_base, warp, lane);
    return;
  }
  const int S = p.stages;
  const int t_first = kPair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;   // persistent schedule of this CTA
  const int t_stride = kPair ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const uint32_t bars = base + p.off_bars;
  auto tfull_bar = [&](int a) { return bars + 8u * (2 * S + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (2 * S + 2 + a); };
  constexpr int CH = kOutF32 ? 32 : 64;  // columns per staged chunk (128 B per row)
  const int quad = warp & 3;             // TMEM lane quadrant this warp may access
  const int nsub = p.epi_sub;            // warps per quadrant (2 here)
  const int sub = ((warp - 2) >> 2) % nsub;
  const int wid = quad * nsub + sub;     // private slab index
  const int cpt = (p.block_n + CH - 1) / CH;
  const float* s_bias = reinterpret_cast<const float*>(gbase + p.off_bias);
  constexpr bool has_res = kRes != 0;
  // slab origin inside the (tn, th, tw) tile: rows are ordered n, h, w (w fastest)
  const int so = quad * 32;
  const int w_off = so % p.tw, h_off = (so / p.tw) % p.th, n_off = so / (p.tw * p.th);
  constexpr uint32_t kSlab = 32 * 128;  // 4 KiB: 32 rows x 128 B
  const uint32_t obufs = p.epi_obufs > 0 ? (uint32_t)p.epi_obufs : 2u / (uint32_t)nsub;   // staging slabs of this warp
  const uint32_t out_u32 = base + p.off_out + (uint32_t)wid * obufs * kSlab;
  uint8_t* out_g = gbase + p.off_out + (uint32_t)wid * obufs * kSlab;
  const uint32_t res_u32 = base + p.off_res + (uint32_t)wid * 2u * kSlab;   // residual ring: 2 slabs per warp
  const uint8_t* res_g = gbase + p.off_res + (uint32_t)wid * 2u * kSlab;
  auto rbar = [&](uint32_t b) { return bars + 8u * (48u + (uint32_t)wid * 2u + b); };

  // this warp's chunk sequence: local index l -> global chunk g = nsub*l + sub of the CTA's tile sequence
  auto issue_res = [&](uint32_t l) {  // called by ONE lane
    const int g = nsub * (int)l + sub;
    const int ti = g / cpt, c = g - ti * cpt;
    const long long tile = (long long)t_first + (long long)ti * t_stride;
    if (tile >= p.num_tiles) return;
    const TileCoord t = decode_tile<kPair>(p, (int)tile);
    const uint32_t b = l & 1u;
    mbar_expect_tx(rbar(b), kSlab);
    tma_load_4d(res_u32 + b * kSlab, &p.tmR, rbar(b), t.ncol0 + c * CH, t.w0 + w_off, t.h0 + h_off, t.n0 + n_off);
  };
  auto release_acc = [&](int ti) {   // all of this warp's TMEM reads of tile ti are done
    tc_fence_before();
    if constexpr (kPair) {
      mbar_arrive_leader(tempty_bar(ti & 1));   // the leader's issuer waits for both CTAs' epilogues
    } else {
      mbar_arrive(tempty_bar(ti & 1));
    }
  };

  const bool tma_res = has_res && !(!kOutF32 && p.direct_out);
  if (tma_res && lane == 0) {
    issue_res(0);
    issue_res(1);
  }
  __syncwarp();
  int cur_ti = -1;   // last tile whose accumulator this warp has seen complete
  int ti = sub / cpt, c = sub - ti * cpt;   // chunk g = nsub*l + sub, advanced incrementally
  int dec_ti = -1;
  TileCoord t{};
  for (uint32_t l = 0;; ++l) {
    // tile coordinates (integer divisions) are resolved once per tile and BEFORE the accumulator wait, off
    // the critical path; a tile index past the end decodes to harmless numbers and is never used
    if (ti != dec_ti) {
      t = decode_tile<kPair>(p, t_first + ti * t_stride);
      dec_ti = ti;
    }
    // Only the warps that own a chunk of a tile wait for its accumulator and hand it back (tempty counts 128 x
    // min(nsub, chunks per tile) arrivals, see the kernels' barrier init). When every warp arrived for every tile, a warp
    // WITHOUT a chunk in tile i (64-channel layers with two warps per quadrant: every second tile) released it only after
    // finishing its own previous tile's whole epilogue chain, and the MMA of tile i+2 waited for that. (With two warps
    // per quadrant and two accumulators every barrier has ONE owner group, so parity waits stay in step. Measured: the
    // 64-channel 1x1 layers gain a little, the ResNet stem does not move - 171 us either way.)
    if (cur_ti < ti) {
      cur_ti = ti;
      if ((long long)t_first + (long long)ti * t_stride >= p.num_tiles) break;
      mbar_wait(tfull_bar(ti & 1), (uint32_t)((ti >> 1) & 1));
      tc_fence_after();
      if (warp == 2 + 4 * sub && lane == 0) IG_STAMP(ti, 6);
    }
    const uint32_t t_acc = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)((ti & 1) * p.acc_stride);
    const uint32_t ob = l & (obufs - 1u);
    const uint32_t rb = l & 1u, rphase = (l >> 1) & 1u;
    const int ncols = min(CH, p.block_n - c * CH);
    float v[CH];
#pragma unroll
    for (int j = 0; j < CH / 16; ++j) {
      if (j * 16 < ncols) {
        tmem_ld_x16(t_acc + (uint32_t)(c * CH + j * 16), &v[j * 16]);
      } else {
#pragma unroll
        for (int q = 0; q < 16; ++q) v[j * 16 + q] = 0.f;
      }
    }
    tmem_ld_wait();
    if (c + nsub >= cpt) release_acc(ti);   // this warp's last chunk of the tile
    const float* bias_c = s_bias + t.ncol0 + c * CH;
    uint8_t* out_row = out_g + ob * kSlab;
    if constexpr (!kOutF32) {
      if (p.direct_out) {   // CTA-uniform: see IgemmParams::direct_out
        const int col0 = t.ncol0 + c * CH;
        const int row = (t.n0 == 0 && t.h0 == 0) ? t.w0 + so + lane : p.ln_rows;
        if (row < p.ln_rows) {
          __nv_bfloat16* orow = p.y_ptr + (long long)row * p.y_pitch + col0;
          const __nv_bfloat16* rrow = has_res ? p.res_ptr + (long long)row * p.res_pitch + col0 : nullptr;
          uint4 rv[8];
#pragma unroll
          for (int j = 0; j < 8; ++j)
            rv[j] = (has_res && col0 + j * 8 < p.cout) ? __ldg(reinterpret_cast<const uint4*>(rrow + j * 8))
                                                       : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (col0 + j * 8 < p.cout)
              *reinterpret_cast<uint4*>(orow + j * 8) = epilogue8<kAct, kRes>(&v[j * 8], bias_c + j * 8, rv[j]);
        }
        c += nsub;
        while (c >= cpt) {
          c -= cpt;
          ++ti;
        }
        continue;
      }
    }
    if constexpr (!kOutF32) {
      uint4 packed[8];
      if (has_res) mbar_wait(rbar(rb), rphase);
      const uint8_t* res_row = res_g + rb * kSlab;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        uint4 rv = make_uint4(0u, 0u, 0u, 0u);
        if constexpr (has_res) rv = *reinterpret_cast<const uint4*>(res_row + sw128_off(lane, j));
        packed[j] = epilogue8<kAct, kRes>(&v[j * 8], bias_c + j * 8, rv);
      }
      if (lane == 0) {   // the store that last used this staging buffer has read it
        if (obufs == 2u) tma_store_wait_read<1>(); else tma_store_wait_read<0>();
      }
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 8; ++j) *reinterpret_cast<uint4*>(out_row + sw128_off(lane, j)) = packed[j];
    } else {
#pragma unroll
      for (int q = 0; q < CH; ++q) v[q] = apply_act<kAct>(v[q] + bias_c[q]);
      if (lane == 0) {
        if (obufs == 2u) tma_store_wait_read<1>(); else tma_store_wait_read<0>();
      }
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<float4*>(out_row + sw128_off(lane, j)) =
            make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      tma_store_4d(&p.tmC, out_u32 + ob * kSlab, t.ncol0 + c * CH, t.w0 + w_off, t.h0 + h_off, t.n0 + n_off);
      tma_store_commit();
      if (tma_res) issue_res(l + 2);
      if (warp == 2 + 4 * sub) IG_STAMP(ti, 7);
    }
    __syncwarp();
    c += nsub;
    while (c >= cpt) {
      c -= cpt;
      ++ti;
    }
  }
  if (lane == 0) tma_store_wait_all();
  __syncwarp();
}

// ============================== epilogue, sixteen warps (epi_sub == 4) ==============================
// Shallow-K layers without a residual whose tile is ALL epilogue (EfficientNet / MobileNet 1x1 expansions: one K block,
// 128 x N outputs through an activation; ncu on EfficientNet-B4's 24 -> 144 @112^2: 1.6 TB/s of output with eight warps,
// issue slots 20 % busy - a latency chain per chunk with two warps per SM sub-partition to hide it). Four warps per
// TMEM lane quadrant take every fourth 64-column chunk; the chunk goes through in four 16-column steps with the next
// step's tcgen05.ld in flight (the fused bottleneck kernel's e3, csrc/bottleneck.cu): ~60 live registers, which is what
// lets 18 warps share the register file. One staging slab per warp, bf16 output, no residual, no LayerNorm folding.
constexpr int kThreads16 = 64 + 32 * 16;
// kRes != 0 (CTA-pair kernels, ResNet conv3 layers with K = 256 / 512): the warp's staging slab is also where the TMA
// puts its residual - arithmetic in place, as in the fused bottleneck kernel - so sixteen warps need 64 KiB where the
// eight-warp epilogue needs 32 + 64 KiB (staging + residual ring), and the K loop keeps its stages. The residual of the
// warp's NEXT chunk is requested as soon as its store has read the slab; the K loop of the next tile hides the trip.
template <int kAct, int kRes = 0, bool kPair = false>
__device__ __forceinline__ void epilogue_warps16(const IgemmParams& p, const uint32_t base, uint8_t* gbase,
                                                 const uint32_t tmem_base, const int warp, const int lane) {
  const int S = p.stages;
  const int t_first = kPair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int t_stride = kPair ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const uint32_t bars = base + p.off_bars;
  auto tfull_bar = [&](int a) { return bars + 8u * (2 * S + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (2 * S + 2 + a); };
  constexpr int CH = 64;
  constexpr int nsub = 4;
  constexpr bool has_res = kRes != 0;
  const int quad = warp & 3;
  const int sub = (warp - 2) >> 2;
  const int wid = quad * nsub + sub;
  const int cpt = (p.block_n + CH - 1) / CH;
  const float* s_bias = reinterpret_cast<const float*>(gbase + p.off_bias);
  const int so = quad * 32;
  const int w_off = so % p.tw, h_off = (so / p.tw) % p.th, n_off = so / (p.tw * p.th);
  constexpr uint32_t kSlab = 32 * 128;
  const uint32_t out_u32 = base + p.off_out + (uint32_t)wid * kSlab;
  uint8_t* out_row = gbase + p.off_out + (uint32_t)wid * kSlab;
  const uint32_t rbar = bars + 8u * (48u + (uint32_t)wid);
  auto release_acc = [&](int ti) {
    tc_fence_before();
    if constexpr (kPair) {
      mbar_arrive_leader(tempty_bar(ti & 1));
    } else {
      mbar_arrive(tempty_bar(ti & 1));
    }
  };
  auto issue_res = [&](int ti, int c) {   // ONE lane: the residual of chunk (ti, c) into this warp's slab
    const long long tile = (long long)t_first + (long long)ti * t_stride;
    if (tile >= p.num_tiles) return;
    const TileCoord t = decode_tile<kPair>(p, (int)tile);
    mbar_expect_tx(rbar, kSlab);
    tma_load_4d(out_u32, &p.tmR, rbar, t.ncol0 + c * CH, t.w0 + w_off, t.h0 + h_off, t.n0 + n_off);
  };
  int cur_ti = -1;
  int ti = sub / cpt, c = sub - ti * cpt;
  int dec_ti = -1;
  TileCoord t{};
  if (has_res && lane == 0) issue_res(ti, c);
  __syncwarp();
  for (uint32_t l = 0;; ++l) {
    if (ti != dec_ti) {
      t = decode_tile<kPair>(p, t_first + ti * t_stride);
      dec_ti = ti;
    }
    // every warp waits for and releases EVERY tile here: with four warps per quadrant and two accumulators a barrier is
    // shared by two owner groups, and a parity wait cannot skip phases (owner-only hand-back as in epilogue_warps would
    // need one accumulator per group)
    bool done = false;
    while (cur_ti < ti) {
      ++cur_ti;
      if ((long long)t_first + (long long)cur_ti * t_stride >= p.num_tiles) {
        done = true;
        break;
      }
      mbar_wait(tfull_bar(cur_ti & 1), (uint32_t)((cur_ti >> 1) & 1));
      tc_fence_after();
      if (cur_ti < ti) release_acc(cur_ti);
    }
    if (done) break;
    const uint32_t t_acc = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)((ti & 1) * p.acc_stride) + (uint32_t)(c * CH);
    const int ncols = min(CH, p.block_n - c * CH);   // multiple of 16
    const float* bias_c = s_bias + t.ncol0 + c * CH;
    float va[16], vb[16];
    tmem_ld_x16(t_acc, va);
    if constexpr (has_res) {
      mbar_wait(rbar, l & 1u);   // the residual slab has landed (requested after the previous store had read the slab)
    } else {
      if (lane == 0) tma_store_wait_read<0>();   // the store that last used this warp's slab has read it
      __syncwarp();
    }
#pragma unroll
    for (int st = 0; st < 4; ++st) {
      float* cur = (st & 1) ? vb : va;
      float* nxt = (st & 1) ? va : vb;
      uint4 o0 = make_uint4(0u, 0u, 0u, 0u), o1 = o0;
      if (st * 16 < ncols) {   // warp-uniform
        uint4 r0 = o0, r1 = o0;
        if constexpr (has_res) {
          r0 = *reinterpret_cast<const uint4*>(out_row + sw128_off(lane, 2 * st));
          r1 = *reinterpret_cast<const uint4*>(out_row + sw128_off(lane, 2 * st + 1));
        }
        tmem_ld_wait();
        if (st < 3 && (st + 1) * 16 < ncols) tmem_ld_x16(t_acc + (uint32_t)(16 * (st + 1)), nxt);
        o0 = epilogue8<kAct, kRes>(&cur[0], bias_c + 16 * st, r0);
        o1 = epilogue8<kAct, kRes>(&cur[8], bias_c + 16 * st + 8, r1);
      }
      *reinterpret_cast<uint4*>(out_row + sw128_off(lane, 2 * st)) = o0;
      *reinterpret_cast<uint4*>(out_row + sw128_off(lane, 2 * st + 1)) = o1;
    }
    if (c + nsub >= cpt) release_acc(ti);   // this warp's last chunk of the tile
    fence_proxy_async_smem();
    __syncwarp();
    const int c_old = c, ti_old = ti;
    c += nsub;
    while (c >= cpt) {
      c -= cpt;
      ++ti;
    }
    if (lane == 0) {
      tma_store_4d(&p.tmC, out_u32, t.ncol0 + c_old * CH, t.w0 + w_off, t.h0 + h_off, t.n0 + n_off);
      tma_store_commit();
      if constexpr (has_res) {
        tma_store_wait_read<0>();
        issue_res(ti, c);
      }
    }
    (void)ti_old;
    __syncwarp();
  }
  if (lane == 0) tma_store_wait_all();
  __syncwarp();
}

// kRes: 0 = no residual, 1 = act(acc + bias + res), 2 = act(acc + bias) + res. Compile-time so that the
// unrolled epilogue is straight-line code (a runtime flag doubled its instruction count and made the
// epilogue warps issue-bound on the HBM-bound layers: profiles/r01_layers_v4).
template <bool kOutF32, int kAct, int kRes, int kLN = 0, bool kGate = false, bool kEpi16 = false>
__global__ void __launch_bounds__(kEpi16 ? kThreads16 : (kGate ? kThreads + 128 : kThreads), 1) igemm_kernel(const __grid_constant__ IgemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - raw);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int S = p.stages;
  const bool gather = (kEpi16 || kGate) && p.ga_on;   // narrow-K layers: A gathered by threads, filter resident (IgemmParams::ga_*)
  const uint32_t stage_bytes = gather ? (uint32_t)kABytes : (uint32_t)(kABytes + p.block_n * 128);

  const uint32_t bars = base + p.off_bars;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (S + s); };
  auto tfull_bar = [&](int a) { return bars + 8u * (2 * S + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (2 * S + 2 + a); };
  const uint32_t tmem_slot = bars + 8u * (2 * S + 12);
  volatile uint32_t* tmem_slot_g =
      reinterpret_cast<volatile uint32_t*>(gbase + p.off_bars + 8 * (2 * S + 12));

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB);
    tma_prefetch_desc(&p.tmC);
    if (p.has_res) tma_prefetch_desc(&p.tmR);
    for (int s = 0; s < S; ++s) {
      mbar_init(full_bar(s), gather ? 32u : 1u);   // gathered A: + the 32 lanes of warp 0
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 128u * (uint32_t)(p.epi_sub == 4 ? 4 : min(p.epi_sub, (p.block_n + (kOutF32 ? 31 : 63)) / (kOutF32 ? 32 : 64))));   // the warps that own a chunk of the tile
    }
    for (int b = 0; b < 16; ++b) mbar_init(bars + 8u * (48 + b), 1);   // residual ring (epilogue_warps)
    if (kGate)
      for (int b = 0; b < S; ++b) mbar_init(bars + 8u * (32 + b), 128);   // xf[stage]: the 128 gate threads
    mbar_fence_init();
  }
  // folded-BN shift / bias for every output column, once per CTA (zero beyond cout)
  {
    const int ncols_pad = p.n_tiles * p.block_n + 64;
    stage_columns(reinterpret_cast<float*>(gbase + p.off_bias), p.bias, p.cout, ncols_pad);
    if (kLN == 1)   // column sums of the gamma-folded filter, next to the bias
      stage_columns(reinterpret_cast<float*>(gbase + p.off_wsum), p.ln_wsum, p.cout, ncols_pad);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    tmem_relinquish();
  }
  griddep_wait();     // PDL: everything above overlapped the previous kernel's tail
  tc_fence_before();
  __syncthreads();
  griddep_launch();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_g;


  if (kGate && gather && warp == 0) {
    // gated narrow-K layers: the gate warps below fetch, scale and stage the A rows themselves; nothing to do here
  } else if (gather && warp == 0) {
    // ============================== gathering producer (narrow K, see IgemmParams::ga_*) ==============================
    // lane <-> rows lane, lane + 32, ... of a 128-row tile; chunk j of row r sits at r*128 + ((j ^ (r & 7)) << 4) (the
    // layout a SWIZZLE_128B TMA box would have written). The chunks beyond K are zeroed once per stage and never touched
    // again. The filter (<= 256 rows of k < 64) is gathered the same way ONCE, before the first tile is signalled.
    const int k16 = p.ga_k16;
    const uint32_t zero_chunks = (uint32_t)(8 - k16);
    const uint32_t rows_total = (uint32_t)S * 128u + (uint32_t)p.block_n;   // A stages (16 KiB each), then the filter slab
    for (uint32_t idx = (uint32_t)lane; idx < rows_total * zero_chunks; idx += 32u) {
      const uint32_t rr = idx / zero_chunks, j = (uint32_t)k16 + idx % zero_chunks;
      uint8_t* rowp = rr < (uint32_t)S * 128u ? gbase + rr * 128u : gbase + p.off_bres + (rr - (uint32_t)S * 128u) * 128u;
      const uint32_t r = rr < (uint32_t)S * 128u ? (rr & 127u) : rr - (uint32_t)S * 128u;
      *reinterpret_cast<uint4*>(rowp + ((j ^ (r & 7u)) << 4)) = make_uint4(0u, 0u, 0u, 0u);
    }
    for (int r = lane; r < p.block_n; r += 32) {
      const bool ok = r < p.cout;
      const __nv_bfloat16* src = p.gb_ptr + (long long)(ok ? r : 0) * p.gb_pitch;
      const uint32_t drow = base + (uint32_t)p.off_bres + (uint32_t)r * 128u;
      for (int j = 0; j < k16; ++j) cp_async_16(drow + (((uint32_t)j ^ ((uint32_t)r & 7u)) << 4), src + 8 * j, ok);
    }
    cp_async_commit();
    cp_async_wait<0>();
    fence_proxy_async_smem();
    constexpr int kLookahead = 3;   // tiles in flight (cp.async groups)
    int stage = 0, cstage = 0, pending = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const TileCoord t = decode_tile(p, tile);
      mbar_wait(empty_bar(stage), phase ^ 1u);
      const uint32_t dst = base + (uint32_t)stage * stage_bytes;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = lane + 32 * i;
        const int gr = t.w0 + r;
        const bool ok = gr < p.ga_rows;
        const __nv_bfloat16* src = p.ga_ptr + (long long)(ok ? gr : 0) * p.ga_pitch;
        const uint32_t drow = dst + (uint32_t)r * 128u;
        for (int j = 0; j < k16; ++j) cp_async_16(drow + (((uint32_t)j ^ ((uint32_t)r & 7u)) << 4), src + 8 * j, ok);
      }
      cp_async_commit();
      if (++pending == kLookahead) {
        cp_async_wait<kLookahead - 1>();
        fence_proxy_async_smem();
        mbar_arrive(full_bar(cstage));
        if (++cstage == S) cstage = 0;
        --pending;
      }
      if (++stage == S) {
        stage = 0;
        phase ^= 1u;
      }
    }
    cp_async_wait<0>();
    fence_proxy_async_smem();
    for (; pending > 0; --pending) {
      mbar_arrive(full_bar(cstage));
      if (++cstage == S) cstage = 0;
    }
  } else if (warp == 0) {
    // ============================== TMA producer ==============================
    // One elected thread runs the whole role (same reasoning as the MMA issuer below): per K block a
    // barrier wait, an expect_tx and two bulk-tensor loads; ring addresses advance incrementally.
    if (elect_one()) {
      uint32_t dst = base, fb = full_bar(0), eb = empty_bar(0);
      const uint32_t dst_end = base + (uint32_t)S * stage_bytes, fb0 = fb, eb0 = eb;
      uint32_t phase = 0;
      const int kh = p.kh, kw = p.kw, kchunks = p.kchunks;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const TileCoord t = decode_tile(p, tile);
        for (int r = 0; r < kh; ++r) {
          const int hc = t.h0 * p.mul_h + r * p.dil_h - p.pad_h;
          if (hc + (p.th - 1) * p.mul_h < 0 || hc >= p.in_h) continue;  // tap row entirely in the padding
          for (int s = 0; s < kw; ++s) {
            const int wc = t.w0 * p.mul_w + s * p.dil_w - p.pad_w;
            if (wc + (p.tw - 1) * p.mul_w < 0 || wc >= p.in_w) continue;
            int kb = (r * kw + s) * p.cin_pack;
            for (int c = 0; c < kchunks; ++c, kb += kBlockK) {
              mbar_wait(eb, phase ^ 1u);
              mbar_expect_tx(fb, stage_bytes);
              const int ac = (c == kchunks - 1 && p.a_tail >= 0) ? p.a_tail : c * kBlockK;   // K_TAIL_SHIFT: in-bounds box
              tma_load_4d(dst, &p.tmA, fb, ac + (p.grouped ? t.ncol0 : 0), wc, hc, t.n0);
              tma_load_2d(dst + kABytes, &p.tmB, fb, kb, t.ncol0);
              dst += stage_bytes, fb += 8, eb += 8;
              if (dst == dst_end) {
                dst = base, fb = fb0, eb = eb0;
                phase ^= 1u;
              }
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (kGate && warp >= 6 && gather) {
    // ============================== SE gate + gather (narrow K: k < 64, one K chunk, <= 256 outputs) ==============================
    // EfficientNet-B4's first two projections (48 -> 24 and 24 -> 24 @112^2, 1.6 M rows): the 64-channel TMA boxes of
    // both operands reach past their tensors and crawl (116 + 201 us for 36 + 24 us of traffic). Here the four gate
    // warps ARE the producers: thread = one row of the tile, read with 16-byte loads one tile ahead, multiplied by the
    // image's gate (HMUL2: the rounding the separate pass had) and written straight into the 128B-swizzled operand
    // layout; the K-pad chunks are zeroed once, the filter is staged once and stays resident.
    const int r = (warp - 6) * 32 + lane;
    const int k16 = p.ga_k16;
    const uint32_t xor7 = (uint32_t)(r & 7);
    for (int st = 0; st < S; ++st)
      for (int j = k16; j < 8; ++j)
        *reinterpret_cast<uint4*>(gbase + (uint32_t)st * stage_bytes + r * 128 + (((uint32_t)j ^ xor7) << 4)) =
            make_uint4(0u, 0u, 0u, 0u);
    for (int n = r; n < p.block_n; n += 128) {
      uint8_t* brow = gbase + p.off_bres + n * 128;
      const __nv_bfloat16* wrow = p.gb_ptr + (long long)min(n, p.cout - 1) * p.gb_pitch;
      for (int j = 0; j < 8; ++j) {
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (j < k16 && n < p.cout) v = __ldg(reinterpret_cast<const uint4*>(wrow + 8 * j));
        *reinterpret_cast<uint4*>(brow + (((uint32_t)j ^ (uint32_t)(n & 7)) << 4)) = v;
      }
    }
    fence_proxy_async_smem();
    auto load_row = [&](int tile, uint4* a) {
      int row = tile < p.num_tiles ? decode_tile(p, tile).w0 + r : p.ga_rows;
      const bool ok = row < p.ga_rows;
      const __nv_bfloat16* src = p.ga_ptr + (long long)(ok ? row : 0) * p.ga_pitch;
#pragma unroll
      for (int j = 0; j < 7; ++j)
        a[j] = (ok && j < k16) ? __ldg(reinterpret_cast<const uint4*>(src + 8 * j)) : make_uint4(0u, 0u, 0u, 0u);
    };
    uint4 a[7], an[7];
    load_row((int)blockIdx.x, a);
    uint32_t eb = empty_bar(0), xb = bars + 8u * 32u;
    const uint32_t eb0 = eb, xb0 = xb;
    uint32_t phase = 0;
    uint8_t* rowp = gbase + r * 128;
    uint8_t* const row0 = rowp;
    int sidx = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const TileCoord t = decode_tile(p, tile);
      const int row = min(t.w0 + r, p.ln_rows - 1);
      const __nv_bfloat16* grow = p.gate + (long long)(row / p.gate_rpi) * p.gate_pitch;
      uint4 g[7];
#pragma unroll
      for (int j = 0; j < 7; ++j)
        g[j] = j < k16 ? __ldg(reinterpret_cast<const uint4*>(grow + 8 * j)) : make_uint4(0u, 0u, 0u, 0u);
      load_row(tile + (int)gridDim.x, an);   // next tile's row in flight behind this tile's arithmetic
      mbar_wait(eb, phase ^ 1u);
#pragma unroll
      for (int j = 0; j < 7; ++j) {
        if (j < k16) {
          __nv_bfloat162* a2 = reinterpret_cast<__nv_bfloat162*>(&a[j]);
          const __nv_bfloat162* g2 = reinterpret_cast<const __nv_bfloat162*>(&g[j]);
#pragma unroll
          for (int e = 0; e < 4; ++e) a2[e] = __hmul2(a2[e], g2[e]);
          *reinterpret_cast<uint4*>(rowp + (((uint32_t)j ^ xor7) << 4)) = a[j];
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(xb);
#pragma unroll
      for (int j = 0; j < 7; ++j) a[j] = an[j];
      eb += 8, xb += 8, rowp += stage_bytes;
      if (++sidx == S) {
        sidx = 0, eb = eb0, xb = xb0, rowp = row0;
        phase ^= 1u;
      }
    }
  } else if (kGate && warp >= 6) {
    // ============================== SE gate on the A operand (warps 6..9, kGate kernels) ==============================
    // SqueezeExcitation `x * scale` (layers/squeeze.py:61) feeding the project convolution (efficientnet.py:161-170):
    // instead of a separate pass that reads and rewrites the expanded tensor, these four warps scale every staged A
    // tile IN SHARED MEMORY - thread = one row of the tile; 16-byte chunk j of row r of the 128B-swizzled tile sits at
    // r*128 + ((j ^ (r & 7)) << 4); bf16 x bf16 products rounded to bf16 with HMUL2, exactly what the separate pass
    // stored - make the writes visible to the async proxy and arrive on xf[stage], which the MMA issuer waits for in
    // place of full[stage]. Flat GEMMs (1x1 convolutions) only; the epilogue runs on four warps (epi_sub == 1).
    // One or two groups of four warps (blockDim decides): with two, the groups take alternate K blocks - the chain
    // wait -> 8 LDS -> 32 HMUL2 -> 8 STS -> proxy fence -> arrive is ~500 cycles per K block and thread against 64..256
    // tensor cycles, and it, not TMA or the tensor pipe, bounded every gated projection (144 -> 32 @56^2: 47 us gated,
    // 27 us ungated).
    const int ngrp = ((int)blockDim.x - 192) >> 7;
    const int grp = (warp - 6) >> 2;
    const int r = ((warp - 6) & 3) * 32 + lane;
    const uint32_t xor7 = (uint32_t)(r & 7);
    uint32_t fb = full_bar(0);
    const uint32_t fb0 = fb;
    uint32_t xb = bars + 8u * 32u;
    const uint32_t xb0 = xb;
    uint32_t phase = 0;
    uint8_t* rowp = gbase + r * 128;
    uint8_t* const row0 = rowp;
    int sidx = 0;
    int turn = 0;   // K blocks are dealt round-robin to the groups
    const int kchunks = p.kchunks;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const TileCoord t = decode_tile(p, tile);
      const int row = min(t.w0 + r, p.ln_rows - 1);   // rows past the end: zero-filled A, any gate will do
      const __nv_bfloat16* grow = p.gate + (long long)(row / p.gate_rpi) * p.gate_pitch;
      for (int i = 0; i < kchunks; ++i) {
        const bool mine = turn == grp;
        if (++turn == ngrp) turn = 0;
        if (!mine) {
          fb += 8, xb += 8, rowp += stage_bytes;
          if (++sidx == S) {
            sidx = 0, fb = fb0, xb = xb0, rowp = row0;
            phase ^= 1u;
          }
          continue;
        }
        const int k0 = (i == kchunks - 1 && p.a_tail >= 0) ? p.a_tail : i * kBlockK;
        uint4 g[8];
#pragma unroll
        for (int j = 0; j < 8; ++j)   // the gate does not depend on the tile data: in flight during the barrier wait
          g[j] = (k0 + j * 8 < p.gate_cols) ? __ldg(reinterpret_cast<const uint4*>(grow + k0 + j * 8))
                                            : make_uint4(0u, 0u, 0u, 0u);
        mbar_wait(fb, phase);
        uint4 a[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = *reinterpret_cast<const uint4*>(rowp + (((uint32_t)j ^ xor7) << 4));
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          __nv_bfloat162* a2 = reinterpret_cast<__nv_bfloat162*>(&a[j]);
          const __nv_bfloat162* g2 = reinterpret_cast<const __nv_bfloat162*>(&g[j]);
#pragma unroll
          for (int e = 0; e < 4; ++e) a2[e] = __hmul2(a2[e], g2[e]);
          *reinterpret_cast<uint4*>(rowp + (((uint32_t)j ^ xor7) << 4)) = a[j];
        }
        fence_proxy_async_smem();   // generic-proxy writes -> visible to the tensor core's async-proxy reads
        mbar_arrive(xb);
        fb += 8, xb += 8, rowp += stage_bytes;
        if (++sidx == S) {
          sidx = 0, fb = fb0, xb = xb0, rowp = row0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    // ============================== MMA issuer ==============================
    // ONE thread runs the whole role. Per 64-deep K block it needs: a barrier wait, 4 UTCHMMA, a
    // commit. Everything else is kept out of the loop (running descriptor / barrier addresses, the
    // number of K blocks of the tile computed once per tile), because this thread's instruction
    // latency -- not the tensor pipe -- bounded every layer (profiles/r01_stem_v3: ~70 SASS
    // instructions and ~650 cycles per K block against 128..512 cycles of MMA work).
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_bf16_m128((uint32_t)p.block_n);
      const uint32_t desc_hi = (uint32_t)(umma_desc_sw128(0) >> 32);
      const uint32_t lbo = 1u << 16;
      const uint32_t step = stage_bytes >> 4;
      const uint32_t a_lo0 = ((base & 0x3FFFF) >> 4) | lbo;
      const uint32_t a_end = a_lo0 + (uint32_t)S * step;
      const uint32_t b_off = kABytes >> 4;
      const uint32_t bres_lo = (((base + (uint32_t)p.off_bres) & 0x3FFFF) >> 4) | lbo;   // resident filter (gather mode)
      uint32_t a_lo = a_lo0;
      uint32_t fb = kGate ? bars + 8u * 32u : full_bar(0), eb = empty_bar(0);
      const uint32_t fb0 = fb, eb0 = eb;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      const int kh = p.kh, kw = p.kw, kchunks = p.kchunks;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        int nk = kchunks;
        if (kh * kw > 1) {  // count the taps the producer does not skip (same predicate)
          const TileCoord t = decode_tile(p, tile);
          int vr = 0, vs = 0;
          for (int r = 0; r < kh; ++r) {
            const int hc = t.h0 * p.mul_h + r * p.dil_h - p.pad_h;
            vr += !(hc + (p.th - 1) * p.mul_h < 0 || hc >= p.in_h);
          }
          for (int s2 = 0; s2 < kw; ++s2) {
            const int wc = t.w0 * p.mul_w + s2 * p.dil_w - p.pad_w;
            vs += !(wc + (p.tw - 1) * p.mul_w < 0 || wc >= p.in_w);
          }
          nk = vr * vs * kchunks;
        }
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.acc_stride);
        uint32_t accumulate = 0;
        for (int i = 0; i < nk; ++i) {
          mbar_wait(fb, phase);      // kGate: xf[stage], raised by the gate warps once the A tile is scaled
          tc_fence_after();
          umma_bf16_kblock64(d_tmem, a_lo, gather ? bres_lo : a_lo + b_off, desc_hi, desc_hi, idesc, accumulate, eb);
          accumulate = 1;
          a_lo += step, fb += 8, eb += 8;
          if (a_lo == a_end) {
            a_lo = a_lo0, fb = fb0, eb = eb0;
            phase ^= 1u;
          }
        }
        umma_commit(tfull_bar(acc));  // accumulator complete -> epilogue
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
    __syncwarp();
  } else {
    if constexpr (kEpi16) {
      epilogue_warps16<kAct, 0, false>(p, base, gbase, tmem_base, warp, lane);
    } else {
      epilogue_warps<kOutF32, kAct, kRes, false, kLN>(p, base, gbase, tmem_base, warp, lane);
    }
  }

  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2) for the K-deep layers.
//
// With one CTA per tile every 64-deep K block moves 16 KiB of A plus block_n*128 B of B into the SM
// (48 KiB for a 128x256 tile) against 512 tensor-pipe cycles of work; the layers with K >= 256 ran at
// ~900 cycles per K block, i.e. at the L2->SM fill rate (~54 B/cycle/SM), not at the MMA rate. Here
// two CTAs of a cluster (one TPC) compute a 256 x block_n tile with ONE tcgen05.mma.cta_group::2
// stream issued by the leader: each CTA loads only its own 128 A rows and HALF of the B slab, so the
// fill per SM and K block drops to 16 KiB + block_n*64 B (32 KiB) for the same 512 cycles.
// Protocol: full[s] lives in the leader (both producers' TMA bytes are credited to it), the MMA
// commits are multicast to empty[s] / tfull[a] of both CTAs, both epilogues arrive on the leader's
// tempty[a]. The epilogue itself is the single-CTA one (each CTA drains its own 128 TMEM lanes).
template <int kAct, int kRes, int kLN = 0, bool kEpi16 = false>
__device__ __forceinline__ void pair_kernel_body(const IgemmParams& p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - raw);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int S = p.stages;
  const uint32_t rank = blockIdx.x & 1u;            // == %cluster_ctarank for a (2,1,1) cluster
  const uint32_t b_half = (uint32_t)p.block_n * 64u;   // bytes of this CTA's half of the B slab
  const uint32_t stage_bytes = kABytes + b_half;

  const uint32_t bars = base + p.off_bars;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (S + s); };
  auto tfull_bar = [&](int a) { return bars + 8u * (2 * S + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (2 * S + 2 + a); };
  const uint32_t tmem_slot = bars + 8u * (2 * S + 12);
  volatile uint32_t* tmem_slot_g =
      reinterpret_cast<volatile uint32_t*>(gbase + p.off_bars + 8 * (2 * S + 12));

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB);
    tma_prefetch_desc(&p.tmC);
    if (p.has_res) tma_prefetch_desc(&p.tmR);
    for (int s = 0; s < S; ++s) {
      mbar_init(full_bar(s), 1);    // used in the leader only: its producer's arrive.expect_tx
      mbar_init(empty_bar(s), 1);   // one multicast commit per phase
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 256u * (uint32_t)(p.epi_sub == 4 ? 4 : min(p.epi_sub, (p.block_n + 63) / 64)));  // leader: the chunk-owning epilogue warps of both CTAs
    }
    for (int b = 0; b < 16; ++b) mbar_init(bars + 8u * (48 + b), 1);   // residual ring (epilogue_warps)
    mbar_fence_init();
  }
  {
    const int ncols_pad = p.n_tiles * p.block_n + 64;
    stage_columns(reinterpret_cast<float*>(gbase + p.off_bias), p.bias, p.cout, ncols_pad);
    if (kLN == 1)   // column sums of the gamma-folded filter, next to the bias
      stage_columns(reinterpret_cast<float*>(gbase + p.off_wsum), p.ln_wsum, p.cout, ncols_pad);
  }
  if (warp == 1) {
    tmem_alloc_pair(tmem_slot, (uint32_t)p.tmem_cols);
    tmem_relinquish_pair();
  }
  griddep_wait();     // PDL: everything above overlapped the previous kernel's tail
  tc_fence_before();
  __syncthreads();
  griddep_launch();
  cluster_sync_all();   // the peer's barriers are initialised before anything is signalled remotely
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_g;
  const int t_first = (int)(blockIdx.x >> 1), t_stride = (int)(gridDim.x >> 1);

  if (warp == 0) {
    // ============================== TMA producer (both CTAs) ==============================
    if (elect_one()) {
      uint32_t dst = base, fb = full_bar(0), eb = empty_bar(0);
      const uint32_t dst_end = base + (uint32_t)S * stage_bytes, fb0 = fb, eb0 = eb;
      uint32_t phase = 0;
      const int kh = p.kh, kw = p.kw, kchunks = p.kchunks;
      const int nrow = (int)rank * (p.block_n / 2);   // this CTA's rows of the B slab
      for (int tile = t_first; tile < p.num_tiles; tile += t_stride) {
        const TileCoord t = decode_tile<true>(p, tile);
        for (int r = 0; r < kh; ++r) {
          const int hc = t.h0 * p.mul_h + r * p.dil_h - p.pad_h;
          for (int s = 0; s < kw; ++s) {
            const int wc = t.w0 * p.mul_w + s * p.dil_w - p.pad_w;
            int kb = (r * kw + s) * p.cin_pack;
            for (int c = 0; c < kchunks; ++c, kb += kBlockK) {
              mbar_wait(eb, phase ^ 1u);
              if (rank == 0) mbar_expect_tx(fb, 2u * stage_bytes);   // both CTAs' bytes land on this barrier
              tma_load_4d_pair(dst, &p.tmA, fb, (c == kchunks - 1 && p.a_tail >= 0) ? p.a_tail : c * kBlockK, wc, hc, t.n0);
              tma_load_2d_pair(dst + kABytes, &p.tmB, fb, kb, t.ncol0 + nrow);
              dst += stage_bytes, fb += 8, eb += 8;
              if (dst == dst_end) {
                dst = base, fb = fb0, eb = eb0;
                phase ^= 1u;
              }
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ============================== MMA issuer (leader CTA only) ==============================
    if (rank == 0 && elect_one()) {
      const uint32_t idesc = umma_idesc_bf16_m256((uint32_t)p.block_n);
      const uint32_t desc_hi = (uint32_t)(umma_desc_sw128(0) >> 32);
      const uint32_t lbo = 1u << 16;
      const uint32_t step = stage_bytes >> 4;
      const uint32_t a_lo0 = ((base & 0x3FFFF) >> 4) | lbo;
      const uint32_t a_end = a_lo0 + (uint32_t)S * step;
      const uint32_t b_off = kABytes >> 4;
      uint32_t a_lo = a_lo0;
      uint32_t fb = full_bar(0), eb = empty_bar(0);
      const uint32_t fb0 = fb, eb0 = eb;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      const int nk = p.kh * p.kw * p.kchunks;
      for (int tile = t_first; tile < p.num_tiles; tile += t_stride) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.acc_stride);
        uint32_t accumulate = 0;
        for (int i = 0; i < nk; ++i) {
          mbar_wait(fb, phase);
          tc_fence_after();
          umma_bf16_kblock64_pair(d_tmem, a_lo, a_lo + b_off, desc_hi, desc_hi, idesc, accumulate, eb);
          accumulate = 1;
          a_lo += step, fb += 8, eb += 8;
          if (a_lo == a_end) {
            a_lo = a_lo0, fb = fb0, eb = eb0;
            phase ^= 1u;
          }
        }
        umma_commit_pair(tfull_bar(acc));
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
    __syncwarp();
  } else {
    if constexpr (kEpi16) {
      epilogue_warps16<kAct, kRes, true>(p, base, gbase, tmem_base, warp, lane);
    } else {
      epilogue_warps<false, kAct, kRes, true, kLN>(p, base, gbase, tmem_base, warp, lane);
    }
  }

  // ---- teardown: neither CTA may leave while its peer can still touch its smem / TMEM / barriers ----
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, (uint32_t)p.tmem_cols);
  }
}

template <int kAct, int kRes, int kLN = 0>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
    pair_kernel(const __grid_constant__ IgemmParams p) {
  pair_kernel_body<kAct, kRes, kLN>(p);
}
template <int kAct, int kRes>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads16, 1)
    pair_kernel16(const __grid_constant__ IgemmParams p) {
  pair_kernel_body<kAct, kRes, 0, true>(p);
}

// ------------------------------------------------------------------------------------------------
// First-layer convolution (Cin <= 8 on the raw image): halo tile + overlapping UMMA descriptors.
//
// The generic path re-reads every input pixel once per filter row through TMA (7x for the ResNet
// stem: measured L2->SM bound, 0.46 ms for B=256). Here the padded NHWC8 image tile needed by 128
// outputs (8 wide x 16 high) is staged ONCE, un-swizzled, and every (filter row r, tap pair 2j/2j+1)
// MMA reads it in place through a SWIZZLE_NONE K-major descriptor:
//   * a core matrix = 8 consecutive output columns x one tap (8 channels, 16 B): with the image
//     columns de-interleaved into `stride` phase planes (a TMA box with element stride `stride`),
//     those 8 rows are 8 consecutive 16-byte pixels of one plane -> the canonical 128-byte core matrix;
//   * LBO (second core matrix along K = the next tap) = one pixel (stride 1) or one plane (stride 2/4);
//   * SBO (next 8 rows = next output row) = `stride` input rows of the plane.
// Different taps are the same bytes at shifted start addresses; nothing is replicated in shared
// memory. The whole packed filter (kh x [N x 64] slabs) stays resident for the lifetime of the CTA.
// Call sites replaced: resnet.py:243-251 (7x7 s2), vgg.py:137 (3x3 s1), efficientnet.py:327 /
// mobilenetv3.py:193 (3x3 s2), swin.py:705-713 (4x4 s4).
constexpr int kStemProducers = 4;                       // warps 0, 6, 7, 8
constexpr int kStemThreads = kThreads + 32 * (kStemProducers - 1);   // warps 0 (producer), 1 (MMA), 2..9 epilogue, 10..12 producers
                                                                   // (epi_sub == 1: 2..5 epilogue, 6..8 producers, 288 threads)

// The halo tile is gathered with cp.async: its natural granule is one pixel (16 B), which a TMA box can
// only move as one request per pixel (measured ~5900 cycles per tile). Here a warp instruction moves
// one contiguous image-row segment and de-interleaves the column phases on the fly. One warp alone is
// issue-bound on the address arithmetic (profiles/r01_stem_v2: the producer warp never idles), so the
// rows of a tile are dealt round-robin to kStemProducers warps; every lane arrives on the stage's
// barrier after ITS copies have landed and been fenced for the async proxy (tcgen05.mma reads).
template <int kLookahead>   // tiles in flight per warp (cp.async groups); needs kLookahead + 2 stages
__device__ __forceinline__ void stem_gather_la(const IgemmParams& p, const uint32_t base, const int pw) {
  const int S = p.stages;
  const uint32_t bars = base + p.off_bars;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (S + s); };
  const int lane = threadIdx.x & 31;
  const int hp = p.in_h, wp = p.in_w;
  const uint8_t* img = reinterpret_cast<const uint8_t*>(p.h_src);
  const int span = p.h_px * p.h_stride;  // input pixels per halo row
  const long long row_b = (long long)wp * 16;
  const uint32_t drow_b = (uint32_t)(p.h_px * 16);
  int stage = 0, cstage = 0, pending = 0;
  uint32_t phase = 0;
  for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
    const TileCoord t = decode_tile(p, tile);
    mbar_wait(empty_bar(stage), phase ^ 1u);
    const uint32_t a_dst = base + stage * p.h_stage_bytes;
    const int gy0 = t.h0 * p.h_stride_y, gx0 = t.w0 * p.h_stride;
    const int rows_ok = min(p.h_rows, hp - gy0);
    for (int px = lane; px < span; px += 32) {
      const int gx = gx0 + px;
      const bool okx = gx < wp;
      uint32_t dst = a_dst + (uint32_t)((px % p.h_stride) * p.h_plane_pitch + (px / p.h_stride) * 16) +
                     (uint32_t)pw * drow_b;
      const uint8_t* src = img + (((long long)t.n0 * hp + gy0 + pw) * wp + (okx ? gx : 0)) * 16;
      int y = pw;
      if (okx) {
        for (; y < rows_ok; y += kStemProducers) {
          cp_async_16(dst, src, true);
          dst += kStemProducers * drow_b;
          src += kStemProducers * row_b;
        }
      }
      for (; y < p.h_rows; y += kStemProducers) {  // beyond the padded image: zeros
        cp_async_16(dst, img, false);
        dst += kStemProducers * drow_b;
      }
    }
    cp_async_commit();
    if (++pending == kLookahead) {
      cp_async_wait<kLookahead - 1>();
      fence_proxy_async_smem();
      mbar_arrive(full_bar(cstage));
      if (++cstage == S) cstage = 0;
      --pending;
    }
    if (++stage == S) {
      stage = 0;
      phase ^= 1u;
    }
  }
  cp_async_wait<0>();
  fence_proxy_async_smem();
  for (; pending > 0; --pending) {
    mbar_arrive(full_bar(cstage));
    if (++cstage == S) cstage = 0;
  }
}

// Three tiles in flight per gather warp (a deeper ring bought nothing: the gather warps are ISSUE bound, see the TMA path
// of the pixel-pair layout in stem_kernel).
__device__ __forceinline__ void stem_gather(const IgemmParams& p, const uint32_t base, const int pw) {
  stem_gather_la<3>(p, base, pw);   // six in flight measured SLOWER on the 8-channel layout (199 vs 174 us)
}

// First-layer epilogue with the max-pool that follows it (resnet.py:243-253: conv1 -> bn1 -> relu -> maxpool 3x3 / 2 / 1):
// the 411 MB conv1 output of a 256-image batch is neither written nor read back. The four warps that drain a tile
// (8 columns x 16 rows of conv outputs) stage it in shared memory, then the same 128 threads reduce the 3x3 windows:
// pooled pixel (ph, pw) needs conv rows 2ph-1..2ph+1 and columns 2pw-1..2pw+1, so with tile origins at multiples of
// (16, 8) the pooled pixels 1..7 x 1..3 of a tile are complete inside it (plain 16-byte stores), the first / last
// pooled row and column (ph % 8 == 0, pw % 4 == 0) collect contributions from two or four tiles: those go through
// red.global.max (bf16 x 8 per instruction) on memory the host entry zeroed - exact and order-independent because
// ReLU outputs are >= 0 (0 is the identity of max on them; the pool's own padding is -inf, i.e. "skip the tap").
// Requires conv output extents that are multiples of (16, 8), cout == 64, ReLU: the entry checks.
__device__ __forceinline__ uint4 bf16x8_max(const uint4 a, const uint4 b) {
  uint4 r;
  const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&a);
  const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(&b);
  __nv_bfloat162* pr = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) pr[i] = __hmax2(pa[i], pb[i]);
  return r;
}
__device__ __forceinline__ void red_max_bf16x8(void* ptr, const uint4 v) {
  asm volatile("red.global.max.noftz.v4.bf16x2 [%0], {%1, %2, %3, %4};" ::"l"(ptr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ void stem_pool_epilogue(const IgemmParams& p, const uint32_t base, uint8_t* gbase,
                                                   const uint32_t tmem_base, const int warp, const int lane) {
  const int S = p.stages;
  const int t_first = (int)blockIdx.x, t_stride = (int)gridDim.x;
  const uint32_t bars = base + p.off_bars;
  auto tfull_bar = [&](int a) { return bars + 8u * (2 * S + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (2 * S + 2 + a); };
  const int quad = warp & 3;
  const int sub = ((warp - 2) >> 2) & 1;        // two groups of four warps take alternate tiles
  const int r = quad * 32 + lane;               // accumulator row = conv pixel (r / 8, r % 8) of the tile
  uint8_t* tb = gbase + p.off_out + (uint32_t)sub * kStageBuf;   // this group's tile: 128 rows x 128 B, XOR-swizzled
  const float* s_bias = reinterpret_cast<const float*>(gbase + p.off_bias);
  const uint32_t bar_id = 1u + (uint32_t)sub;
  for (int ti = 0;; ++ti) {
    const long long tile = (long long)t_first + (long long)ti * t_stride;
    if (tile >= p.num_tiles) break;
    if ((ti & 1) != sub) continue;   // the other group's tile (only the owning group waits for / releases an accumulator)
    mbar_wait(tfull_bar(ti & 1), (uint32_t)((ti >> 1) & 1));
    tc_fence_after();
    const TileCoord t = decode_tile<false>(p, (int)tile);
    const uint32_t t_acc = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)((ti & 1) * p.acc_stride);
    named_bar_sync(bar_id, 128);   // the group's pooling reads of its previous tile are done
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      float v[32];
      tmem_ld_x16(t_acc + (uint32_t)(32 * hf), &v[0]);
      tmem_ld_x16(t_acc + (uint32_t)(32 * hf + 16), &v[16]);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 4; ++j)
        *reinterpret_cast<uint4*>(tb + sw128_off((uint32_t)r, (uint32_t)(4 * hf + j))) =
            epilogue8<EQXV_ACT_RELU, 0>(&v[8 * j], s_bias + 32 * hf + 8 * j, make_uint4(0u, 0u, 0u, 0u));
    }
    tc_fence_before();
    mbar_arrive(tempty_bar(ti & 1));
    named_bar_sync(bar_id, 128);   // the tile is complete in shared memory
    const int ph0 = t.h0 >> 1, pw0 = t.w0 >> 1;
    for (int it = r; it < 45 * 8; it += 128) {
      const int j = it & 7, pix = it >> 3;
      const int phi = pix / 5, pwi = pix - phi * 5;
      const int ph = ph0 + phi, pw = pw0 + pwi;
      if (ph >= p.pool_h || pw >= p.pool_w) continue;
      uint4 m = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
      for (int dr = -1; dr <= 1; ++dr) {
        const int hr = 2 * phi + dr;
        if (hr < 0 || hr > 15) continue;
#pragma unroll
        for (int dc = -1; dc <= 1; ++dc) {
          const int wc = 2 * pwi + dc;
          if (wc < 0 || wc > 7) continue;
          const uint32_t rr = (uint32_t)(hr * 8 + wc);
          m = bf16x8_max(m, *reinterpret_cast<const uint4*>(tb + sw128_off(rr, (uint32_t)j)));
        }
      }
      __nv_bfloat16* dst = p.pool_y + (((long long)t.n0 * p.pool_h + ph) * p.pool_w + pw) * p.pool_pitch + 8 * j;
      if (phi >= 1 && phi <= 7 && pwi >= 1 && pwi <= 3) {
        *reinterpret_cast<uint4*>(dst) = m;
      } else {
        red_max_bf16x8(dst, m);
      }
    }
  }
}

constexpr int kStemThreads16 = kThreads16 + 32 * (kStemProducers - 1);   // sixteen epilogue warps (see epilogue_warps16)
template <int kAct, bool kPool = false, bool kEpi16 = false>
__global__ void __launch_bounds__(kEpi16 ? kStemThreads16 : kStemThreads, 1) stem_kernel(const __grid_constant__ IgemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - raw);
  const int warp = threadIdx.x >> 5;
  const int S = p.stages;
  const uint32_t bars = base + p.off_bars;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (S + s); };
  auto tfull_bar = [&](int a) { return bars + 8u * (2 * S + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (2 * S + 2 + a); };
  const uint32_t tmem_slot = bars + 8u * (2 * S + 12);
  const uint32_t wfull_bar = bars + 8u * (2 * S + 14);
  volatile uint32_t* tmem_slot_g =
      reinterpret_cast<volatile uint32_t*>(gbase + p.off_bars + 8 * (2 * S + 12));

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.tmB);
    if (!kPool) tma_prefetch_desc(&p.tmC);
    for (int s = 0; s < S; ++s) {
      mbar_init(full_bar(s), p.h_tma ? 1u : 32u * kStemProducers);  // gather: every producer lane arrives once its copies landed
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 128u * (uint32_t)(p.epi_sub == 4 ? 4 : min(p.epi_sub, (p.block_n + 63) / 64)));   // the warps that own a chunk of the tile (sixteen-warp epilogue: all)
    }
    mbar_init(wfull_bar, 1);
    mbar_fence_init();
  }
  {
    float* sb = reinterpret_cast<float*>(gbase + p.off_bias);
    for (int i = threadIdx.x; i < p.block_n + 64; i += blockDim.x)
      sb[i] = (p.bias != nullptr && i < p.cout) ? __ldg(p.bias + i) : 0.f;
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    tmem_relinquish();
  }
  griddep_wait();     // PDL: everything above overlapped the previous kernel's tail
  tc_fence_before();
  __syncthreads();
  griddep_launch();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_g;
  const uint32_t b_smem = base + p.h_off_b;
  const uint32_t b_slab = (uint32_t)p.block_n * 128u;

  if (warp == 0) {
    // ============================== TMA producer ==============================
    if (elect_one()) {
      mbar_expect_tx(wfull_bar, (uint32_t)p.kh * b_slab);
      for (int r = 0; r < p.kh; ++r) tma_load_2d(b_smem + r * b_slab, &p.tmB, wfull_bar, r * kBlockK, 0);
    }
    __syncwarp();
    if (p.h_tma) {
      // Pixel-pair layout: the halo tile is a dense [h_rows x h_px units] window of the packed image - one TMA box per
      // tile (out-of-bounds = zero fill). The clock64 timeline of the gather (tools/stem_timeline.py) showed each of the
      // four gather warps busy ~1900 cycles per tile (~950 to get ten 16-byte LDGSTS per lane accepted, the rest in
      // group waits / fences / arrives): THAT was the tile period, whatever else was changed.
      if (elect_one()) {
        tma_prefetch_desc(&p.tmA);
        const uint32_t bytes = (uint32_t)(p.h_px * 16 * p.h_rows);
        int stage = 0, it = 0;
        uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
          const TileCoord t = decode_tile(p, tile);
          mbar_wait(empty_bar(stage), phase ^ 1u);
          IG_STAMP(it, 3);
          mbar_expect_tx(full_bar(stage), bytes);
          tma_load_4d(base + stage * p.h_stage_bytes, &p.tmA, full_bar(stage), t.w0 * p.h_stride * 8, t.h0 * p.h_stride_y, t.n0, 0);
          IG_STAMP(it, 4);
          if (++stage == S) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
      __syncwarp();
    } else {
      stem_gather(p, base, 0);
    }
  } else if (warp == 1 || (p.h_tma && warp == 2 + 4 * p.epi_sub)) {
    // ============================== MMA issuer(s) ==============================
    const uint32_t idesc = umma_idesc_bf16_m128((uint32_t)p.block_n);
    const uint64_t bdesc_hi = umma_desc_sw128(0);
    // A: SWIZZLE_NONE, K-major. LBO = byte distance tap 2j -> tap 2j+1, SBO = next output row.
    const uint32_t lbo = (p.h_stride == 1) ? 16u : (uint32_t)p.h_plane_pitch;
    const uint32_t sbo = (uint32_t)(p.h_stride_y * p.h_px * 16);
    const uint64_t adesc_hi = ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | ((uint64_t)1 << 46);
    // per-K-step start offsets (tap 2j: plane (2j % stride), pixel 2j / stride), hoisted out of the
    // issue loop: the single issuing thread must not spend cycles on integer division
    uint32_t joff[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
      joff[j] = (uint32_t)(((2 * j) % p.h_stride) * p.h_plane_pitch + ((2 * j) / p.h_stride) * 16) >> 4;
    const uint32_t row_step = (uint32_t)(p.h_px * 16) >> 4;
    const int ksteps = p.h_ksteps, kh = p.kh;
    // TMA mode: TWO issuing threads (this warp: even tiles / accumulator 0; the first idle gather warp: odd tiles /
    // accumulator 1). One thread spends ~500 cycles per tile between its last UMMA and the next tile's first (commits,
    // two barrier waits, loop) while the tensor pipe drains: with two, the other tile's UMMAs are already queued.
    const int it_step = p.h_tma ? 2 : 1;
    const int it0 = (p.h_tma && warp != 1) ? 1 : 0;
    if (elect_one()) {
      mbar_wait(wfull_bar, 0);
      const uint64_t bdesc0 = bdesc_hi | (uint64_t)((b_smem & 0x3FFFF) >> 4);
      for (int it = it0;; it += it_step) {
        const long long tile = (long long)blockIdx.x + (long long)it * gridDim.x;
        if (tile >= p.num_tiles) break;
        const int stage = it % S, acc = it & 1;
        const uint32_t phase = (uint32_t)((it / S) & 1), acc_phase = (uint32_t)((it >> 1) & 1);
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        IG_STAMP(it, 0);
        mbar_wait(full_bar(stage), phase);
        IG_STAMP(it, 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.acc_stride);
        const uint32_t a_src = base + stage * p.h_stage_bytes;
        uint64_t adesc = adesc_hi | (uint64_t)((a_src & 0x3FFFF) >> 4);
        uint64_t bdesc = bdesc0;
        if (p.h_stride == 1 && (ksteps == 2 || ksteps == 4)) {
          // Lean issue for the unit-stride layouts (joff[j] == 2 j, the same +2 steps as the filter): 32-bit descriptor
          // words, one asm block per 2 / 4 K steps. The generic loop below costs ~60 cycles per UMMA in the issuing
          // thread (tools/stem_timeline.py: 850 cycles for the 14 UMMAs of a ResNet stem tile against 672 of execution).
          uint32_t a_lo = (uint32_t)adesc, b_lo = (uint32_t)bdesc;
          const uint32_t a_hi = (uint32_t)(adesc_hi >> 32), b_hi = (uint32_t)(bdesc_hi >> 32);
          const uint32_t b_step = b_slab >> 4;
          if (ksteps == 2) {
            for (int r = 0; r < kh; ++r, a_lo += row_step, b_lo += b_step)
              umma_bf16_2steps_nc(d_tmem, a_lo, b_lo, a_hi, b_hi, idesc, (uint32_t)(r != 0));
          } else {
            for (int r = 0; r < kh; ++r, a_lo += row_step, b_lo += b_step)
              umma_bf16_kblock64_nc(d_tmem, a_lo, b_lo, a_hi, b_hi, idesc, (uint32_t)(r != 0));
          }
        } else
        for (int r = 0; r < kh; ++r) {
          umma_bf16(d_tmem, adesc + joff[0], bdesc, idesc, (uint32_t)(r != 0));
          if (ksteps > 1) umma_bf16(d_tmem, adesc + joff[1], bdesc + 2, idesc, 1u);
          if (ksteps > 2) umma_bf16(d_tmem, adesc + joff[2], bdesc + 4, idesc, 1u);
          if (ksteps > 3) umma_bf16(d_tmem, adesc + joff[3], bdesc + 6, idesc, 1u);
          adesc += row_step;
          bdesc += b_slab >> 4;
        }
        umma_commit(empty_bar(stage));
        umma_commit(tfull_bar(acc));
        IG_STAMP(it, 2);
      }
    }
    __syncwarp();
  } else if (warp < 2 + 4 * p.epi_sub) {
    if constexpr (kPool) {
      stem_pool_epilogue(p, base, gbase, tmem_base, warp, threadIdx.x & 31);
    } else if constexpr (kEpi16) {
      epilogue_warps16<kAct, 0, false>(p, base, gbase, tmem_base, warp, threadIdx.x & 31);
    } else {
      epilogue_warps<false, kAct, 0>(p, base, gbase, tmem_base, warp, threadIdx.x & 31);
    }
  } else {
    if (!p.h_tma) stem_gather(p, base, warp - (1 + 4 * p.epi_sub));  // producer warps 1..3
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------
// Halo-tile convolution (kh x kw, stride 1, small dilation) for large feature maps.
//
// The generic kernel moves the A tile once per filter tap (9x for a 3x3), which on the 56x56..224x224
// layers is bound by the L2->SM path, not by HBM or the tensor pipe (ResNet-50 layer1 3x3: 1.5 TB/s of
// HBM, 19% tensor). Here ONE 4-D TMA box (64 ch, tw + (kw-1)*dil, th + (kh-1)*dil, 1 image) per 64-deep
// channel block serves all taps: tap (r, s) is the same 128B-swizzled tile read through a descriptor
// whose start address is shifted by ((r*dil)*box_w + s*dil) rows and whose 8-row-group stride (SBO) is
// box_w rows, i.e. one output row (tw = 8). The UMMA swizzle is a function of the shared-memory
// address bits, exactly like the TMA write pattern, so an unaligned start row is legal with
// base_offset = 0 (verified on B200 by csrc/debug_umma.cu: gpurun_out/umma_shift.log).
// B slabs ([block_n x 64] per tap) run through their own ring.
template <int kAct, int kRes>
__global__ void __launch_bounds__(kThreads, 1) halo_kernel(const __grid_constant__ IgemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - raw);
  const int warp = threadIdx.x >> 5;
  const int S = p.stages;        // B ring depth (barrier layout shared with the generic kernel)
  const int SA = p.h_planes;     // A ring depth
  const uint32_t bars = base + p.off_bars;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (S + s); };
  auto tfull_bar = [&](int a) { return bars + 8u * (2 * S + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (2 * S + 2 + a); };
  const uint32_t tmem_slot = bars + 8u * (2 * S + 12);
  auto afull_bar = [&](int s) { return bars + 8u * (2 * S + 14 + s); };
  auto aempty_bar = [&](int s) { return bars + 8u * (2 * S + 14 + SA + s); };
  volatile uint32_t* tmem_slot_g =
      reinterpret_cast<volatile uint32_t*>(gbase + p.off_bars + 8 * (2 * S + 12));

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB);
    tma_prefetch_desc(&p.tmC);
    if (p.has_res) tma_prefetch_desc(&p.tmR);
    for (int s = 0; s < S; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < SA; ++s) {
      mbar_init(afull_bar(s), 1);
      mbar_init(aempty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 128u * (uint32_t)(p.epi_sub == 4 ? 4 : min(p.epi_sub, (p.block_n + 63) / 64)));   // the warps that own a chunk of the tile (sixteen-warp epilogue: all)
    }
    for (int b = 0; b < 16; ++b) mbar_init(bars + 8u * (48 + b), 1);   // residual ring (epilogue_warps)
    mbar_fence_init();
  }
  {
    const int ncols_pad = p.n_tiles * p.block_n + 64;
    stage_columns(reinterpret_cast<float*>(gbase + p.off_bias), p.bias, p.cout, ncols_pad);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    tmem_relinquish();
  }
  griddep_wait();     // PDL: everything above overlapped the previous kernel's tail
  tc_fence_before();
  __syncthreads();
  griddep_launch();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_g;
  const uint32_t a_smem = base + p.h_off_b;          // A ring follows the B ring
  const uint32_t b_slab = (uint32_t)p.block_n * 128u;
  const uint32_t a_bytes = (uint32_t)(p.h_px * p.h_rows * 128);
  const int taps = p.kh * p.kw;

  if (warp == 0) {
    // ============================== TMA producer (one elected thread) ==============================
    if (elect_one()) {
      int sa = 0;
      uint32_t pa = 0, pb = 0;
      uint32_t bdst = base, fb = full_bar(0), eb = empty_bar(0);
      const uint32_t bdst_end = base + (uint32_t)S * b_slab, fb0 = fb, eb0 = eb;
      const int kchunks = p.kchunks;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const TileCoord t = decode_tile(p, tile);
        for (int c = 0; c < kchunks; ++c) {
          mbar_wait(aempty_bar(sa), pa ^ 1u);
          mbar_expect_tx(afull_bar(sa), a_bytes);
          tma_load_4d(a_smem + sa * p.h_stage_bytes, &p.tmA, afull_bar(sa),
                      (c == kchunks - 1 && p.a_tail >= 0) ? p.a_tail : c * kBlockK, t.w0 - p.pad_w,
                      t.h0 - p.pad_h, t.n0);
          if (++sa == SA) {
            sa = 0;
            pa ^= 1u;
          }
          int kb = c * kBlockK;
          for (int tap = 0; tap < taps; ++tap, kb += p.cin_pack) {
            if (!p.b_resident || tile == (int)blockIdx.x) {   // resident filter: one pass over the ring, ever
              mbar_wait(eb, pb ^ 1u);
              mbar_expect_tx(fb, b_slab);
              tma_load_2d(bdst, &p.tmB, fb, kb, t.ncol0);
            }
            bdst += b_slab, fb += 8, eb += 8;
            if (bdst == bdst_end) {
              bdst = base, fb = fb0, eb = eb0;
              pb ^= 1u;
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ============================== MMA issuer (one elected thread) ==============================
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_bf16_m128((uint32_t)p.block_n);
      const uint32_t b_hi = (uint32_t)(umma_desc_sw128(0) >> 32);
      // A: 128B swizzle, K-major, 8-row groups strided by one halo row (SBO = box_w * 128 B)
      const uint32_t a_hi = ((uint32_t)(p.h_px * 128) >> 4) | (1u << 14) | (2u << 29);
      const uint32_t lbo = 1u << 16;
      const uint32_t bstep = b_slab >> 4;
      const uint32_t b_lo0 = ((base & 0x3FFFF) >> 4) | lbo, b_end = b_lo0 + (uint32_t)S * bstep;
      uint32_t b_lo = b_lo0, fb = full_bar(0), eb = empty_bar(0);
      const uint32_t fb0 = fb, eb0 = eb;
      const uint32_t row_step = (uint32_t)(p.dil_h * p.h_px) * 8u, col_step = (uint32_t)p.dil_w * 8u;
      const int kh = p.kh, kw = p.kw, kchunks = p.kchunks;
      int sa = 0;
      uint32_t pa = 0, pb = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      if (p.b_resident && kh == 3 && kw == 3 && kchunks == 1) {
        // Lean path for the resident 3x3 filter (ResNet layer1, VGG, DenseNet 64-channel layers): the generic
        // loop below spends ~60 SASS instructions per tap on ring bookkeeping, barrier waits and
        // register->uniform-register moves against 128 tensor-pipe cycles of work (4 MMAs of N = 64): the
        // issuing thread, not the tensor pipe, bounded the kernel (ncu r01s12: tensor pipe 33 % active).
        // Here the nine taps are straight-line code: constant filter descriptors, no waits, no commits.
        for (int sl = 0; sl < 9; ++sl) mbar_wait(full_bar(sl), 0u);   // the filter has landed (once per CTA)
        tc_fence_after();
        const uint32_t stage_step = (uint32_t)p.h_stage_bytes >> 4;
        const uint32_t a_base0 = ((a_smem & 0x3FFFF) >> 4) | lbo;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
          mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
          mbar_wait(afull_bar(sa), pa);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.acc_stride);
          const uint32_t a0 = a_base0 + (uint32_t)sa * stage_step;
#pragma unroll
          for (int r = 0; r < 3; ++r) {
#pragma unroll
            for (int s2 = 0; s2 < 3; ++s2) {
              umma_bf16_kblock64_nc(d_tmem, a0 + (uint32_t)r * row_step + (uint32_t)s2 * col_step,
                                    b_lo0 + (uint32_t)(r * 3 + s2) * bstep, a_hi, b_hi, idesc,
                                    (r | s2) != 0 ? 1u : 0u);
            }
          }
          umma_commit(aempty_bar(sa));
          umma_commit(tfull_bar(acc));
          if (++sa == SA) {
            sa = 0;
            pa ^= 1u;
          }
          acc ^= 1;
          if (acc == 0) acc_phase ^= 1u;
        }
      } else
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.acc_stride);
        uint32_t accumulate = 0;
        for (int c = 0; c < kchunks; ++c) {
          mbar_wait(afull_bar(sa), pa);
          uint32_t a_row = (((a_smem + sa * p.h_stage_bytes) & 0x3FFFF) >> 4) | lbo;
          for (int r = 0; r < kh; ++r, a_row += row_step) {
            uint32_t a_lo = a_row;
            for (int s2 = 0; s2 < kw; ++s2, a_lo += col_step) {
              mbar_wait(fb, p.b_resident ? 0u : pb);   // resident slabs: phase 0 completed once and for all
              tc_fence_after();
              umma_bf16_kblock64(d_tmem, a_lo, b_lo, a_hi, b_hi, idesc, accumulate, eb);
              accumulate = 1;
              b_lo += bstep, fb += 8, eb += 8;
              if (b_lo == b_end) {
                b_lo = b_lo0, fb = fb0, eb = eb0;
                pb ^= 1u;
              }
            }
          }
          umma_commit(aempty_bar(sa));
          if (++sa == SA) {
            sa = 0;
            pa ^= 1u;
          }
        }
        umma_commit(tfull_bar(acc));
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
    __syncwarp();
  } else {
    epilogue_warps<false, kAct, kRes>(p, base, gbase, tmem_base, warp, threadIdx.x & 31);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
constexpr int kNumActs = 8;
using KernelFn = void (*)(const IgemmParams);
struct KernelTable {
  KernelFn bf16[3][kNumActs];  // [residual mode][activation]
  KernelFn f32[kNumActs];
};
#define EQXV_ACT_ROW(F32, RES)                                                                        \
  {                                                                                                   \
    igemm_kernel<F32, 0, RES>, igemm_kernel<F32, 1, RES>, igemm_kernel<F32, 2, RES>,                  \
        igemm_kernel<F32, 3, RES>, igemm_kernel<F32, 4, RES>, igemm_kernel<F32, 5, RES>,              \
        igemm_kernel<F32, 6, RES>, igemm_kernel<F32, 7, RES>                                          \
  }
static const KernelTable& kernel_table() {
  static const KernelTable t = {{EQXV_ACT_ROW(false, 0), EQXV_ACT_ROW(false, 1), EQXV_ACT_ROW(false, 2)},
                                EQXV_ACT_ROW(true, 0)};
  return t;
}

struct IgemmProblem {
  // A operand tensor map (4-D)
  TmapSpec a;
  // B operand: packed weights [cout, ktot]
  const void* wgt;
  int ktot;
  const float* bias;
  // output / residual: [n, ho, wo, pitch] viewed with the M-tile geometry
  void* y;
  const void* res;
  int y_pitch, res_pitch;
  int out_n, out_h, out_w;  // logical M grid (flat views use out_w = M, out_h = out_n = 1)
  int tw, th, tn;
  int cout;
  int kh, kw, dil_h, dil_w, pad_h, pad_w, mul_h, mul_w;
  int in_h, in_w;
  int kchunks, cin_pack;
  int cin;      // logical input channels (gate extent)
  int a_tail;   // see IgemmParams
  int act, flags;
  int grouped;
  // LayerNorm folding (flat GEMMs): 0 none, 1 consumer, 2 producer (see IgemmParams)
  int ln_mode;
  float2* ln_stats;
  const float* ln_wsum;
  int ln_slots;
  float ln_inv_d, ln_eps;
  // SE gate on the A operand (flat GEMMs)
  const void* gate;
  int gate_pitch, gate_rpi;
};

static void choose_tile(int n, int ho, int wo, int& tw, int& th, int& tn) {
  long long best = -1;
  for (int w = 128; w >= 1; w >>= 1) {
    for (int h = 128 / w; h >= 1; h >>= 1) {
      const int b = 128 / (w * h);
      const long long padded = (long long)ceil_div(wo, w) * w * (long long)ceil_div(ho, h) * h *
                               (long long)ceil_div(n, b) * b;
      if (best < 0 || padded < best) {  // ties keep the wider (more contiguous) tile found first
        best = padded;
        tw = w;
        th = h;
        tn = b;
      }
    }
  }
}

// N tile. A single n-tile may be any multiple of 16 (the 64-wide store boxes are clipped at the
// tensor edge). With several n-tiles the tile width must be a multiple of 64 so that no store box
// reaches into the neighbouring tile's columns.
// The width is chosen against the persistent schedule: the launch takes ceil(tiles / SMs) rounds and a
// round lasts as long as one tile, so a narrower tile that fills the last round beats a wide one that
// leaves most SMs idle in it (ViT-B: 99 m-tiles x N=768 is 2.007 rounds of 128x256 but 2.68 of 128x192;
// ResNet layer4: 98 m-tiles x N=512). The tile time model is fitted to B200 measurements.
static int choose_block_n(int cout, long long m_tiles, int kblocks, int sms) {
  const int single = std::max(32, ceil_div(cout, 16) * 16);
  int best_bn = 0;
  double best_cost = 0;
  auto consider = [&](int bn) {
    const long long nt = ceil_div(cout, bn);
    const long long rounds = (m_tiles * nt + sms - 1) / sms;
    // measured per-K-block time of the single-CTA kernel: ~245 + 2.6*bn cycles (908 @256, 575 @128)
    const double tile = (double)kblocks * (245.0 + 2.6 * bn) + 1000.0;
    const double cost = (double)rounds * tile;
    if (best_bn == 0 || cost < best_cost * 0.97) {  // prefer the wider tile unless clearly beaten
      best_bn = bn;
      best_cost = cost;
    }
  };
  if (cout <= 256) consider(single);
  for (int bn = 256; bn >= 64; bn -= 64)
    if (bn < cout) consider(bn);
  return best_bn;
}

#define EQXV_PAIR_ROW(RES)                                                                              \
  {                                                                                                     \
    pair_kernel<0, RES>, pair_kernel<1, RES>, pair_kernel<2, RES>, pair_kernel<3, RES>,                 \
        pair_kernel<4, RES>, pair_kernel<5, RES>, pair_kernel<6, RES>, pair_kernel<7, RES>              \
  }
static KernelFn pair_table(int act, int res_mode) {
  static const KernelFn t[3][kNumActs] = {EQXV_PAIR_ROW(0), EQXV_PAIR_ROW(1), EQXV_PAIR_ROW(2)};
  return t[res_mode][act];
}

// Tuning overrides, read at every launch (i.e. at plan-build / graph-capture time) so that a sweep can change
// them inside one process (tools/sweep_igemm.py): EQXV_NO_PAIR=1, EQXV_FORCE_PAIR=1, EQXV_BLOCK_N=<n>, EQXV_EPI_SUB=<1|2>.
static int env_int(const char* name) {
  const char* v = getenv(name);
  return v ? atoi(v) : 0;
}

// pair variant: the tile is 256 x bn for two SMs; per SM and K block: 16 KiB of A + bn*64 B of B
static int choose_block_n_pair(int cout, long long pair_m_tiles, int kblocks, int clusters) {
  int best_bn = 0;
  double best_cost = 0;
  for (int bn = 256; bn >= 64; bn -= 64) {
    if (bn > 64 && bn >= 2 * cout) continue;
    const long long nt = ceil_div(cout, bn);
    const long long rounds = (pair_m_tiles * nt + clusters - 1) / clusters;
    // measured on B200 (gpurun_out/ab_pair.log): ~190 + 2.2*bn cycles per K block (753 @256, 472 @128):
    // operand reads + TMA fill share the 128 B/cycle shared-memory port
    const double cost = (double)rounds * ((double)kblocks * (190.0 + 2.2 * bn) + 1000.0);
    if (best_bn == 0 || cost < best_cost * 0.97) {
      best_bn = bn;
      best_cost = cost;
    }
  }
  return best_bn;
}

static KernelFn pair16_table(int act, int res_mode) {   // act none / relu / silu, residual before / after the activation
  static const KernelFn t[2][3] = {{pair_kernel16<0, 1>, pair_kernel16<1, 1>, pair_kernel16<2, 1>},
                                   {pair_kernel16<0, 2>, pair_kernel16<1, 2>, pair_kernel16<2, 2>}};
  return (act >= 0 && act <= 2 && res_mode >= 1 && res_mode <= 2) ? t[res_mode - 1][act] : nullptr;
}
static KernelFn epi16_table(int act) {
  static const KernelFn t[kNumActs] = {
      igemm_kernel<false, 0, 0, 0, false, true>, igemm_kernel<false, 1, 0, 0, false, true>,
      igemm_kernel<false, 2, 0, 0, false, true>, igemm_kernel<false, 3, 0, 0, false, true>,
      igemm_kernel<false, 4, 0, 0, false, true>, igemm_kernel<false, 5, 0, 0, false, true>,
      igemm_kernel<false, 6, 0, 0, false, true>, igemm_kernel<false, 7, 0, 0, false, true>};
  return t[act];
}

static int launch_igemm(const IgemmProblem& q, cudaStream_t stream) {
  IgemmParams p;
  memset(&p, 0, sizeof(p));
  const bool out_f32 = (q.flags & EQXV_FLAG_OUT_F32) != 0;
  EQXV_CHECK_ARG(!(out_f32 && q.res), "igemm: residual is not supported with fp32 output");
  const long long m_tiles =
      (long long)ceil_div(q.out_w, q.tw) * ceil_div(q.out_h, q.th) * ceil_div(q.out_n, q.tn);
  const int kblocks = q.kh * q.kw * q.kchunks;
  // CTA pairs pay off when the K loop (not HBM or the epilogue) dominates: deep K, wide N, enough tiles
  const bool pair_ok = !out_f32 && q.cout >= 128 && m_tiles >= 2 && q.dil_h == 1 && !q.grouped;
  const bool pair = pair_ok && !q.gate && !env_int("EQXV_NO_PAIR") && (kblocks >= 4 || env_int("EQXV_FORCE_PAIR"));
  int block_n;
  if (pair) {
    block_n = choose_block_n_pair(q.cout, (m_tiles + 1) / 2, kblocks, device_sm_count() / 2);
  } else {
    block_n = choose_block_n(q.cout, m_tiles, kblocks, device_sm_count());
  }
  {
    const int bn = env_int("EQXV_BLOCK_N");   // multiple of 64 below cout (several n-tiles) -- sweeps only
    if (bn >= 64 && bn <= 256 && bn % 64 == 0 && bn < q.cout && !(pair && bn >= 2 * q.cout)) block_n = bn;
  }
  if (q.grouped) block_n = 64;   // one n-tile = one 64-channel block of the block-diagonal filter
  if (q.ln_mode == 2 && block_n % 64 != 0) block_n = std::min(256, ceil_div(block_n, 64) * 64);   // 64-column stat chunks
  p.grouped = q.grouped;
  p.block_n = block_n;
  p.acc_stride = ceil_div(block_n, 32) * 32;
  int cols = 32;
  while (cols < 2 * p.acc_stride) cols <<= 1;
  p.tmem_cols = cols;
  p.n_tiles = ceil_div(q.cout, block_n);
  p.tw = q.tw, p.th = q.th, p.tn = q.tn;
  p.tiles_w = ceil_div(q.out_w, q.tw);
  p.tiles_h = ceil_div(q.out_h, q.th);
  p.tiles_n = ceil_div(q.out_n, q.tn);
  const long long num_tiles = (pair ? (m_tiles + 1) / 2 : m_tiles) * p.n_tiles;
  EQXV_CHECK_ARG(num_tiles > 0 && num_tiles < (1ll << 30), "igemm: bad tile count %lld", num_tiles);
  p.num_tiles = (int)num_tiles;
  p.kh = q.kh, p.kw = q.kw, p.dil_h = q.dil_h, p.dil_w = q.dil_w;
  p.pad_h = q.pad_h, p.pad_w = q.pad_w, p.mul_h = q.mul_h, p.mul_w = q.mul_w;
  p.in_h = q.in_h, p.in_w = q.in_w;
  p.kchunks = q.kchunks, p.cin_pack = q.cin_pack;
  p.a_tail = q.a_tail;
  p.cout = q.cout;
  p.act = q.act;
  p.res_after_act = (q.flags & EQXV_FLAG_RES_AFTER_ACT) ? 1 : 0;
  p.has_res = q.res ? 1 : 0;
  p.bias = q.bias;

  // shared memory carve-up
  const int stage_bytes = kABytes + (pair ? block_n * 64 : block_n * 128);
  const int bias_one = ceil_div((p.n_tiles * block_n + 64) * 4, 1024) * 1024;
  const int bias_bytes = bias_one * (q.ln_mode == 1 ? 2 : 1);   // consumer: + the filter's column sums
  EQXV_CHECK_ARG(bias_bytes <= (q.ln_mode == 1 ? 28 : 20) * 1024, "igemm: cout %d too large for the bias staging area",
                 q.cout);
  p.ln_stats = q.ln_stats, p.ln_wsum = q.ln_wsum, p.ln_slots = q.ln_slots, p.ln_rows = q.out_w;
  p.ln_inv_d = q.ln_inv_d, p.ln_eps = q.ln_eps;
  // direct epilogue: flat bf16 GEMMs whose K loop is too short to hide per-slab TMA operations
  {
    static const int direct_env = getenv("EQXV_DIRECT_OUT") ? atoi(getenv("EQXV_DIRECT_OUT")) : -1;
    const bool flat = q.tw == 128 && q.th == 1 && q.tn == 1 && q.kh == 1 && q.kw == 1 && !q.grouped;
    const bool ok = flat && !out_f32 && q.ln_mode == 0 && q.cout % 8 == 0 && q.y_pitch % 8 == 0 &&
                    (!q.res || q.res_pitch % 8 == 0);
    // Only where a warp's 32 rows are one contiguous run in memory (cout <= 32: <= 64 bytes per row): measured on B200,
    // wider rows turn every 16-byte store of a warp into 32 scattered sectors and the direct path LOSES to the TMA slab
    // (ResNet-50 layer1 c1/c3, K = 64: 3.33 -> 4.13 ms per step; EfficientNet 24 -> 144: 303 -> 389 us), while the
    // 24-channel projections gain (192 -> 159 us, 107 -> 97 us; what remains there is the TMA fetching 48-byte rows).
    p.direct_out = ok && (direct_env >= 0 ? direct_env != 0 : (kblocks <= 3 && q.cout <= 32)) ? 1 : 0;
    p.y_ptr = static_cast<__nv_bfloat16*>(q.y), p.res_ptr = static_cast<const __nv_bfloat16*>(q.res);
    p.y_pitch = q.y_pitch, p.res_pitch = q.res_pitch;
  }
  if (q.gate) {
    EQXV_CHECK_ARG(q.tw == 128 && q.th == 1 && q.tn == 1 && q.kh == 1 && q.kw == 1 && !out_f32 && !q.grouped &&
                       q.ln_mode == 0 && q.act == EQXV_ACT_NONE && !(q.flags & EQXV_FLAG_RES_AFTER_ACT),
                   "igemm: the gated A operand needs a flat bf16 GEMM without activation");
    p.gate = static_cast<const __nv_bfloat16*>(q.gate), p.gate_pitch = q.gate_pitch, p.gate_rpi = q.gate_rpi;
    p.gate_cols = q.cin;
    p.ln_rows = q.out_w;
  }
  if (q.ln_mode != 0) {
    EQXV_CHECK_ARG(q.tw == 128 && q.th == 1 && q.tn == 1 && q.kh == 1 && q.kw == 1 && !out_f32 && !q.grouped,
                   "igemm: LayerNorm folding needs a flat bf16 GEMM");
    EQXV_CHECK_ARG(q.ln_mode == 1 ? (q.res == nullptr && (q.act == EQXV_ACT_NONE || q.act == EQXV_ACT_GELU_TANH))
                                  : (q.res != nullptr && q.act == EQXV_ACT_NONE &&
                                     !(q.flags & EQXV_FLAG_RES_AFTER_ACT)),
                   "igemm: LayerNorm folding: consumer = no residual, act none/gelu; producer = residual, no act");
  }
  // epilogue warps per TMEM lane quadrant; every warp owns 2 residual slabs, the 8 staging slabs are shared out
  const int forced_sub = env_int("EQXV_EPI_SUB");
  // Two warps per quadrant where the epilogue bounds the tile (shallow K, <= 4 K blocks: ResNet c3 / downsample
  // layers went from 78 % to 99 % of their HBM roofline); one where the K loop hides it: the leaner tile-major
  // loop wins there by 2-7 % with or without a residual (tools/sweep_igemm.py, profiles/r01_sweep_igemm_v18.txt).
  p.epi_sub = (forced_sub == 1 || forced_sub == 2) ? forced_sub : (kblocks <= 4 ? 2 : 1);
  if (q.ln_mode != 0) p.epi_sub = 1;   // the folding lives in the four-warp epilogue
  if (q.gate) p.epi_sub = 1;           // warps 6..9 scale the A tiles
  // sixteen warps where the tile is all epilogue: one or two K blocks, no residual, enough 64-column chunks per CTA
  // for the extra warps to matter (see epilogue_warps16). EQXV_EPI_SUB=4 forces it where it is legal, =1/2 disables it.
  bool epi16 = false;
  {
    const bool legal = !pair && !out_f32 && !q.res && q.ln_mode == 0 && !q.gate && !p.direct_out && !q.grouped &&
                       block_n % 16 == 0;
    const long long chunks = num_tiles * ceil_div(block_n, 64);
    if (legal && (forced_sub == 4 || (forced_sub == 0 && kblocks <= 2 && chunks >= 8ll * device_sm_count()))) {
      epi16 = true;
      p.epi_sub = 4;
    }
  }
  // ... and on the CTA-pair kernels with a residual (ResNet conv3 layers: 128 x 256 outputs + shortcut per CTA and tile
  // against a 4..8-block K loop): ncu on l4.c3 showed the tensor pipe 38 % active with a flat stall profile - the
  // four-warp epilogue (8000 cycles per tile) paced the 6000-cycle K loop, and eight warps cost a pipeline stage for
  // their residual ring. Sixteen warps with the residual landing in the staging slab need 64 KiB in all.
  bool pair16 = false;
  {
    static const bool no_pair16 = getenv("EQXV_NO_PAIR16") != nullptr;
    const int res_mode_ = q.res ? ((q.flags & EQXV_FLAG_RES_AFTER_ACT) ? 2 : 1) : 0;
    if (pair && q.res && q.ln_mode == 0 && !out_f32 && block_n % 16 == 0 && pair16_table(q.act, res_mode_) != nullptr &&
        !no_pair16 && forced_sub == 0 && kblocks >= 8) {   // measured: l4.c3 (K = 512) 39.6 -> 36.2 us, l3.c3 (K = 256) 47.0 -> 48.7
      pair16 = true;
      p.epi_sub = 4;
    }
  }
  // eight-warp epilogue: ONE staging slab per warp means every chunk waits until the TMA store of the previous chunk
  // has finished reading it; with a second slab (32 KiB more, one pipeline stage less) the store of chunk l overlaps the
  // math of chunk l+1. Worth it where the epilogue bounds the layer and the K loop is short (ResNet c3 layers).
  static const int obuf_env = getenv("EQXV_EPI_OBUFS") ? atoi(getenv("EQXV_EPI_OBUFS")) : 0;
  {
    static const bool no_gather = getenv("EQXV_NO_GATHER_A") != nullptr;
    const bool flat = q.tw == 128 && q.th == 1 && q.tn == 1 && q.kh == 1 && q.kw == 1;
    const long long a_pitch = (long long)(q.a.strides_bytes[0] / 2);
    if ((epi16 || q.gate) && flat && q.kchunks == 1 && q.cin_pack < 64 && q.cin_pack % 8 == 0 && a_pitch % 8 == 0 &&
        !no_gather && p.n_tiles == 1 && q.ktot % 8 == 0) {
      p.ga_on = 1;
      p.ga_ptr = static_cast<const __nv_bfloat16*>(q.a.base);
      p.gb_ptr = static_cast<const __nv_bfloat16*>(q.wgt);
      p.ga_pitch = (int)a_pitch, p.ga_rows = q.out_w, p.ga_k16 = q.cin_pack / 8, p.gb_pitch = q.ktot;
    }
  }
  int out_bytes = (epi16 || pair16) ? 4 * kStageBuf : 2 * kStageBuf;   // sixteen warps: one 4 KiB slab each
  p.epi_obufs = 0;
  // Measured on ResNet-50 (A/B in one call, 3.457 vs 3.43 ms per step): no gain - the c3 layers are not waiting on
  // their staging slab - so the second slab stays opt-in (EQXV_EPI_OBUFS=2).
  if (p.epi_sub == 2 && obuf_env == 2) {
    const int fixed2 = 4 * kStageBuf + (p.has_res ? 2 * p.epi_sub * kStageBuf : 0) + bias_bytes + 512;
    if ((kMaxSmem - 1024 - fixed2) / stage_bytes >= 3) {
      out_bytes = 4 * kStageBuf;
      p.epi_obufs = 2;
    }
  }
  const int bres_bytes = p.ga_on ? ceil_div(block_n * 128, 1024) * 1024 : 0;   // resident filter slab (gather mode)
  const int ring_stage = p.ga_on ? kABytes : stage_bytes;
  const int res_ring = (p.has_res && !pair16) ? 2 * p.epi_sub * kStageBuf : 0;   // pair16: the residual lands in the staging slab
  const int fixed = out_bytes + res_ring + bias_bytes + 512 + bres_bytes;
  int stages = (kMaxSmem - 1024 - fixed) / ring_stage;
  stages = std::min(stages, 8);
  EQXV_CHECK_ARG(stages >= 2, "igemm: not enough shared memory for block_n=%d", block_n);
  EQXV_CHECK_ARG(!p.ga_on || stages >= 4, "igemm: internal: gather mode needs four stages");
  p.stages = stages;
  p.off_bres = stages * ring_stage;
  p.off_out = p.off_bres + bres_bytes;
  p.off_res = p.off_out + out_bytes;
  p.off_bias = p.off_res + res_ring;
  p.off_wsum = p.off_bias + bias_one;
  p.off_bars = p.off_bias + bias_bytes;
  const int smem_bytes = p.off_bars + 512 + 1024;

  int rc = encode_tmap(&p.tmA, q.a);
  if (rc) return rc;
  TmapSpec b{};
  b.base = const_cast<void*>(q.wgt);
  b.dtype = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  b.rank = 2;
  b.dims[0] = (uint64_t)q.ktot, b.dims[1] = (uint64_t)q.cout;
  b.strides_bytes[0] = (uint64_t)q.ktot * 2;
  b.box[0] = kBlockK, b.box[1] = (uint32_t)(pair ? block_n / 2 : block_n);
  b.estride[0] = b.estride[1] = 1;
  b.swizzle = CU_TENSOR_MAP_SWIZZLE_128B;
  rc = encode_tmap(&p.tmB, b);
  if (rc) return rc;

  auto make_out = [&](CUtensorMap* m, void* ptr, int pitch, bool f32) -> int {
    TmapSpec c{};
    const uint64_t es = f32 ? 4 : 2;
    c.base = ptr;
    c.dtype = f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    c.rank = 4;
    c.dims[0] = (uint64_t)q.cout, c.dims[1] = (uint64_t)q.out_w, c.dims[2] = (uint64_t)q.out_h,
    c.dims[3] = (uint64_t)q.out_n;
    c.strides_bytes[0] = (uint64_t)pitch * es;
    c.strides_bytes[1] = c.strides_bytes[0] * (uint64_t)q.out_w;
    c.strides_bytes[2] = c.strides_bytes[1] * (uint64_t)q.out_h;
    // one box = one epilogue warp's slab: 32 consecutive rows of the (tn, th, tw) tile
    const int bw = std::min(q.tw, 32), bh = std::min(q.th, 32 / bw), bn = 32 / (bw * bh);
    c.box[0] = f32 ? 32 : 64, c.box[1] = (uint32_t)bw, c.box[2] = (uint32_t)bh, c.box[3] = (uint32_t)bn;
    c.estride[0] = c.estride[1] = c.estride[2] = c.estride[3] = 1;
    c.swizzle = CU_TENSOR_MAP_SWIZZLE_128B;
    return encode_tmap(m, c);
  };
  rc = make_out(&p.tmC, q.y, q.y_pitch, out_f32);
  if (rc) return rc;
  if (q.res) {
    rc = make_out(&p.tmR, const_cast<void*>(q.res), q.res_pitch, false);
    if (rc) return rc;
  }

  EQXV_CHECK_ARG(q.act >= 0 && q.act < kNumActs, "igemm: unknown activation %d", q.act);
  const int res_mode = q.res ? (p.res_after_act ? 2 : 1) : 0;
  if (pair) {
    const int clusters = std::min(p.num_tiles, device_sm_count() / 2);
    KernelFn pfn = pair_table(q.act, res_mode);
    if (q.ln_mode == 1) pfn = q.act == EQXV_ACT_NONE ? pair_kernel<0, 0, 1> : pair_kernel<EQXV_ACT_GELU_TANH, 0, 1>;
    if (q.ln_mode == 2) pfn = pair_kernel<0, 1, 2>;
    if (pair16) pfn = pair16_table(q.act, res_mode);
    EQXV_CUDA(launch_kernel(pfn, dim3(2 * clusters), dim3(64 + 128 * p.epi_sub), (size_t)(smem_bytes), stream, p));
    EQXV_CUDA(cudaGetLastError());
    return EQXV_OK;
  }
  const int grid = std::min(p.num_tiles, device_sm_count());
  KernelFn fn = out_f32 ? kernel_table().f32[q.act] : kernel_table().bf16[res_mode][q.act];
  if (q.ln_mode == 1)
    fn = q.act == EQXV_ACT_NONE ? igemm_kernel<false, 0, 0, 1> : igemm_kernel<false, EQXV_ACT_GELU_TANH, 0, 1>;
  if (q.ln_mode == 2) fn = igemm_kernel<false, 0, 1, 2>;
  if (q.gate) fn = q.res ? igemm_kernel<false, 0, 1, 0, true> : igemm_kernel<false, 0, 0, 0, true>;
  if (epi16) fn = epi16_table(q.act);
  // gate warps: two groups of four (alternate K blocks) unless the rows are gathered by them (one K block per tile)
  static const int gate_groups_env = getenv("EQXV_GATE_GROUPS") ? atoi(getenv("EQXV_GATE_GROUPS")) : 0;
  const int gate_groups = !q.gate ? 0 : ((p.ga_on || gate_groups_env == 1 || kblocks < 2) ? 1 : 2);
  EQXV_CUDA(launch_kernel(fn, dim3(grid), dim3(64 + 128 * p.epi_sub + 128 * gate_groups), (size_t)(smem_bytes), stream, p));
  EQXV_CUDA(cudaGetLastError());
  return EQXV_OK;
}

static KernelFn ln_kernels(int i) {
  static const KernelFn t[8] = {pair_kernel<0, 0, 1>, pair_kernel<EQXV_ACT_GELU_TANH, 0, 1>, pair_kernel<0, 1, 2>,
                                igemm_kernel<false, 0, 0, 1>, igemm_kernel<false, EQXV_ACT_GELU_TANH, 0, 1>,
                                igemm_kernel<false, 0, 1, 2>, igemm_kernel<false, 0, 0, 0, true>,
                                igemm_kernel<false, 0, 1, 0, true>};
  return t[i];
}

// halo kernel instantiations: the activations that follow large-map 3x3 convolutions
static KernelFn halo_table(int act, int res_mode) {
  static const KernelFn t[3][3] = {{halo_kernel<0, 0>, halo_kernel<1, 0>, halo_kernel<2, 0>},
                                   {halo_kernel<0, 1>, halo_kernel<1, 1>, halo_kernel<2, 1>},
                                   {halo_kernel<0, 2>, halo_kernel<1, 2>, halo_kernel<2, 2>}};
  return (act >= 0 && act <= 2) ? t[res_mode][act] : nullptr;
}

using StemFn = void (*)(const IgemmParams);
static StemFn stem_table16(int act);
static StemFn stem_table(int act) {
  static const StemFn t[kNumActs] = {stem_kernel<0>, stem_kernel<1>, stem_kernel<2>, stem_kernel<3>,
                                     stem_kernel<4>, stem_kernel<5>, stem_kernel<6>, stem_kernel<7>};
  return t[act];
}

static StemFn stem_table16(int act) {
  static const StemFn t[kNumActs] = {stem_kernel<0, false, true>, stem_kernel<1, false, true>, stem_kernel<2, false, true>,
                                     stem_kernel<3, false, true>, stem_kernel<4, false, true>, stem_kernel<5, false, true>,
                                     stem_kernel<6, false, true>, stem_kernel<7, false, true>};
  return t[act];
}

int igemm_init() {
  for (int a = 0; a < kNumActs; ++a)
    EQXV_CUDA(cudaFuncSetAttribute(stem_table16(a), cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
  for (int a = 0; a < kNumActs; ++a)
    EQXV_CUDA(cudaFuncSetAttribute(epi16_table(a), cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
  for (int a = 0; a < 3; ++a)
    for (int r = 1; r <= 2; ++r)
      EQXV_CUDA(cudaFuncSetAttribute(pair16_table(a, r), cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
  for (int i = 0; i < 8; ++i)
    EQXV_CUDA(cudaFuncSetAttribute(ln_kernels(i), cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
  for (int a = 0; a < kNumActs; ++a)
    EQXV_CUDA(cudaFuncSetAttribute(stem_table(a), cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
  EQXV_CUDA(cudaFuncSetAttribute(stem_kernel<EQXV_ACT_RELU, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
  for (int a = 0; a < 3; ++a)
    for (int r = 0; r < 3; ++r)
      EQXV_CUDA(cudaFuncSetAttribute(halo_table(a, r), cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
  for (int a = 0; a < kNumActs; ++a)
    for (int r = 0; r < 3; ++r)
      EQXV_CUDA(cudaFuncSetAttribute(pair_table(a, r), cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
  for (int a = 0; a < kNumActs; ++a) {
    for (int r = 0; r < 3; ++r)
      EQXV_CUDA(cudaFuncSetAttribute(kernel_table().bf16[r][a], cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     kMaxSmem));
    EQXV_CUDA(cudaFuncSetAttribute(kernel_table().f32[a], cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   kMaxSmem));
  }
  return EQXV_OK;
}

// Halo-tile variant (see halo_kernel): worthwhile when the map tiles well into 8 x 16 output blocks.
static bool halo_eligible(const eqxv_conv_desc* d, int ho, int wo) {
  if (d->stride != 1 || d->dil > 2 || d->kh * d->kw < 4 || d->kh > 5 || d->kw > 5) return false;
  if (d->flags & EQXV_FLAG_OUT_F32) return false;
  if (d->act < 0 || d->act > 2) return false;
  const long long padded = (long long)ceil_div(wo, 8) * 8 * ceil_div(ho, 16) * 16;
  if (padded * 100 > (long long)wo * ho * 116) return false;           // <= 16 % wasted rows
  const int bw = 8 + (d->kw - 1) * d->dil, bh = 16 + (d->kh - 1) * d->dil;
  if (bw * bh * 128 > 48 * 1024) return false;
  return true;
}

static int launch_halo(const eqxv_conv_desc* d, int ho, int wo, cudaStream_t stream) {
  IgemmParams p;
  memset(&p, 0, sizeof(p));
  p.tw = 8, p.th = 16, p.tn = 1;
  p.tiles_w = ceil_div(wo, 8), p.tiles_h = ceil_div(ho, 16), p.tiles_n = d->n;
  const long long m_tiles = (long long)p.tiles_w * p.tiles_h * p.tiles_n;
  const int kchunks = ceil_div(d->cin, kBlockK);
  const int block_n = choose_block_n(d->cout, m_tiles, d->kh * d->kw * kchunks, device_sm_count());
  p.block_n = block_n;
  p.acc_stride = ceil_div(block_n, 32) * 32;
  int cols = 32;
  while (cols < 2 * p.acc_stride) cols <<= 1;
  p.tmem_cols = cols;
  p.n_tiles = ceil_div(d->cout, block_n);
  const long long num_tiles = m_tiles * p.n_tiles;
  EQXV_CHECK_ARG(num_tiles > 0 && num_tiles < (1ll << 30), "conv: bad tile count %lld", num_tiles);
  p.num_tiles = (int)num_tiles;
  p.kh = d->kh, p.kw = d->kw, p.dil_h = p.dil_w = d->dil, p.pad_h = p.pad_w = d->pad;
  p.mul_h = p.mul_w = 1;
  p.in_h = d->h, p.in_w = d->w;
  const bool tail = (d->flags & EQXV_FLAG_K_TAIL_SHIFT) != 0;
  const int cin_pack = tail ? kchunks * kBlockK : d->cin;
  p.kchunks = kchunks, p.cin_pack = cin_pack;
  p.a_tail = tail ? d->cin - kBlockK : -1;
  p.cout = d->cout, p.act = d->act, p.bias = d->bias;
  p.res_after_act = (d->flags & EQXV_FLAG_RES_AFTER_ACT) ? 1 : 0;
  p.has_res = d->residual ? 1 : 0;
  p.h_px = 8 + (d->kw - 1) * d->dil;     // halo box width (pixels = 128-byte rows)
  p.h_rows = 16 + (d->kh - 1) * d->dil;  // halo box height
  p.h_stage_bytes = ceil_div(p.h_px * p.h_rows * 128, 1024) * 1024;

  const int b_slab = block_n * 128;
  const int bias_bytes = ceil_div((p.n_tiles * block_n + 64) * 4, 1024) * 1024;
  EQXV_CHECK_ARG(bias_bytes <= 20 * 1024, "conv: cout %d too large for the bias staging area", d->cout);
  const int forced_sub = env_int("EQXV_EPI_SUB");
  // K >= 9 blocks: the MMA loop bounds the tile, except on the lean resident-filter path (N = 64: ~1150
  // tensor cycles per tile against a ~2000-cycle epilogue chain per warp), which gets two warps per quadrant
  const bool lean = d->kh == 3 && d->kw == 3 && kchunks == 1 && block_n <= 64 && p.n_tiles == 1;
  p.epi_sub = (forced_sub == 1 || forced_sub == 2) ? forced_sub : (lean ? 2 : 1);
  const int fixed = 2 * kStageBuf + (p.has_res ? 2 * p.epi_sub * kStageBuf : 0) + bias_bytes + 512;
  int sa = 3;
  int sb = (kMaxSmem - 1024 - fixed - sa * p.h_stage_bytes) / b_slab;
  if (sb < 4) {
    sa = 2;
    sb = (kMaxSmem - 1024 - fixed - sa * p.h_stage_bytes) / b_slab;
  }
  // The B slabs ([block_n x 64] per tap and K chunk) used to stream through an 8-deep ring for EVERY tile
  // (ResNet layer1 3x3: 72 KiB of filter per 23 KiB of activation tile, L2->SM bound at 43 % of its
  // roofline, profiles/r01_layer_roofline_v3.txt). When the whole filter fits it is loaded once per CTA.
  const int slabs = d->kh * d->kw * kchunks;
  static const bool no_bres = getenv("EQXV_NO_BRES") != nullptr;
  if (p.n_tiles == 1 && slabs <= sb && slabs <= 12 && !no_bres) {
    sb = slabs;
    p.b_resident = 1;
  } else {
    sb = std::min(sb, 8);
  }
  EQXV_CHECK_ARG(sb >= 2, "conv: not enough shared memory for the halo pipeline");
  p.stages = sb;
  p.h_planes = sa;
  p.h_off_b = sb * b_slab;
  p.off_out = p.h_off_b + sa * p.h_stage_bytes;
  p.off_res = p.off_out + 2 * kStageBuf;
  p.off_bias = p.off_res + (p.has_res ? 2 * p.epi_sub * kStageBuf : 0);
  p.off_bars = p.off_bias + bias_bytes;
  const int smem_bytes = p.off_bars + 512 + 1024;

  TmapSpec a{};
  a.base = const_cast<void*>(d->x);
  a.dtype = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  a.rank = 4;
  a.swizzle = CU_TENSOR_MAP_SWIZZLE_128B;
  a.dims[0] = (uint64_t)d->cin, a.dims[1] = (uint64_t)d->w, a.dims[2] = (uint64_t)d->h, a.dims[3] = (uint64_t)d->n;
  a.strides_bytes[0] = (uint64_t)d->x_pitch * 2;
  a.strides_bytes[1] = a.strides_bytes[0] * (uint64_t)d->w;
  a.strides_bytes[2] = a.strides_bytes[1] * (uint64_t)d->h;
  a.box[0] = kBlockK, a.box[1] = (uint32_t)p.h_px, a.box[2] = (uint32_t)p.h_rows, a.box[3] = 1;
  a.estride[0] = a.estride[1] = a.estride[2] = a.estride[3] = 1;
  int rc = encode_tmap(&p.tmA, a);
  if (rc) return rc;
  TmapSpec b{};
  b.base = const_cast<void*>(d->wgt);
  b.dtype = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  b.rank = 2;
  const int ktot = d->kh * d->kw * cin_pack;
  b.dims[0] = (uint64_t)ktot, b.dims[1] = (uint64_t)d->cout;
  b.strides_bytes[0] = (uint64_t)ktot * 2;
  b.box[0] = kBlockK, b.box[1] = (uint32_t)block_n;
  b.estride[0] = b.estride[1] = 1;
  b.swizzle = CU_TENSOR_MAP_SWIZZLE_128B;
  rc = encode_tmap(&p.tmB, b);
  if (rc) return rc;
  auto make_out = [&](CUtensorMap* m, const void* ptr, int pitch) -> int {
    TmapSpec c{};
    c.base = const_cast<void*>(ptr);
    c.dtype = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    c.rank = 4;
    c.dims[0] = (uint64_t)d->cout, c.dims[1] = (uint64_t)wo, c.dims[2] = (uint64_t)ho, c.dims[3] = (uint64_t)d->n;
    c.strides_bytes[0] = (uint64_t)pitch * 2;
    c.strides_bytes[1] = c.strides_bytes[0] * (uint64_t)wo;
    c.strides_bytes[2] = c.strides_bytes[1] * (uint64_t)ho;
    c.box[0] = 64, c.box[1] = 8, c.box[2] = 4, c.box[3] = 1;
    c.estride[0] = c.estride[1] = c.estride[2] = c.estride[3] = 1;
    c.swizzle = CU_TENSOR_MAP_SWIZZLE_128B;
    return encode_tmap(m, c);
  };
  rc = make_out(&p.tmC, d->y, d->y_pitch);
  if (rc) return rc;
  if (d->residual) {
    rc = make_out(&p.tmR, d->residual, d->res_pitch);
    if (rc) return rc;
  }
  const int res_mode = d->residual ? (p.res_after_act ? 2 : 1) : 0;
  const int grid = std::min(p.num_tiles, device_sm_count());
  EQXV_CUDA(launch_kernel(halo_table(d->act, res_mode), dim3(grid), dim3(64 + 128 * p.epi_sub), (size_t)smem_bytes, stream, p));
  EQXV_CUDA(cudaGetLastError());
  return EQXV_OK;
}

}  // namespace eqxv

using namespace eqxv;

struct LnFold {
  int mode;   // 1 LayerNorm consumer, 2 LayerNorm producer, 0 with `gate`: SE gate on the A operand
  float2* stats;
  const float* wsum;
  int slots;
  float inv_d, eps;
  const void* gate = nullptr;
  int gate_pitch = 0, gate_rpi = 0;
};
static int conv_impl(const eqxv_conv_desc* d, const LnFold* ln, void* stream);
namespace eqxv {
bool gemv_applies(long long m, int n, int k);
int launch_gemv(const void* a, long long lda, const void* w, long long ldw, int w_head, int w_off, const float* bias,
                void* out, long long ldo, long long m, int n, int k, int act, bool out_f32, cudaStream_t stream);
}  // namespace eqxv

extern "C" int eqxv_conv2d_igemm_bf16(const eqxv_conv_desc* d, void* stream) { return conv_impl(d, nullptr, stream); }

static int conv_impl(const eqxv_conv_desc* d, const LnFold* ln, void* stream) {
  EQXV_CHECK_ARG(d != nullptr, "conv: null descriptor");
  EQXV_CHECK_ARG(d->x && d->wgt && d->y, "conv: null tensor pointer");
  EQXV_CHECK_ARG(d->n > 0 && d->h > 0 && d->w > 0 && d->cin > 0 && d->cout > 0, "conv: bad shape");
  EQXV_CHECK_ARG(d->kh >= 1 && d->kw >= 1 && d->stride >= 1 && d->dil >= 1 && d->pad >= 0,
                 "conv: bad kernel geometry");
  EQXV_CHECK_ARG(d->cin % 8 == 0, "conv: cin (%d) must be a multiple of 8 (pad the input)", d->cin);
  EQXV_CHECK_ARG(d->x_pitch % 8 == 0 && d->x_pitch >= d->cin, "conv: bad x_pitch %d", d->x_pitch);
  const bool f32 = (d->flags & EQXV_FLAG_OUT_F32) != 0;
  EQXV_CHECK_ARG(d->y_pitch >= d->cout && d->y_pitch % (f32 ? 4 : 8) == 0, "conv: bad y_pitch %d",
                 d->y_pitch);
  EQXV_CHECK_ARG(((uintptr_t)d->x & 15) == 0 && ((uintptr_t)d->y & 15) == 0 &&
                     ((uintptr_t)d->wgt & 15) == 0,
                 "conv: pointers must be 16-byte aligned");
  if (d->residual) {
    EQXV_CHECK_ARG(d->res_pitch % 8 == 0 && d->res_pitch >= d->cout, "conv: bad res_pitch");
    EQXV_CHECK_ARG(((uintptr_t)d->residual & 15) == 0, "conv: residual must be 16-byte aligned");
  }
  const int ho = (d->h + 2 * d->pad - d->dil * (d->kh - 1) - 1) / d->stride + 1;
  const int wo = (d->w + 2 * d->pad - d->dil * (d->kw - 1) - 1) / d->stride + 1;
  EQXV_CHECK_ARG(ho > 0 && wo > 0, "conv: empty output");
  const bool grouped = (d->flags & EQXV_FLAG_GROUPED_BLOCK64) != 0;
  if (grouped)
    EQXV_CHECK_ARG(d->cin == d->cout && d->cin % 64 == 0 && !f32,
                   "conv: GROUPED_BLOCK64 needs cin == cout, a multiple of 64, bf16 output");
  const bool tail = (d->flags & EQXV_FLAG_K_TAIL_SHIFT) != 0;
  if (tail)
    EQXV_CHECK_ARG(!grouped && d->cin > kBlockK && d->cin % kBlockK != 0,
                   "conv: K_TAIL_SHIFT needs a dense filter with cin > 64, cin %% 64 != 0 (cin = %d)", d->cin);
  if (!grouped && !ln && halo_eligible(d, ho, wo)) return launch_halo(d, ho, wo, (cudaStream_t)stream);

  IgemmProblem q{};
  if (ln) {
    q.ln_mode = ln->mode, q.ln_stats = ln->stats, q.ln_wsum = ln->wsum, q.ln_slots = ln->slots;
    q.ln_inv_d = ln->inv_d, q.ln_eps = ln->eps;
    q.gate = ln->gate, q.gate_pitch = ln->gate_pitch, q.gate_rpi = ln->gate_rpi;
  }
  q.grouped = grouped ? 1 : 0;
  q.wgt = d->wgt;
  const int cin_pack = grouped ? 64 : (tail ? ceil_div(d->cin, kBlockK) * kBlockK : d->cin);
  q.ktot = d->kh * d->kw * cin_pack;
  q.cin = d->cin;
  q.a_tail = tail ? d->cin - kBlockK : -1;
  q.bias = d->bias;
  q.y = d->y;
  q.res = d->residual;
  q.y_pitch = d->y_pitch;
  q.res_pitch = d->res_pitch;
  q.cout = d->cout;
  q.act = d->act;
  q.flags = d->flags;
  q.kchunks = grouped ? 1 : ceil_div(d->cin, kBlockK);
  q.cin_pack = cin_pack;
  q.a.base = const_cast<void*>(d->x);
  q.a.dtype = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  q.a.rank = 4;
  q.a.swizzle = CU_TENSOR_MAP_SWIZZLE_128B;

  const bool pointwise = (d->kh == 1 && d->kw == 1 && d->stride == 1 && d->pad == 0);
  if (pointwise && !ln && !grouped && !d->residual) {
    // ONE row (single-image classifier heads): one warp per output column, csrc/gemv.cu
    const long long m_rows = (long long)d->n * d->h * d->w;
    if (gemv_applies(m_rows, d->cout, d->cin)) {
      const int kc = ceil_div(d->cin, kBlockK);
      return launch_gemv(d->x, d->x_pitch, d->wgt, tail ? kc * kBlockK : d->cin, kBlockK * (kc - 1),
                         tail ? kc * kBlockK - d->cin : 0, d->bias, d->y, d->y_pitch, m_rows, d->cout, d->cin, d->act, f32,
                         (cudaStream_t)stream);
    }
  }
  if (pointwise) {
    // plain GEMM over the flattened pixel index
    const long long m = (long long)d->n * d->h * d->w;
    EQXV_CHECK_ARG(m < (1ll << 31), "conv: too many pixels");
    q.out_w = (int)m, q.out_h = 1, q.out_n = 1;
    q.tw = 128, q.th = 1, q.tn = 1;
    q.kh = q.kw = 1, q.dil_h = q.dil_w = 1, q.pad_h = q.pad_w = 0, q.mul_h = q.mul_w = 1;
    q.in_h = 1, q.in_w = (int)m;
    q.a.dims[0] = (uint64_t)d->cin, q.a.dims[1] = (uint64_t)m, q.a.dims[2] = 1, q.a.dims[3] = 1;
    q.a.strides_bytes[0] = (uint64_t)d->x_pitch * 2;
    q.a.strides_bytes[1] = q.a.strides_bytes[0] * (uint64_t)m;
    q.a.strides_bytes[2] = q.a.strides_bytes[1];
    q.a.box[0] = kBlockK, q.a.box[1] = 128, q.a.box[2] = 1, q.a.box[3] = 1;
    q.a.estride[0] = q.a.estride[1] = q.a.estride[2] = q.a.estride[3] = 1;
  } else {
    q.out_w = wo, q.out_h = ho, q.out_n = d->n;
    choose_tile(d->n, ho, wo, q.tw, q.th, q.tn);
    EQXV_CHECK_ARG(q.tw * d->stride <= 256 && q.th * d->stride <= 256, "conv: stride too large");
    q.kh = d->kh, q.kw = d->kw, q.dil_h = q.dil_w = d->dil, q.pad_h = q.pad_w = d->pad;
    q.mul_h = q.mul_w = d->stride;
    q.in_h = d->h, q.in_w = d->w;
    q.a.dims[0] = (uint64_t)d->cin, q.a.dims[1] = (uint64_t)d->w, q.a.dims[2] = (uint64_t)d->h,
    q.a.dims[3] = (uint64_t)d->n;
    q.a.strides_bytes[0] = (uint64_t)d->x_pitch * 2;
    q.a.strides_bytes[1] = q.a.strides_bytes[0] * (uint64_t)d->w;
    q.a.strides_bytes[2] = q.a.strides_bytes[1] * (uint64_t)d->h;
    q.a.box[0] = kBlockK;
    q.a.box[1] = (uint32_t)(q.tw * d->stride);
    q.a.box[2] = (uint32_t)(q.th * d->stride);
    q.a.box[3] = (uint32_t)q.tn;
    q.a.estride[0] = 1, q.a.estride[1] = (uint32_t)d->stride, q.a.estride[2] = (uint32_t)d->stride,
    q.a.estride[3] = 1;
  }
  return launch_igemm(q, (cudaStream_t)stream);
}

extern "C" int eqxv_gemm_bias_act_res_bf16(const void* a, int64_t lda, const void* w,
                                           const float* bias, const void* residual, int64_t ldr,
                                           void* out, int64_t ldo, int64_t m, int32_t n, int32_t k,
                                           int32_t act, int32_t flags, void* stream) {
  EQXV_CHECK_ARG(m > 0 && m < (1ll << 31) && n > 0 && k > 0, "gemm: bad shape");
  EQXV_CHECK_ARG(k % 8 == 0, "gemm: k (%d) must be a multiple of 8", k);
  eqxv_conv_desc d{};
  d.x = a, d.wgt = w, d.bias = bias, d.residual = residual, d.y = out;
  d.n = 1, d.h = 1, d.w = (int32_t)m, d.cin = k, d.cout = n;
  d.kh = d.kw = 1, d.stride = 1, d.pad = 0, d.dil = 1;
  d.x_pitch = (int32_t)lda, d.y_pitch = (int32_t)ldo, d.res_pitch = (int32_t)ldr;
  d.act = act, d.flags = flags;
  return eqxv_conv2d_igemm_bf16(&d, stream);
}

static void gemm_desc(eqxv_conv_desc& d, const void* a, int64_t lda, const void* w, const float* bias,
                      const void* residual, int64_t ldr, void* out, int64_t ldo, int64_t m, int32_t n, int32_t k,
                      int32_t act) {
  d.x = a, d.wgt = w, d.bias = bias, d.residual = residual, d.y = out;
  d.n = 1, d.h = 1, d.w = (int32_t)m, d.cin = k, d.cout = n;
  d.kh = d.kw = 1, d.stride = 1, d.pad = 0, d.dil = 1;
  d.x_pitch = (int32_t)lda, d.y_pitch = (int32_t)ldo, d.res_pitch = (int32_t)ldr;
  d.act = act, d.flags = 0;
}

extern "C" int eqxv_gemm_res_rowstats_bf16(const void* a, int64_t lda, const void* w, const float* bias,
                                           const void* residual, int64_t ldr, void* out, int64_t ldo,
                                           float* row_stats, int64_t m, int32_t n, int32_t k, void* stream) {
  EQXV_CHECK_ARG(m > 0 && m < (1ll << 31) && n > 0 && k > 0 && k % 8 == 0, "gemm_res_rowstats: bad shape");
  EQXV_CHECK_ARG(residual && row_stats && ((uintptr_t)row_stats & 7) == 0, "gemm_res_rowstats: residual and row_stats are required");
  eqxv_conv_desc d{};
  gemm_desc(d, a, lda, w, bias, residual, ldr, out, ldo, m, n, k, EQXV_ACT_NONE);
  LnFold ln{2, reinterpret_cast<float2*>(row_stats), nullptr, (n + 63) / 64, 0.f, 0.f};
  return conv_impl(&d, &ln, stream);
}

extern "C" int eqxv_gemm_ln_act_bf16(const void* a, int64_t lda, const void* w, const float* bias, const float* wsum,
                                     const float* row_stats, int32_t slots, float eps, void* out, int64_t ldo,
                                     int64_t m, int32_t n, int32_t k, int32_t act, void* stream) {
  EQXV_CHECK_ARG(m > 0 && m < (1ll << 31) && n > 0 && k > 0 && k % 8 == 0, "gemm_ln_act: bad shape");
  EQXV_CHECK_ARG(bias && wsum && row_stats && slots == (k + 63) / 64,
                 "gemm_ln_act: bias, wsum and row_stats[m][ceil(k/64)] are required (slots %d, k %d)", slots, k);
  eqxv_conv_desc d{};
  gemm_desc(d, a, lda, w, bias, nullptr, 0, out, ldo, m, n, k, act);
  LnFold ln{1, reinterpret_cast<float2*>(const_cast<float*>(row_stats)), wsum, slots, 1.f / (float)k, eps};
  return conv_impl(&d, &ln, stream);
}

extern "C" int eqxv_gemm_gated_bf16(const void* a, int64_t lda, const void* gate, int64_t ldg, int32_t rows_per_image,
                                    const void* w, const float* bias, const void* residual, int64_t ldr, void* out,
                                    int64_t ldo, int64_t m, int32_t n, int32_t k, int32_t flags, void* stream) {
  EQXV_CHECK_ARG(m > 0 && m < (1ll << 31) && n > 0 && k > 0 && k % 8 == 0, "gemm_gated: bad shape");
  EQXV_CHECK_ARG((flags & ~EQXV_FLAG_K_TAIL_SHIFT) == 0, "gemm_gated: only EQXV_FLAG_K_TAIL_SHIFT is accepted");
  EQXV_CHECK_ARG(gate && rows_per_image > 0 && ldg >= k && ldg % 8 == 0 && ((uintptr_t)gate & 15) == 0,
                 "gemm_gated: gate must be a 16-byte aligned bf16 [images, >= k] matrix");
  eqxv_conv_desc d{};
  gemm_desc(d, a, lda, w, bias, residual, ldr, out, ldo, m, n, k, EQXV_ACT_NONE);
  d.flags = flags;
  LnFold ln{0, nullptr, nullptr, 0, 0.f, 0.f, gate, (int)ldg, rows_per_image};
  return conv_impl(&d, &ln, stream);
}

static int conv_stem_impl(const void* xpad, const void* wgt, const float* bias, void* y, int32_t n, int32_t h, int32_t w,
                          int32_t cout, int32_t kh, int32_t kw, int32_t stride, int32_t pad, int32_t y_pitch, int32_t act,
                          bool pool, void* stream, bool c4 = false);
static long long* g_stem_dbg = nullptr;
extern "C" int eqxv_debug_stem_timeline(long long* ts) {
  g_stem_dbg = ts;
  return EQXV_OK;
}

extern "C" int eqxv_conv_stem_c4_bf16(const void* xpad4, const void* wgt, const float* bias, void* y, int32_t n, int32_t h,
                                      int32_t w, int32_t cout, int32_t kh, int32_t kw, int32_t stride, int32_t pad,
                                      int32_t y_pitch, int32_t act, void* stream) {
  return conv_stem_impl(xpad4, wgt, bias, y, n, h, w, cout, kh, kw, stride, pad, y_pitch, act, false, stream, true);
}

extern "C" int eqxv_conv_stem_bf16(const void* xpad, const void* wgt, const float* bias, void* y, int32_t n,
                                   int32_t h, int32_t w, int32_t cout, int32_t kh, int32_t kw, int32_t stride,
                                   int32_t pad, int32_t y_pitch, int32_t act, void* stream) {
  return conv_stem_impl(xpad, wgt, bias, y, n, h, w, cout, kh, kw, stride, pad, y_pitch, act, false, stream);
}

extern "C" int eqxv_conv_stem_maxpool_bf16(const void* xpad, const void* wgt, const float* bias, void* y_pooled, int32_t n,
                                           int32_t h, int32_t w, int32_t cout, int32_t kh, int32_t kw, int32_t stride,
                                           int32_t pad, int32_t y_pitch, void* stream) {
  return conv_stem_impl(xpad, wgt, bias, y_pooled, n, h, w, cout, kh, kw, stride, pad, y_pitch, EQXV_ACT_RELU, true, stream);
}

static int conv_stem_impl(const void* xpad, const void* wgt, const float* bias, void* y, int32_t n, int32_t h, int32_t w,
                          int32_t cout, int32_t kh, int32_t kw, int32_t stride, int32_t pad, int32_t y_pitch, int32_t act,
                          bool pool, void* stream, bool c4) {
  EQXV_CHECK_ARG(xpad && wgt && y, "stem: null pointer");
  if (c4)
    EQXV_CHECK_ARG(stride == 2 && w % 2 == 0 && cout <= 256 && !pool,
                   "stem(c4): the pixel-pair layout is for stride-2 first layers on even widths (stride %d, w %d)", stride, w);
  EQXV_CHECK_ARG(n > 0 && h > 0 && w > 0 && cout > 0, "stem: bad shape");
  EQXV_CHECK_ARG(kh >= 1 && kh <= 8 && kw >= 1 && kw <= 8 && stride >= 1 && stride <= 4 && pad >= 0 &&
                     2 * pad <= kw,
                 "stem: unsupported geometry k=%dx%d stride %d pad %d", kh, kw, stride, pad);
  EQXV_CHECK_ARG(y_pitch % 8 == 0 && y_pitch >= cout, "stem: bad y_pitch");
  const int ho = (h + 2 * pad - kh) / stride + 1, wo = (w + 2 * pad - kw) / stride + 1;
  EQXV_CHECK_ARG(ho > 0 && wo > 0, "stem: empty output");
  // layout written by eqxv_pack_stem_input; c4: [n, hp, (w + 8) / 2 pixel pairs, 2 x 4 channels] (eqxv_pack_stem_input_c4)
  const int hp = h + 2 * pad, wp = c4 ? (w + 8) / 2 : w + 8;
  EQXV_CHECK_ARG(act >= 0 && act < kNumActs, "stem: unknown activation %d", act);
  if (pool)
    EQXV_CHECK_ARG(cout == 64 && ho % 16 == 0 && wo % 8 == 0 && (stride == 1 || stride == 2 || stride == 4) && y_pitch % 8 == 0 &&
                       ((uintptr_t)y & 15) == 0,
                   "stem+maxpool: needs cout == 64 and a conv output of (16 a) x (8 b) pixels (got %d x %d x %d)", cout, ho, wo);
  if (cout <= 256 && (stride == 1 || stride == 2 || stride == 4)) {
    // ---- halo kernel: one staged image tile serves every filter tap ----
    IgemmParams p;
    memset(&p, 0, sizeof(p));
    const int block_n = std::max(32, ceil_div(cout, 16) * 16);
    p.block_n = block_n;
    p.acc_stride = ceil_div(block_n, 32) * 32;
    int cols = 32;
    while (cols < 2 * p.acc_stride) cols <<= 1;
    p.tmem_cols = cols;
    p.n_tiles = 1;
    p.tw = 8, p.th = 16, p.tn = 1;
    p.tiles_w = ceil_div(wo, p.tw), p.tiles_h = ceil_div(ho, p.th), p.tiles_n = n;
    const long long num_tiles = (long long)p.tiles_w * p.tiles_h * p.tiles_n;
    EQXV_CHECK_ARG(num_tiles < (1ll << 30), "stem: too many tiles");
    p.num_tiles = (int)num_tiles;
    p.kh = kh, p.kw = kw;
    p.cout = cout, p.act = act, p.bias = bias;
    // c4: a 16-byte unit is a PAIR of 4-channel pixels, so a stride-2 convolution walks the units with stride 1 (no
    // phase planes), a filter row of kw taps is ceil(kw / 2) unit taps = ceil(kw / 4) K steps - half the MMAs, half the
    // halo bytes and half the packed image of the 8-channel layout (ResNet stem: 28 -> 14 UMMAs per tile).
    const int sx = c4 ? 1 : stride;
    const int kw_u = c4 ? ceil_div(kw, 2) : kw;   // filter width in 16-byte units
    p.h_stride = sx, p.h_planes = sx, p.h_stride_y = stride;
    p.h_ksteps = ceil_div(kw_u, 2);
    const int kw_pad = 2 * p.h_ksteps;
    p.h_px = p.tw + (kw_pad - 1) / sx;
    p.h_rows = (p.th - 1) * stride + kh;
    p.h_plane_pitch = ceil_div(p.h_rows * p.h_px * 16, 128) * 128;
    p.h_stage_bytes = ceil_div(p.h_plane_pitch * p.h_planes, 1024) * 1024;
    const int b_bytes = kh * block_n * 128;
    const int bias_bytes = ceil_div((block_n + 64) * 4, 1024) * 1024;
    // EQXV_STEM_SUB: 1 / 2 / 4 epilogue warps per TMEM lane quadrant (default 2; 4 measured slower, see stem_gather)
    static const int stem_sub = getenv("EQXV_STEM_SUB") ? atoi(getenv("EQXV_STEM_SUB")) : 0;
    const bool epi16 = !pool && stem_sub == 4 && block_n % 16 == 0;   // measured: 182 vs 170 us with eight warps - opt-in only
    const int out_bytes = epi16 ? 4 * kStageBuf : 2 * kStageBuf;
    int stages = (kMaxSmem - 1024 - b_bytes - out_bytes - bias_bytes - 512) / p.h_stage_bytes;
    stages = std::min(stages, c4 ? 10 : 6);
    EQXV_CHECK_ARG(stages >= 2, "stem: not enough shared memory (k=%dx%d cout=%d)", kh, kw, cout);
    p.stages = stages;
    p.h_off_b = stages * p.h_stage_bytes;
    p.off_out = p.h_off_b + b_bytes;
    p.off_res = p.off_out + out_bytes;
    // N = 64 is one chunk per tile and 28 MMAs (~900 tensor cycles): with one epilogue warp per quadrant its
    // ~2000-cycle chain per tile bounded the kernel (188 us against 98 us of HBM traffic); two warps per
    // quadrant take alternate tiles.
    p.epi_sub = epi16 ? 4 : ((stem_sub == 1 && !pool) ? 1 : 2);   // the pooling epilogue is written for two groups of four warps
    p.off_bias = p.off_res;
  p.off_bars = p.off_bias + bias_bytes;
    const int smem_bytes = p.off_bars + 512 + 1024;   // barriers: 2 S + 15 slots of 8 bytes (S <= 10)

    p.h_src = xpad;
    p.dbg = g_stem_dbg;
    p.in_h = hp, p.in_w = wp;
    static const bool no_stem_tma = getenv("EQXV_NO_STEM_TMA") != nullptr;
    if (c4 && !no_stem_tma && p.h_px * 8 <= 256 && p.h_rows <= 256) {
      TmapSpec a{};
      a.base = const_cast<void*>(xpad);
      a.dtype = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
      a.rank = 4;
      a.swizzle = CU_TENSOR_MAP_SWIZZLE_NONE;
      a.dims[0] = (uint64_t)wp * 8, a.dims[1] = (uint64_t)hp, a.dims[2] = (uint64_t)n, a.dims[3] = 1;
      a.strides_bytes[0] = (uint64_t)wp * 16;
      a.strides_bytes[1] = a.strides_bytes[0] * (uint64_t)hp;
      a.strides_bytes[2] = a.strides_bytes[1] * (uint64_t)n;
      a.box[0] = (uint32_t)(p.h_px * 8), a.box[1] = (uint32_t)p.h_rows, a.box[2] = 1, a.box[3] = 1;
      a.estride[0] = a.estride[1] = a.estride[2] = a.estride[3] = 1;
      int rc_a = encode_tmap(&p.tmA, a);
      if (rc_a) return rc_a;
      p.h_tma = 1;
    }
    EQXV_CHECK_ARG(stages >= 4, "stem: pipeline too shallow");
    int rc;
    TmapSpec b{};
    b.base = const_cast<void*>(wgt);
    b.dtype = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    b.rank = 2;
    b.dims[0] = (uint64_t)(kh * 64), b.dims[1] = (uint64_t)cout;
    b.strides_bytes[0] = (uint64_t)(kh * 64) * 2;
    b.box[0] = kBlockK, b.box[1] = (uint32_t)block_n;
    b.estride[0] = b.estride[1] = 1;
    b.swizzle = CU_TENSOR_MAP_SWIZZLE_128B;
    rc = encode_tmap(&p.tmB, b);
    if (rc) return rc;
    if (pool) {
      // pooled rows ph % 8 == 0 and columns pw % 4 == 0 collect red.max contributions from several tiles: zero them
      // (ReLU outputs are >= 0); everything else is written exactly once with plain stores
      const int ph = ho / 2, pw = wo / 2;
      p.pool_y = static_cast<__nv_bfloat16*>(y), p.pool_h = ph, p.pool_w = pw, p.pool_pitch = y_pitch;
      const size_t px = (size_t)y_pitch * 2;
      EQXV_CUDA(cudaMemset2DAsync(y, 8 * pw * px, 0, pw * px, (size_t)n * ph / 8, (cudaStream_t)stream));
      EQXV_CUDA(cudaMemset2DAsync(y, 4 * px, 0, (size_t)cout * 2, (size_t)n * ph * pw / 4, (cudaStream_t)stream));
      const int grid = std::min(p.num_tiles, device_sm_count());
      EQXV_CUDA(launch_kernel(stem_kernel<EQXV_ACT_RELU, true>, dim3(grid),
                              dim3(64 + 128 * p.epi_sub + 32 * (kStemProducers - 1)), (size_t)(smem_bytes),
                              (cudaStream_t)stream, p));
      EQXV_CUDA(cudaGetLastError());
      return EQXV_OK;
    }
    TmapSpec c{};
    c.base = y;
    c.dtype = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    c.rank = 4;
    c.dims[0] = (uint64_t)cout, c.dims[1] = (uint64_t)wo, c.dims[2] = (uint64_t)ho, c.dims[3] = (uint64_t)n;
    c.strides_bytes[0] = (uint64_t)y_pitch * 2;
    c.strides_bytes[1] = c.strides_bytes[0] * (uint64_t)wo;
    c.strides_bytes[2] = c.strides_bytes[1] * (uint64_t)ho;
    c.box[0] = 64, c.box[1] = 8, c.box[2] = 4, c.box[3] = 1;   // one epilogue warp's slab: 4 rows x 8 columns
    c.estride[0] = c.estride[1] = c.estride[2] = c.estride[3] = 1;
    c.swizzle = CU_TENSOR_MAP_SWIZZLE_128B;
    rc = encode_tmap(&p.tmC, c);
    if (rc) return rc;
    const int grid = std::min(p.num_tiles, device_sm_count());
    EQXV_CUDA(launch_kernel(epi16 ? stem_table16(act) : stem_table(act), dim3(grid),
                            dim3(64 + 128 * p.epi_sub + 32 * (kStemProducers - 1)), (size_t)(smem_bytes), (cudaStream_t)stream, p));
    EQXV_CUDA(cudaGetLastError());
    return EQXV_OK;
  }
  EQXV_CHECK_ARG(!pool, "stem+maxpool: unsupported geometry");
  IgemmProblem q{};
  q.wgt = wgt;
  q.ktot = kh * 64;
  q.bias = bias;
  q.y = y;
  q.res = nullptr;
  q.y_pitch = y_pitch;
  q.cout = cout;
  q.act = act;
  q.flags = 0;
  q.kchunks = 1;
  q.cin_pack = 64;
  q.out_w = wo, q.out_h = ho, q.out_n = n;
  choose_tile(n, ho, wo, q.tw, q.th, q.tn);
  EQXV_CHECK_ARG(q.th * stride <= 256, "stem: tile too tall");
  // kh taps = the filter rows; one "channel block" = an 8-pixel x 8-channel window (64 elements,
  // 128 B) starting at padded column stride*wo. Windows of neighbouring outputs overlap, so the "w"
  // dimension of the map is the OUTPUT column (stride = `stride` pixels); rows are traversed with
  // element stride `stride`. Filter columns >= kw and channels >= cin carry zero weights.
  q.kh = kh, q.kw = 1, q.dil_h = q.dil_w = 1, q.pad_h = q.pad_w = 0, q.mul_h = stride, q.mul_w = 1;
  q.in_h = hp, q.in_w = wo;
  q.a.base = const_cast<void*>(xpad);
  q.a.dtype = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  q.a.rank = 4;
  q.a.swizzle = CU_TENSOR_MAP_SWIZZLE_128B;
  q.a.dims[0] = 64, q.a.dims[1] = (uint64_t)wo, q.a.dims[2] = (uint64_t)hp, q.a.dims[3] = (uint64_t)n;
  q.a.strides_bytes[0] = (uint64_t)stride * 16;                // `stride` pixels x 8 ch x 2 B
  q.a.strides_bytes[1] = (uint64_t)wp * 16;                    // one padded row
  q.a.strides_bytes[2] = (uint64_t)wp * 16 * (uint64_t)hp;     // one padded image
  q.a.box[0] = 64, q.a.box[1] = (uint32_t)q.tw, q.a.box[2] = (uint32_t)(q.th * stride),
  q.a.box[3] = (uint32_t)q.tn;
  q.a.estride[0] = 1, q.a.estride[1] = 1, q.a.estride[2] = (uint32_t)stride, q.a.estride[3] = 1;
  return launch_igemm(q, (cudaStream_t)stream);
}
