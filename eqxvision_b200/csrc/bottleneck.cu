// ResNet bottleneck (64-channel trunk, stride 1) as ONE kernel on a CTA pair: resnet.py:144-162 composed over
// resnet.py:288-296 (layer1 of ResNet-50/101/152).
//
//   t1  --3x3 conv + BN + ReLU-->  t2  --1x1 conv + BN (+ identity | + downsample(x0)) + ReLU-->  y
//                                                        y  --next block's 1x1 conv + BN + ReLU-->  t1'
//
// Layer by layer these tensors cross HBM seven times per block (t2 written + read, y written, read as the next block's
// c1 input and again as its residual, the 256-channel downsample output written + read); at batch 256 that is 1.1 ms of
// a 3.4 ms ResNet-50 step, every layer already at its own HBM roofline (profiles/r01_layer_roofline_v20.txt). Here t2
// never leaves the SM, y is written once and consumed in place by the next block's c1, and the downsample of the first
// block is two more K blocks of the c3 accumulation.
//
// One CTA pair (cta_group::2, one TPC) owns two 8 x 16 pixel tiles; the three filters stay resident, split between the
// two CTAs along N (each CTA holds HALF of every filter: 36 + 16..32 + 16 KiB), which is what makes them fit next to
// the tiles. Per tile and CTA:
//   c2   9 taps x 4 UMMA (M 256, N 64, K 16) on ONE halo tile of t1 (10 x 18 pixels x 64 ch, a single 4-D TMA box; tap
//        (r, s) is the same 128B-swizzled tile through a descriptor shifted by r*10 + s rows, SBO = one halo row)
//        -> D2 (TMEM, 64 columns)
//   e2   epilogue warps: D2 + b2 -> ReLU -> bf16 -> the A operand of c3, written over the consumed halo tile
//   c3   4 UMMA (N 256) [+ 4 more on the x0 tile with the downsample filter] -> D3 (256 columns)
//   e3   D3 + b3 (+ residual, TMA-prefetched INTO the output tile) -> ReLU -> bf16 -> output tile (4 x [128 x 64] SW128
//        chunks) -> TMA store of y; the same tile is the A operand of
//   c1'  16 UMMA (N 64, K 256) -> D1 (64 columns)
//   e4   D1 + b1' -> ReLU -> bf16 -> t1' (16-byte global stores, 128 contiguous bytes per thread)
// The issuer runs c2 of tile i+1 behind c3 of tile i, so the tensor pipe works through e3.
// Hand-offs: TMA -> issuer and issuer -> epilogue by mbarrier (tcgen05.commit multicast to both CTAs); epilogue ->
// issuer (operand tiles written with st.shared in BOTH CTAs, each consumed by its own SM's tensor core under the leader's
// MMAs) by fence.proxy.async + a plain arrive on the leader's barrier. (An arrive.release.cluster here compiles to
// MEMBAR.ALL + ERRBAR per thread and cost 19 % of the kernel's stall samples, profiles/r02_ncu_bneck_v1.txt; the data
// never crosses SMs, only the notification does.)
#include <cstdlib>
#include <cstring>

#include "common.h"
#include "ptx.cuh"
#include "epilogue.cuh"

namespace eqxv {

constexpr int kBnThreads = 576;                          // warp 0 TMA loads, warp 1 MMA issuer, warps 2..17 epilogue
constexpr int kHaloW = 10, kHaloH = 18;                  // halo of an 8 x 16 tile under a 3x3 filter
constexpr uint32_t kT1Bytes = kHaloW * kHaloH * 128;     // 23040
constexpr uint32_t kT1Slot = 23552;                      // rounded up to 1 KiB (swizzle atoms stay aligned)
constexpr uint32_t kTile = 16384;                        // 128 rows x 128 B
constexpr uint32_t kW2Tap = 32 * 128;                    // this CTA's half (32 of 64 filter rows) of one tap
constexpr int kBnMaxSmem = 232448;

struct alignas(64) BneckParams {
  CUtensorMap tmT1, tmX0, tmW2, tmW3, tmW1, tmY, tmR;
  const float *b2, *b3, *b1n;
  __nv_bfloat16* next_out;
  int next_pitch;
  int tiles_w, tiles_h, num_mtiles, num_pairs;
  int n, h, w;
  int w3_chunks;   // 1, or 2 with the downsample filter concatenated along K
  uint32_t off_w3, off_w1, off_t1, off_x0, off_y, off_bias, off_bars;
  long long* dbg;   // eqxv_debug_bottleneck_timeline: clock64 stamps of CTA 0 (tools/bneck_timeline.py), normally null
};

struct BnTile {
  int w0, h0, n0;
};
__device__ __forceinline__ BnTile bn_decode(const BneckParams& p, int m) {
  BnTile t;
  t.w0 = (m % p.tiles_w) * 8;
  m /= p.tiles_w;
  t.h0 = (m % p.tiles_h) * 16;
  t.n0 = m / p.tiles_h;   // >= p.n for the phantom tile of an odd count: TMA zero-fills / clips, direct stores check
  return t;
}

#define BN_STAMP(it, ev)                                                              \
  do {                                                                                \
    if (p.dbg != nullptr && blockIdx.x == 0 && (it) < 16) p.dbg[(it) * 16 + (ev)] = clock64(); \
  } while (0)

// barrier slots (8 bytes each from off_bars)
enum : uint32_t {
  kBarW = 0, kBarT1Full = 1, kBarT1Empty = 3, kBarX0Full = 5, kBarX0Empty = 6, kBarD2Full = 7, kBarA3Full = 8,
  kBarD3Full = 9, kBarYFull = 10, kBarD1Full = 11, kBarTmemSlot = 12, kBarRes = 16   // + (warp - 2)
};

// 32 lanes x 32 contiguous bytes each: one full sector per thread in ONE instruction (16-byte stores leave half-written
// sectors in flight and double the L2 write transactions of a pattern where every lane hits a different 128-byte line)
__device__ __forceinline__ void st_global_32B(void* ptr, const uint4& a, const uint4& b) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(ptr), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w),
               "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w)
               : "memory");
}

template <bool kDown, int kNextN>   // kNextN: output channels of the fused next 1x1 convolution (0 = none, 64, 128)
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kBnThreads, 1)
    bneck_kernel(const __grid_constant__ BneckParams p) {
  constexpr bool kNext = kNextN != 0;
  constexpr uint32_t kW1Chunk = (uint32_t)(kNextN / 2) * 128u;   // this CTA's half of one 64-deep K chunk of w1'
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - raw);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = blockIdx.x & 1u;   // == %cluster_ctarank for a (2,1,1) cluster
  const uint32_t bars = base + p.off_bars;
  auto bar = [&](uint32_t i) { return bars + 8u * i; };
  const uint32_t w2_s = base, w3_s = base + p.off_w3, w1_s = base + p.off_w1, t1_s = base + p.off_t1,
                 x0_s = base + p.off_x0, y_s = base + p.off_y;
  volatile uint32_t* tmem_slot_g = reinterpret_cast<volatile uint32_t*>(gbase + p.off_bars + 8 * kBarTmemSlot);
  float* s_bias = reinterpret_cast<float*>(gbase + p.off_bias);   // [b2 64][b3 256][b1' kNextN]

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.tmT1);
    tma_prefetch_desc(&p.tmW2);
    tma_prefetch_desc(&p.tmW3);
    tma_prefetch_desc(&p.tmY);
    if (kDown) tma_prefetch_desc(&p.tmX0); else tma_prefetch_desc(&p.tmR);
    if (kNext) tma_prefetch_desc(&p.tmW1);
    mbar_init(bar(kBarW), 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar(kBarT1Full + s), 1);    // leader: its producer's arrive.expect_tx (both CTAs' bytes)
      mbar_init(bar(kBarT1Empty + s), 1);   // one multicast commit per use
    }
    mbar_init(bar(kBarX0Full), 1);
    mbar_init(bar(kBarX0Empty), 1);
    mbar_init(bar(kBarD2Full), 1);
    mbar_init(bar(kBarA3Full), 32);         // leader: the 2 x 16 epilogue warps of the pair
    mbar_init(bar(kBarD3Full), 1);
    mbar_init(bar(kBarYFull), 32);
    mbar_init(bar(kBarD1Full), 1);
    for (int b = 0; b < 16; ++b) mbar_init(bar(kBarRes + b), 1);
    mbar_fence_init();
    // The filters are constants of the plan (not written by the preceding kernel): fetched before the PDL wait.
    const uint32_t wbytes = 9u * kW2Tap + (uint32_t)p.w3_chunks * kTile + 4u * kW1Chunk;
    mbar_expect_tx(bar(kBarW), wbytes);
    for (int tap = 0; tap < 9; ++tap) tma_load_2d(w2_s + tap * kW2Tap, &p.tmW2, bar(kBarW), tap * 64, (int)rank * 32);
    for (int k = 0; k < p.w3_chunks; ++k) tma_load_2d(w3_s + k * kTile, &p.tmW3, bar(kBarW), k * 64, (int)rank * 128);
    if (kNext)
      for (int k = 0; k < 4; ++k) tma_load_2d(w1_s + k * kW1Chunk, &p.tmW1, bar(kBarW), k * 64, (int)rank * (kNextN / 2));
  }
  for (int i = threadIdx.x; i < 320 + kNextN; i += blockDim.x) {
    float v;
    if (i < 64) v = __ldg(p.b2 + i);
    else if (i < 320) v = __ldg(p.b3 + (i - 64));
    else v = __ldg(p.b1n + (i - 320));
    s_bias[i] = v;
  }
  if (warp == 1) {
    tmem_alloc_pair(bar(kBarTmemSlot), 512u);
    tmem_relinquish_pair();
  }
  griddep_wait();     // PDL: everything above overlapped the previous kernel's tail
  tc_fence_before();
  __syncthreads();
  griddep_launch();
  mbar_wait(bar(kBarW), 0u);   // this CTA's filter halves have landed
  cluster_sync_all();          // ... and the peer's; its barriers are initialised before anything is signalled remotely
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_g;
  // TMEM columns: D2 [0, 64), D1 [64, 64 + kNextN), D3 [256, 512). Y - the finished output tile as bf16 pairs, A operand of
  // c1' - is written IN PLACE over the D3 columns the same thread has just drained: chunk c (64 fp32 columns) leaves 32
  // packed columns at D3 + 64 c. c3 of the next tile overwrites them only after c1' (MMAs execute in issue order).
  const uint32_t d2_t = tmem_base, d1_t = tmem_base + 64u, d3_t = tmem_base + 256u;
  const int p_first = (int)(blockIdx.x >> 1), p_stride = (int)(gridDim.x >> 1);

  if (warp == 0) {
    // ============================== TMA producer (both CTAs, one thread) ==============================
    if (elect_one()) {
      int slot = 0;
      uint32_t ph = 0, xph = 0;
      for (int pi = p_first; pi < p.num_pairs; pi += p_stride) {
        const BnTile t = bn_decode(p, 2 * pi + (int)rank);
        mbar_wait(bar(kBarT1Empty + slot), ph ^ 1u);
        if (rank == 0) mbar_expect_tx(bar(kBarT1Full + slot), 2u * kT1Bytes);   // both CTAs' bytes land on this barrier
        tma_load_4d_pair(t1_s + slot * kT1Slot, &p.tmT1, bar(kBarT1Full + slot), 0, t.w0 - 1, t.h0 - 1, t.n0);
        if (!kDown) {
          // this tile's residual (16 slabs of 32 rows x 64 channels) on its way from HBM to L2: the producer runs one to
          // two tiles ahead of the epilogue warps, which then fetch their slabs from L2. (Issued from the epilogue warps
          // the prefetch instructions sat on their critical path: ~1500 cycles between two tiles.)
#pragma unroll
          for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int qq = 0; qq < 4; ++qq) tma_prefetch_l2_4d(&p.tmR, c * 64, t.w0, t.h0 + 4 * qq, t.n0);
        }
        if (kDown) {
          mbar_wait(bar(kBarX0Empty), xph ^ 1u);
          if (rank == 0) mbar_expect_tx(bar(kBarX0Full), 2u * kTile);
          tma_load_4d_pair(x0_s, &p.tmX0, bar(kBarX0Full), 0, t.w0, t.h0, t.n0);
          xph ^= 1u;
        }
        if (++slot == 2) {
          slot = 0;
          ph ^= 1u;
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ============================== MMA issuer (leader CTA, one thread) ==============================
    if (rank == 0 && elect_one()) {
      const uint32_t idesc64 = umma_idesc_bf16_m256(64u), idesc256 = umma_idesc_bf16_m256(256u);
      const uint32_t k_hi = (uint32_t)(umma_desc_sw128(0) >> 32);   // plain 128B-swizzled K-major tile
      // halo view: 8-row groups strided by one halo row (SBO = 10 x 128 B), see halo_kernel in igemm.cu
      const uint32_t halo_hi = ((uint32_t)(kHaloW * 128) >> 4) | (1u << 14) | (2u << 29);
      const uint32_t lbo = 1u << 16;
      auto lo = [&](uint32_t addr) { return ((addr & 0x3FFFFu) >> 4) | lbo; };
      const uint32_t w2_lo = lo(w2_s), w3_lo = lo(w3_s), w1_lo = lo(w1_s), x0_lo = lo(x0_s);
      auto issue_c2 = [&](int slot) {
        const uint32_t a0 = lo(t1_s + (uint32_t)slot * kT1Slot);
#pragma unroll
        for (int r = 0; r < 3; ++r) {
#pragma unroll
          for (int s = 0; s < 3; ++s) {
            umma_bf16_kblock64_pair_nc(d2_t, a0 + (uint32_t)(r * kHaloW + s) * 8u,
                                       w2_lo + (uint32_t)(r * 3 + s) * (kW2Tap >> 4), halo_hi, k_hi, idesc64,
                                       (r | s) != 0 ? 1u : 0u);
          }
        }
        umma_commit_pair(bar(kBarD2Full));
      };
      int slot = 0;
      uint32_t tph = 0;                 // parity of the once-per-tile barriers
      uint32_t fph[2] = {0u, 0u};       // parity of t1full[slot]
      int it = 0;
      for (int pi = p_first; pi < p.num_pairs; pi += p_stride, ++it) {
        if (it == 0) {
          mbar_wait(bar(kBarT1Full + 0), fph[0]);
          fph[0] ^= 1u;
          tc_fence_after();
          issue_c2(0);
        }
        // c3 (+ downsample): A = the bf16 tile the epilogue wrote over the consumed halo tile
        mbar_wait(bar(kBarA3Full), tph);
        if (kDown) mbar_wait(bar(kBarX0Full), tph);
        tc_fence_after();
        BN_STAMP(it, 0);
        umma_bf16_kblock64_pair_nc(d3_t, lo(t1_s + (uint32_t)slot * kT1Slot), w3_lo, k_hi, k_hi, idesc256, 0u);
        if (kDown) umma_bf16_kblock64_pair_nc(d3_t, x0_lo, w3_lo + (kTile >> 4), k_hi, k_hi, idesc256, 1u);
        umma_commit_pair(bar(kBarT1Empty + slot));
        if (kDown) umma_commit_pair(bar(kBarX0Empty));
        umma_commit_pair(bar(kBarD3Full));
        BN_STAMP(it, 1);
        // c2 of the next tile runs while the epilogue warps work through D3
        if (pi + p_stride < p.num_pairs) {
          const int ns = slot ^ 1;
          mbar_wait(bar(kBarT1Full + ns), fph[ns]);
          fph[ns] ^= 1u;
          tc_fence_after();
          issue_c2(ns);
        }
        BN_STAMP(it, 2);
        if (kNext) {
          // next block's c1: A = the finished output tile as bf16 pairs in TENSOR MEMORY (the epilogue wrote it there next
          // to the staging slab of the TMA store), so the slab is free for the next tile's residual as soon as the store
          // has read it - with A in shared memory every tile waited for c1' AND a TMA round trip before its e3.
          mbar_wait(bar(kBarYFull), tph);
          tc_fence_after();
          BN_STAMP(it, 3);
          const uint32_t idesc_n = umma_idesc_bf16_m256((uint32_t)kNextN);
#pragma unroll
          for (int k = 0; k < 16; ++k)
            umma_bf16_ts_pair(d1_t, d3_t + (uint32_t)(64 * (k >> 2) + 8 * (k & 3)),
                              w1_lo + (uint32_t)(k >> 2) * (kW1Chunk >> 4) + (uint32_t)(k & 3) * 2u, k_hi, idesc_n,
                              k != 0 ? 1u : 0u);
          umma_commit_pair(bar(kBarD1Full));
          BN_STAMP(it, 4);
        }
        slot ^= 1;
        tph ^= 1u;
      }
    }
    __syncwarp();
  } else {
    // ============================== epilogue (warps 2..17, both CTAs) ==============================
    // Four warps per TMEM lane quadrant: `sub` picks 16 of the 64 columns (e2, e4) / ONE 64-column chunk (e3).
    // Thread <-> accumulator row r = 32 q + lane <-> pixel (h0 + r / 8, w0 + r % 8). Every warp is autonomous: it owns
    // the 4 KiB slab (its 32 rows x its chunk) of the output tile - residual load, in-place arithmetic, TMA store - and
    // its own mbarrier; there is no CTA-wide synchronisation on this path.
    // Order per tile: e3(i), e2(i+1), residual load(i+1), e4(i): the issuer's c3(i+1) / c1'(i) run behind e2 / e4.
    const int q = warp & 3, sub = (warp - 2) >> 2;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const int r = q * 32 + lane;
    const int h_loc = r >> 3, w_loc = r & 7;
    const uint32_t rbar = bar(kBarRes + (uint32_t)(warp - 2));
    const uint32_t slab_off = (uint32_t)sub * kTile + (uint32_t)q * 4096u;   // 32 rows x 128 B of chunk `sub`
    uint8_t* const slab = gbase + p.off_y + slab_off;
    const float* const bias3 = s_bias + 64 + 64 * sub;
    auto e2 = [&](int slot, uint32_t ph) {   // D2 -> 16 columns of the A operand of c3, over the consumed halo tile
      mbar_wait(bar(kBarD2Full), ph);
      tc_fence_after();
      float v[16];
      tmem_ld_x16(d2_t + lane_base + (uint32_t)(16 * sub), v);
      tmem_ld_wait();
      tc_fence_before();
      uint8_t* a3 = gbase + p.off_t1 + (uint32_t)slot * kT1Slot;
#pragma unroll
      for (int j = 0; j < 2; ++j)
        *reinterpret_cast<uint4*>(a3 + sw128_off((uint32_t)r, (uint32_t)(2 * sub + j))) =
            epilogue8<EQXV_ACT_RELU, 0>(&v[8 * j], s_bias + 16 * sub + 8 * j, make_uint4(0u, 0u, 0u, 0u));
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(bar(kBarA3Full));
    };
    auto issue_res = [&](const BnTile& t) {   // ONE lane
      mbar_expect_tx(rbar, 4096u);
      tma_load_4d(y_s + slab_off, &p.tmR, rbar, sub * 64, t.w0, t.h0 + 4 * q, t.n0);
    };
    uint32_t tph = 0;
    int slot = 0, it = 0;
    const bool stamp = warp == 2 && lane == 0;
    BnTile t = bn_decode(p, 2 * p_first + (int)rank), tnext = t;
    if (p_first < p.num_pairs) {
      if (!kDown && lane == 0) issue_res(t);
      __syncwarp();
      e2(0, 0u);
    }
    for (int pi = p_first; pi < p.num_pairs; pi += p_stride, ++it) {
      const bool more = pi + p_stride < p.num_pairs;
      if (more) tnext = bn_decode(p, 2 * (pi + p_stride) + (int)rank);
      // ---------------- e3: D3 (+ residual, in place) -> y ----------------
      mbar_wait(bar(kBarD3Full), tph);
      tc_fence_after();
      if (stamp) BN_STAMP(it, 5);
      {
        // four steps of 16 columns, the next step's TMEM load in flight behind the current step's arithmetic
        float va[16], vb[16];
        tmem_ld_x16(d3_t + lane_base + (uint32_t)(64 * sub), va);
        if (!kDown) {
          mbar_wait(rbar, tph);   // the residual slab has landed
        } else {
          // no residual load orders this: the TMA store that last read the slab must be done before it is rewritten
          if (lane == 0) tma_store_wait_read<0>();
          __syncwarp();
        }
        if (stamp) BN_STAMP(it, 6);
#pragma unroll
        for (int st = 0; st < 4; ++st) {
          float* cur = (st & 1) ? vb : va;
          float* nxt = (st & 1) ? va : vb;
          uint4 r0 = make_uint4(0u, 0u, 0u, 0u), r1 = r0;
          if (!kDown) {
            r0 = *reinterpret_cast<const uint4*>(slab + sw128_off((uint32_t)lane, (uint32_t)(2 * st)));
            r1 = *reinterpret_cast<const uint4*>(slab + sw128_off((uint32_t)lane, (uint32_t)(2 * st + 1)));
          }
          tmem_ld_wait();
          if (st < 3) tmem_ld_x16(d3_t + lane_base + (uint32_t)(64 * sub + 16 * (st + 1)), nxt);
          uint4 o[2];
          o[0] = epilogue8<EQXV_ACT_RELU, kDown ? 0 : 1>(&cur[0], bias3 + 16 * st, r0);
          o[1] = epilogue8<EQXV_ACT_RELU, kDown ? 0 : 1>(&cur[8], bias3 + 16 * st + 8, r1);
          *reinterpret_cast<uint4*>(slab + sw128_off((uint32_t)lane, (uint32_t)(2 * st))) = o[0];
          *reinterpret_cast<uint4*>(slab + sw128_off((uint32_t)lane, (uint32_t)(2 * st + 1))) = o[1];
          // ... and, as bf16 pairs, into the A operand of c1' (row = lane, K = 64 sub + 16 st .. + 15)
          if (kNext) tmem_st_x8(d3_t + lane_base + (uint32_t)(64 * sub + 8 * st), reinterpret_cast<const uint32_t*>(o));
        }
        if (kNext) tmem_st_wait();
        tc_fence_before();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_4d(&p.tmY, y_s + slab_off, sub * 64, t.w0, t.h0 + 4 * q, t.n0);
          tma_store_commit();
          if (kNext) mbar_arrive_leader(bar(kBarYFull));
        }
        __syncwarp();
      }
      if (stamp) BN_STAMP(it, 7);
      // ---------------- e2 of the next tile (its c2 ran behind this tile's c3) ----------------
      if (more) e2(slot ^ 1, tph ^ 1u);
      if (stamp) BN_STAMP(it, 8);
      // ---------------- the next tile's residual into the slab this warp's store has just read ----------------
      if (!kDown && more) {
        if (lane == 0) {
          tma_store_wait_read<0>();
          issue_res(tnext);
        }
        __syncwarp();
      }
      if (stamp) BN_STAMP(it, 9);
      // ---------------- e4: D1 -> this thread's kNextN / 4 channels of t1' ----------------
      if constexpr (kNext) {
        constexpr int kCols = kNextN / 4;   // 16 or 32
        mbar_wait(bar(kBarD1Full), tph);
        tc_fence_after();
        float v[kCols];
#pragma unroll
        for (int g = 0; g < kCols / 16; ++g) tmem_ld_x16(d1_t + lane_base + (uint32_t)(kCols * sub + 16 * g), &v[16 * g]);
        tmem_ld_wait();
        tc_fence_before();
        const int hh = t.h0 + h_loc, ww = t.w0 + w_loc;
        if (t.n0 < p.n && hh < p.h && ww < p.w) {
          __nv_bfloat16* orow = p.next_out + (((long long)t.n0 * p.h + hh) * p.w + ww) * p.next_pitch + kCols * sub;
          const float* bn = s_bias + 320 + kCols * sub;
#pragma unroll
          for (int g = 0; g < kCols / 16; ++g) {
            const uint4 o0 = epilogue8<EQXV_ACT_RELU, 0>(&v[16 * g], bn + 16 * g, make_uint4(0u, 0u, 0u, 0u));
            const uint4 o1 = epilogue8<EQXV_ACT_RELU, 0>(&v[16 * g + 8], bn + 16 * g + 8, make_uint4(0u, 0u, 0u, 0u));
            st_global_32B(orow + 16 * g, o0, o1);
          }
        }
      }
      if (stamp) BN_STAMP(it, 10);
      t = tnext;
      slot ^= 1;
      tph ^= 1u;
    }
    if (lane == 0) tma_store_wait_all();
    __syncwarp();
  }

  // ---- teardown: neither CTA may leave while its peer can still touch its smem / TMEM / barriers ----
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512u);
  }
}

static long long* g_bn_dbg = nullptr;
using BneckFn = void (*)(const BneckParams);
static BneckFn bneck_table(bool down, int next_n) {
  static const BneckFn t[2][3] = {{bneck_kernel<false, 0>, bneck_kernel<false, 64>, bneck_kernel<false, 128>},
                                  {bneck_kernel<true, 0>, bneck_kernel<true, 64>, bneck_kernel<true, 128>}};
  return t[down ? 1 : 0][next_n / 64];
}

int bottleneck_init() {
  for (int d = 0; d < 2; ++d)
    for (int n = 0; n <= 128; n += 64)
      EQXV_CUDA(cudaFuncSetAttribute(bneck_table(d != 0, n), cudaFuncAttributeMaxDynamicSharedMemorySize, kBnMaxSmem));
  return EQXV_OK;
}

static int map4d(CUtensorMap* m, const void* ptr, int c, int w, int h, int n, int pitch, int bw, int bh) {
  TmapSpec s{};
  s.base = const_cast<void*>(ptr);
  s.dtype = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  s.rank = 4;
  s.swizzle = CU_TENSOR_MAP_SWIZZLE_128B;
  s.dims[0] = (uint64_t)c, s.dims[1] = (uint64_t)w, s.dims[2] = (uint64_t)h, s.dims[3] = (uint64_t)n;
  s.strides_bytes[0] = (uint64_t)pitch * 2;
  s.strides_bytes[1] = s.strides_bytes[0] * (uint64_t)w;
  s.strides_bytes[2] = s.strides_bytes[1] * (uint64_t)h;
  s.box[0] = 64, s.box[1] = (uint32_t)bw, s.box[2] = (uint32_t)bh, s.box[3] = 1;
  s.estride[0] = s.estride[1] = s.estride[2] = s.estride[3] = 1;
  return encode_tmap(m, s);
}
static int map2d(CUtensorMap* m, const void* ptr, int k, int rows, int box_rows) {
  TmapSpec s{};
  s.base = const_cast<void*>(ptr);
  s.dtype = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  s.rank = 2;
  s.swizzle = CU_TENSOR_MAP_SWIZZLE_128B;
  s.dims[0] = (uint64_t)k, s.dims[1] = (uint64_t)rows;
  s.strides_bytes[0] = (uint64_t)k * 2;
  s.box[0] = 64, s.box[1] = (uint32_t)box_rows;
  s.estride[0] = s.estride[1] = 1;
  return encode_tmap(m, s);
}

}  // namespace eqxv

using namespace eqxv;

extern "C" int eqxv_debug_bottleneck_timeline(long long* ts) {
  g_bn_dbg = ts;
  return EQXV_OK;
}

extern "C" int eqxv_bottleneck64_fused_bf16(const eqxv_bottleneck64_desc* d, void* stream) {
  EQXV_CHECK_ARG(d != nullptr, "bottleneck: null descriptor");
  EQXV_CHECK_ARG(d->t1 && d->w2 && d->b2 && d->w3 && d->b3 && d->y, "bottleneck: null tensor pointer");
  EQXV_CHECK_ARG((d->residual != nullptr) != (d->x0 != nullptr),
                 "bottleneck: exactly one of residual (identity shortcut) / x0 (downsample input) must be given");
  EQXV_CHECK_ARG((d->w1n != nullptr) == (d->next != nullptr) && (d->w1n != nullptr) == (d->b1n != nullptr),
                 "bottleneck: w1n, b1n and next go together");
  EQXV_CHECK_ARG(d->n > 0 && d->h > 0 && d->w > 0, "bottleneck: bad shape");
  const bool down = d->x0 != nullptr, next = d->next != nullptr;
  const int next_n = next ? d->next_channels : 0;
  EQXV_CHECK_ARG(!next || next_n == 64 || next_n == 128, "bottleneck: the fused next convolution has 64 or 128 outputs, not %d",
                 next_n);
  auto ok_ptr = [](const void* p_) { return ((uintptr_t)p_ & 15) == 0; };
  EQXV_CHECK_ARG(ok_ptr(d->t1) && ok_ptr(d->w2) && ok_ptr(d->w3) && ok_ptr(d->y) && ok_ptr(d->residual) &&
                     ok_ptr(d->x0) && ok_ptr(d->w1n) && ok_ptr(d->next),
                 "bottleneck: pointers must be 16-byte aligned");
  EQXV_CHECK_ARG(d->t1_pitch >= 64 && d->t1_pitch % 8 == 0 && d->y_pitch >= 256 && d->y_pitch % 8 == 0,
                 "bottleneck: bad t1 / y pitch");
  if (down) EQXV_CHECK_ARG(d->x0_pitch >= 64 && d->x0_pitch % 8 == 0, "bottleneck: bad x0 pitch");
  else EQXV_CHECK_ARG(d->res_pitch >= 256 && d->res_pitch % 8 == 0, "bottleneck: bad residual pitch");
  if (next) EQXV_CHECK_ARG(d->next_pitch >= next_n && d->next_pitch % 16 == 0 && ((uintptr_t)d->next & 31) == 0,
                           "bottleneck: next is written with 32-byte stores: pitch %% 16 == 0, 32-byte aligned base");

  BneckParams p;
  memset(&p, 0, sizeof(p));
  p.n = d->n, p.h = d->h, p.w = d->w;
  p.tiles_w = ceil_div(d->w, 8), p.tiles_h = ceil_div(d->h, 16);
  const long long mt = (long long)p.tiles_w * p.tiles_h * d->n;
  EQXV_CHECK_ARG(mt < (1ll << 30), "bottleneck: too many tiles");
  p.num_mtiles = (int)mt;
  p.num_pairs = (int)((mt + 1) / 2);
  p.b2 = d->b2, p.b3 = d->b3, p.b1n = d->b1n;
  p.next_out = static_cast<__nv_bfloat16*>(d->next), p.next_pitch = d->next_pitch;
  p.w3_chunks = down ? 2 : 1;
  p.dbg = g_bn_dbg;
  uint32_t off = 9u * kW2Tap;
  p.off_w3 = off, off += (uint32_t)p.w3_chunks * kTile;
  p.off_w1 = off, off += 4u * (uint32_t)(next_n / 2) * 128u;
  p.off_t1 = off, off += 2u * kT1Slot;
  p.off_x0 = off, off += down ? kTile : 0u;
  p.off_y = off, off += 4u * kTile;
  p.off_bias = off, off += 2048u;
  p.off_bars = off, off += 512u;
  const size_t smem_bytes = (size_t)off + 1024;
  EQXV_CHECK_ARG(smem_bytes <= (size_t)kBnMaxSmem, "bottleneck: shared memory budget exceeded (%zu)", smem_bytes);

  int rc = map4d(&p.tmT1, d->t1, 64, d->w, d->h, d->n, d->t1_pitch, kHaloW, kHaloH);
  if (rc) return rc;
  if (down) {
    rc = map4d(&p.tmX0, d->x0, 64, d->w, d->h, d->n, d->x0_pitch, 8, 16);
    if (rc) return rc;
  } else {
    rc = map4d(&p.tmR, d->residual, 256, d->w, d->h, d->n, d->res_pitch, 8, 4);
    if (rc) return rc;
  }
  rc = map4d(&p.tmY, d->y, 256, d->w, d->h, d->n, d->y_pitch, 8, 4);
  if (rc) return rc;
  rc = map2d(&p.tmW2, d->w2, 9 * 64, 64, 32);
  if (rc) return rc;
  rc = map2d(&p.tmW3, d->w3, down ? 128 : 64, 256, 128);
  if (rc) return rc;
  if (next) {
    rc = map2d(&p.tmW1, d->w1n, 256, next_n, next_n / 2);
    if (rc) return rc;
  }
  const int clusters = std::min(p.num_pairs, device_sm_count() / 2);
  EQXV_CUDA(launch_kernel(bneck_table(down, next_n), dim3(2 * clusters), dim3(kBnThreads), smem_bytes,
                          (cudaStream_t)stream, p));
  EQXV_CUDA(cudaGetLastError());
  return EQXV_OK;
}
