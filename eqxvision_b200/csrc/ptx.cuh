// Thin inline-PTX wrappers for the sm_100a features the kernels use:
// mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (MMA / TMEM alloc / TMEM load).
// Everything here is device-side and header-only.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace eqxv {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// one lane of a converged warp (deterministic for a given member mask)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred P;\n"
      "elect.sync _|P, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, P;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// ----------------------------------------------------------------------------------------------
// mbarrier
// ----------------------------------------------------------------------------------------------
// ---- programmatic dependent launch (PDL) --------------------------------------------------------
// Every kernel of the library is launched with cudaLaunchAttributeProgrammaticStreamSerialization
// (common.h: launch_kernel). griddep_wait() blocks until the preceding kernel of the stream has
// completed and its memory is visible; everything a kernel does before it (barrier init, TMEM
// allocation, tensor-map prefetch, constant loads) overlaps the predecessor's tail. EVERY kernel
// executes it (all threads, before the first access to activation memory): completion of kernel i
// then implies completion of kernel i-1, so dependencies further back than one launch stay ordered.
// griddep_launch() lets the successor's CTAs be scheduled as soon as this grid's CTAs are all resident.
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred P;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n"
      "selp.u32 %0, 1, 0, P;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Spin on a phase parity. A watchdog turns a protocol deadlock into a trap (the launch then
// fails with an error instead of hanging the device): ~4 s at 2 GHz. The slow path is kept out of
// line so that the many wait sites do not bloat the instruction stream.
static __device__ __noinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity) {
  const long long t0 = clock64();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (((++spins) & 0x3ff) == 0 && (clock64() - t0) > 8000000000LL) {
      printf("eqxv: mbarrier watchdog: block %d thread %d bar 0x%x parity %u\n", (int)blockIdx.x,
             (int)threadIdx.x, bar, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (!mbar_try_wait(bar, parity)) mbar_wait_slow(bar, parity);
}

// ----------------------------------------------------------------------------------------------
// TMA
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0,
                                            int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, "
      "%4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0,
                                            int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, "
      "%4, %5, %6}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// L2 prefetch of a 4-D box (no shared-memory destination, no barrier): a later TMA load of the same box hits L2
__device__ __forceinline__ void tma_prefetch_l2_4d(const CUtensorMap* m, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global [%0, {%1, %2, %3, %4}];" ::"l"(reinterpret_cast<uint64_t>(m)),
               "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, uint32_t src, int c0, int c1,
                                             int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
          reinterpret_cast<uint64_t>(m)),
      "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
          reinterpret_cast<uint64_t>(m)),
      "r"(src), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_store_commit() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void tma_store_wait_all() {
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
// Make generic-proxy smem writes visible to the async proxy (TMA store reads).
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ----------------------------------------------------------------------------------------------
// cp.async (LDGSTS): 16-byte gathers whose granularity is too fine for TMA boxes
// ----------------------------------------------------------------------------------------------
// copies 16 bytes, or writes 16 zero bytes when `valid` is false (src-size 0)
__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src, bool valid) {
  const uint32_t n = valid ? 16u : 0u;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// ----------------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem] * B[smem]; bf16 inputs, fp32 accumulate; issued by ONE thread.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// One 64-deep K block: 4 x (128 x N x 16) MMAs on 128B-swizzled K-major operands (+32 B per K step
// inside the swizzle row) followed by the commit that releases the smem stage. One asm block so that
// the single issuing thread spends its cycles on UTCHMMA, not on descriptor arithmetic. a_lo / b_lo are
// the low descriptor words ((addr & 0x3FFFF) >> 4); desc_hi the constant high word.
__device__ __forceinline__ void umma_bf16_kblock64(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo,
                                                   uint32_t a_hi, uint32_t b_hi, uint32_t idesc,
                                                   uint32_t accumulate, uint32_t commit_bar) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b64 da, db;\n"
      ".reg .b32 a1, b1;\n"
      "setp.ne.b32 p, %6, 0;\n"
      "mov.b64 da, {%1, %3};\n"
      "mov.b64 db, {%2, %4};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n"
      "setp.eq.b32 p, 0, 0;\n"
      "add.u32 a1, %1, 2;\n"
      "add.u32 b1, %2, 2;\n"
      "mov.b64 da, {a1, %3};\n"
      "mov.b64 db, {b1, %4};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n"
      "add.u32 a1, %1, 4;\n"
      "add.u32 b1, %2, 4;\n"
      "mov.b64 da, {a1, %3};\n"
      "mov.b64 db, {b1, %4};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n"
      "add.u32 a1, %1, 6;\n"
      "add.u32 b1, %2, 6;\n"
      "mov.b64 da, {a1, %3};\n"
      "mov.b64 db, {b1, %4};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n"
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%7];\n"
      "}\n" ::"r"(tmem_d),
      "r"(a_lo), "r"(b_lo), "r"(a_hi), "r"(b_hi), "r"(idesc), "r"(accumulate), "r"(commit_bar)
      : "memory");
}
// Same four MMAs without the commit (operands that stay resident in shared memory need no release).
__device__ __forceinline__ void umma_bf16_kblock64_nc(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo,
                                                      uint32_t a_hi, uint32_t b_hi, uint32_t idesc,
                                                      uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b64 da, db;\n"
      ".reg .b32 a1, b1;\n"
      "setp.ne.b32 p, %6, 0;\n"
      "mov.b64 da, {%1, %3};\n"
      "mov.b64 db, {%2, %4};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n"
      "setp.eq.b32 p, 0, 0;\n"
      "add.u32 a1, %1, 2;\n"
      "add.u32 b1, %2, 2;\n"
      "mov.b64 da, {a1, %3};\n"
      "mov.b64 db, {b1, %4};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n"
      "add.u32 a1, %1, 4;\n"
      "add.u32 b1, %2, 4;\n"
      "mov.b64 da, {a1, %3};\n"
      "mov.b64 db, {b1, %4};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n"
      "add.u32 a1, %1, 6;\n"
      "add.u32 b1, %2, 6;\n"
      "mov.b64 da, {a1, %3};\n"
      "mov.b64 db, {b1, %4};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(a_lo), "r"(b_lo), "r"(a_hi), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Two K steps (2 x 16 deep) of one MMA group with +2 descriptor steps on both operands, no commit: one filter row of the
// first-layer kernel in the pixel-pair layout (a_hi / b_hi are the constant high descriptor words).
__device__ __forceinline__ void umma_bf16_2steps_nc(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t a_hi,
                                                    uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b64 da, db;\n"
      ".reg .b32 a1, b1;\n"
      "setp.ne.b32 p, %6, 0;\n"
      "mov.b64 da, {%1, %3};\n"
      "mov.b64 db, {%2, %4};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n"
      "setp.eq.b32 p, 0, 0;\n"
      "add.u32 a1, %1, 2;\n"
      "add.u32 b1, %2, 2;\n"
      "mov.b64 da, {a1, %3};\n"
      "mov.b64 db, {b1, %4};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(a_lo), "r"(b_lo), "r"(a_hi), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// All previously issued tcgen05.mma of this thread arrive on `bar` when complete.
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
// 32 lanes x 16 consecutive fp32 columns -> 16 registers per thread (thread i <-> lane base+i).
__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, "
      "%13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// registers -> TMEM: 32 lanes x 8 consecutive 32-bit columns (thread i <-> lane base+i)
__device__ __forceinline__ void tmem_st_x8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
               "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem]: the A operand (128 rows = lanes, K = 16 bf16 = 8 packed 32-bit columns per
// instruction) is read from tensor memory (softmax probabilities written by tcgen05.st), B through its descriptor.
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread
__device__ __forceinline__ void tmem_ld_x32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
      "%15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// 32 lanes x 8 consecutive fp32 columns -> 8 registers per thread
__device__ __forceinline__ void tmem_ld_x8(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, 128-byte-swizzled shared-memory matrix descriptor (rows of 64 bf16 = 128 B, 8-row
// swizzle atoms of 1024 B stacked along M/N). Field layout: PTX ISA "tcgen05 shared memory
// descriptor": [0,14) addr>>4, [16,30) LBO>>4, [32,46) SBO>>4, [46,48) version=1, [61,64) swizzle.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;            // LBO (unused for swizzled K-major)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;    // SBO: 8 rows * 128 B
  d |= static_cast<uint64_t>(1) << 46;            // descriptor version (sm_100)
  d |= static_cast<uint64_t>(2) << 61;            // SWIZZLE_128B
  return d;
}
// Instruction descriptor for kind::f16: bf16 x bf16 -> fp32, both operands K-major, M=128.
__host__ __device__ inline uint32_t umma_idesc_bf16_m128(uint32_t n) {
  uint32_t d = 0;
  d |= 1u << 4;          // D format: F32
  d |= 1u << 7;          // A format: BF16
  d |= 1u << 10;         // B format: BF16
  d |= (n >> 3) << 17;   // N / 8
  d |= (128u >> 4) << 24;  // M / 16
  return d;
}

// kind::f16 instruction descriptor with B read MN-major (b_major bit 16): B[n][k] stored with n
// contiguous, e.g. the V operand of attention ([key][head_dim] rows) used as B[head_dim x key].
__host__ __device__ inline uint32_t umma_idesc_bf16_m128_bmn(uint32_t n) {
  return umma_idesc_bf16_m128(n) | (1u << 16);
}

// ----------------------------------------------------------------------------------------------
// CTA pair (cta_group::2): two SMs of one TPC execute one 256-row MMA; the leader (cluster rank 0)
// issues it, both CTAs feed their own shared memory. A shared::cta address with bit 24 cleared
// names the same offset in the leader CTA's shared memory (shared::cluster window).
// ----------------------------------------------------------------------------------------------
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the LEADER CTA's barrier at the same offset (valid from either CTA of the pair)
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar & kPeerBitMask) : "memory");
}
// TMA loads whose completion bytes are credited to the LEADER CTA's barrier
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, "
      "%4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_pair(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1,
                                                 int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, "
      "%4, %5, %6}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// commit of the pair's MMAs, delivered to the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
  asm volatile(
      "{\n"
      ".reg .b16 m;\n"
      "mov.b16 m, 3;\n"
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n"
      "}\n" ::"r"(bar)
      : "memory");
}
// one 64-deep K block of a 256 x N pair MMA (see umma_bf16_kblock64) + multicast release of the stage
__device__ __forceinline__ void umma_bf16_kblock64_pair(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo,
                                                        uint32_t a_hi, uint32_t b_hi, uint32_t idesc,
                                                        uint32_t accumulate, uint32_t commit_bar) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b64 da, db;\n"
      ".reg .b32 a1, b1;\n"
      ".reg .b16 m;\n"
      "mov.b16 m, 3;\n"
      "setp.ne.b32 p, %6, 0;\n"
      "mov.b64 da, {%1, %3};\n"
      "mov.b64 db, {%2, %4};\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n"
      "setp.eq.b32 p, 0, 0;\n"
      "add.u32 a1, %1, 2;\n"
      "add.u32 b1, %2, 2;\n"
      "mov.b64 da, {a1, %3};\n"
      "mov.b64 db, {b1, %4};\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n"
      "add.u32 a1, %1, 4;\n"
      "add.u32 b1, %2, 4;\n"
      "mov.b64 da, {a1, %3};\n"
      "mov.b64 db, {b1, %4};\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n"
      "add.u32 a1, %1, 6;\n"
      "add.u32 b1, %2, 6;\n"
      "mov.b64 da, {a1, %3};\n"
      "mov.b64 db, {b1, %4};\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n"
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%7], m;\n"
      "}\n" ::"r"(tmem_d),
      "r"(a_lo), "r"(b_lo), "r"(a_hi), "r"(b_hi), "r"(idesc), "r"(accumulate), "r"(commit_bar)
      : "memory");
}
// Same four pair MMAs without the commit (operands that stay resident / are released by a later commit).
__device__ __forceinline__ void umma_bf16_kblock64_pair_nc(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo,
                                                           uint32_t a_hi, uint32_t b_hi, uint32_t idesc,
                                                           uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b64 da, db;\n"
      ".reg .b32 a1, b1;\n"
      "setp.ne.b32 p, %6, 0;\n"
      "mov.b64 da, {%1, %3};\n"
      "mov.b64 db, {%2, %4};\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n"
      "setp.eq.b32 p, 0, 0;\n"
      "add.u32 a1, %1, 2;\n"
      "add.u32 b1, %2, 2;\n"
      "mov.b64 da, {a1, %3};\n"
      "mov.b64 db, {b1, %4};\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n"
      "add.u32 a1, %1, 4;\n"
      "add.u32 b1, %2, 4;\n"
      "mov.b64 da, {a1, %3};\n"
      "mov.b64 db, {b1, %4};\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n"
      "add.u32 a1, %1, 6;\n"
      "add.u32 b1, %2, 6;\n"
      "mov.b64 da, {a1, %3};\n"
      "mov.b64 db, {b1, %4};\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(a_lo), "r"(b_lo), "r"(a_hi), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Shared-memory DATA handed from the threads of either CTA of a pair to the leader's MMA issuer (an operand tile written
// with st.shared, then read by tcgen05.mma.cta_group::2 in the writer's own SM): the writer fences its generic-proxy
// writes for the async proxy and arrives with release semantics at cluster scope on the LEADER's barrier; the issuer
// waits with acquire semantics at cluster scope.
__device__ __forceinline__ void mbar_arrive_leader_release(uint32_t bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar & kPeerBitMask) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred P;\n"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P, [%1], %2;\n"
      "selp.u32 %0, 1, 0, P;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
static __device__ __noinline__ void mbar_wait_cluster_slow(uint32_t bar, uint32_t parity) {
  const long long t0 = clock64();
  uint32_t spins = 0;
  while (!mbar_try_wait_cluster(bar, parity)) {
    if (((++spins) & 0x3ff) == 0 && (clock64() - t0) > 8000000000LL) {
      printf("eqxv: mbarrier watchdog (cluster): block %d thread %d bar 0x%x parity %u\n", (int)blockIdx.x,
             (int)threadIdx.x, bar, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  if (!mbar_try_wait_cluster(bar, parity)) mbar_wait_cluster_slow(bar, parity);
}
// pair flavour of umma_bf16_ts: D[tmem] (+)= A[tmem] * B[smem] over both CTAs (each CTA's 128 A rows in its own TMEM)
__device__ __forceinline__ void umma_bf16_ts_pair(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo, uint32_t b_hi,
                                                  uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b64 db;\n"
      "setp.ne.b32 p, %5, 0;\n"
      "mov.b64 db, {%2, %3};\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], db, %4, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// kind::f16 instruction descriptor for the 256-row pair MMA
__host__ __device__ inline uint32_t umma_idesc_bf16_m256(uint32_t n) {
  uint32_t d = 0;
  d |= 1u << 4;
  d |= 1u << 7;
  d |= 1u << 10;
  d |= (n >> 3) << 17;
  d |= (256u >> 4) << 24;
  return d;
}

// Byte offset of 16-byte chunk `j` (0..7) of row `r` inside a 128B-swizzled tile whose base is
// 1024-byte aligned (the layout TMA writes/reads with CU_TENSOR_MAP_SWIZZLE_128B).
__device__ __forceinline__ uint32_t sw128_off(uint32_t r, uint32_t j) {
  return r * 128u + ((j ^ (r & 7u)) << 4);
}

__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

}  // namespace eqxv
