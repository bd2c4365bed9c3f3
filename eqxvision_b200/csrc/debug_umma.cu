// Hardware probe (not on the product path): does a K-major SWIZZLE_128B UMMA descriptor accept an
// A operand that starts at an arbitrary 128-byte row of a TMA-written tile and whose 8-row groups
// are strided by an arbitrary multiple of 128 B?  This is what a halo-tile implicit GEMM needs to
// reuse one shared-memory copy of the input for all kh*kw filter taps.
//   D[128 x 64] = A[rows shift + (m/8)*group_rows + m%8, 0:64] * B[64 x 64]^T
#include "common.h"
#include "ptx.cuh"

namespace eqxv {

struct alignas(64) DbgParams {
  CUtensorMap tmA, tmB;
  float* out;
  int rows;          // rows of A loaded (<= 256)
  int shift_rows;    // first row used
  int group_rows;    // distance between 8-row groups, in rows (8 = dense)
  int base_offset_mode;  // 0: base_offset field = 0, 1: (start >> 7) & 7
};

__global__ void __launch_bounds__(128, 1) dbg_umma_kernel(const __grid_constant__ DbgParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - raw);
  const uint32_t a_s = base;                 // up to 256 rows x 128 B = 32 KiB
  const uint32_t b_s = base + 32768;         // 64 rows x 128 B
  const uint32_t bar_full = base + 32768 + 8192;
  const uint32_t bar_done = bar_full + 8;
  const uint32_t slot = bar_full + 16;
  volatile uint32_t* slot_g = reinterpret_cast<volatile uint32_t*>(gbase + 32768 + 8192 + 16);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(bar_full, 1);
    mbar_init(bar_done, 1);
    mbar_fence_init();
  }
  if (warp == 0) {
    tmem_alloc(slot, 64);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot_g;
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar_full, (uint32_t)(p.rows * 128 + 64 * 128));
    tma_load_2d(a_s, &p.tmA, bar_full, 0, 0);
    tma_load_2d(b_s, &p.tmB, bar_full, 0, 0);
    mbar_wait(bar_full, 0);
    tc_fence_after();
    const uint32_t start = a_s + (uint32_t)p.shift_rows * 128u;
    uint64_t ad = 0;
    ad |= (uint64_t)((start & 0x3FFFF) >> 4);
    ad |= (uint64_t)1 << 16;
    ad |= (uint64_t)((uint32_t)(p.group_rows * 128) >> 4) << 32;
    ad |= (uint64_t)1 << 46;
    if (p.base_offset_mode == 1) ad |= (uint64_t)((start >> 7) & 7u) << 49;
    ad |= (uint64_t)2 << 61;
    const uint64_t bd = umma_desc_sw128(b_s);
    const uint32_t idesc = umma_idesc_bf16_m128(64);
    for (int k = 0; k < 4; ++k) umma_bf16(tmem, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc, k);
    umma_commit(bar_done);
  }
  __syncwarp();
  mbar_wait(bar_done, 0);
  tc_fence_after();
  float v[64];
  const uint32_t t = tmem + ((uint32_t)(warp * 32) << 16);
  for (int j = 0; j < 4; ++j) tmem_ld_x16(t + j * 16, &v[j * 16]);
  tmem_ld_wait();
  const int row = threadIdx.x;
  for (int j = 0; j < 64; ++j) p.out[row * 64 + j] = v[j];
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 64);
  }
}

}  // namespace eqxv

using namespace eqxv;

extern "C" int eqxv_debug_umma_shift(const void* a, int32_t rows, const void* b, float* out,
                                     int32_t shift_rows, int32_t group_rows, int32_t base_offset_mode,
                                     void* stream) {
  EQXV_CHECK_ARG(a && b && out && rows >= 128 && rows <= 256, "debug_umma: bad arguments");
  DbgParams p;
  memset(&p, 0, sizeof(p));
  TmapSpec sa{};
  sa.base = const_cast<void*>(a);
  sa.dtype = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  sa.rank = 2;
  sa.dims[0] = 64, sa.dims[1] = (uint64_t)rows;
  sa.strides_bytes[0] = 128;
  sa.box[0] = 64, sa.box[1] = (uint32_t)rows;
  sa.estride[0] = sa.estride[1] = 1;
  sa.swizzle = CU_TENSOR_MAP_SWIZZLE_128B;
  int rc = encode_tmap(&p.tmA, sa);
  if (rc) return rc;
  TmapSpec sb = sa;
  sb.base = const_cast<void*>(b);
  sb.dims[1] = 64;
  sb.box[1] = 64;
  rc = encode_tmap(&p.tmB, sb);
  if (rc) return rc;
  p.out = out;
  p.rows = rows, p.shift_rows = shift_rows, p.group_rows = group_rows, p.base_offset_mode = base_offset_mode;
  static bool attr = false;
  if (!attr) {
    EQXV_CUDA(cudaFuncSetAttribute(dbg_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    attr = true;
  }
  dbg_umma_kernel<<<1, 128, 32768 + 8192 + 64 + 1024, (cudaStream_t)stream>>>(p);
  EQXV_CUDA(cudaGetLastError());
  return EQXV_OK;
}
