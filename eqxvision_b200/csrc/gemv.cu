// Matrix-vector products (ONE row), on CUDA cores: out[1, n] = act(a[1, :k] @ w[n, :k]^T + bias).
//
// AlexNet's single-image classifier (alexnet.py:57-65, BASELINE configs[0]: 9216 -> 4096 -> 4096 -> 1000) is 117 MB of
// filters read once per image. The tcgen05 path gives such a problem one 128-row tile per 256 output columns: 16 CTAs
// stream the whole filter (77 us for the three layers, 0.11 of the HBM roofline). Here every warp owns one output column
// and walks its filter row with four 16-byte loads in flight per lane (2 KB per warp, ~130 KB per SM: enough to keep
// HBM busy), the activation row sits in shared memory, fp32 accumulation, one shuffle reduction per output.
// Only for m == 1 (the reference README's single-image call): the summation order differs from the tensor-core path, and
// a batch of 2..N images must give every image the bits it gets in any other batch (tests/test_gpu_models.py::
// test_batch_is_a_pure_map_bitwise, the sharding property the multi-GPU path relies on).
// Same contract as the 1x1 path of eqxv_conv2d_igemm_bf16 (incl. EQXV_FLAG_K_TAIL_SHIFT filters); no residual.
#include <cstdlib>

#include "common.h"
#include "ptx.cuh"
#include "epilogue.cuh"

namespace eqxv {

struct GemvParams {
  const __nv_bfloat16* a;
  const __nv_bfloat16* w;
  const float* bias;
  void* out;
  long long lda, ldw, ldo;
  int m, n, k;
  int w_head8, w_off8;   // K_TAIL_SHIFT, in 8-element vectors: logical vector j >= w_head8 lives at packed vector j + w_off8
};

template <int MT, int kAct, bool kOutF32>
__global__ void __launch_bounds__(256) gemv_kernel(const GemvParams p) {
  extern __shared__ uint4 gemv_sa[];   // [MT][k / 8] bf16 vectors
  griddep_wait();   // PDL: the predecessor kernel has completed (ptx.cuh)
  griddep_launch();
  const int kv = p.k >> 3;
  for (int i = threadIdx.x; i < MT * kv; i += 256) {
    const int r = i / kv, j = i - r * kv;
    gemv_sa[i] = r < p.m ? __ldg(reinterpret_cast<const uint4*>(p.a + (long long)r * p.lda) + j) : make_uint4(0u, 0u, 0u, 0u);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int col = blockIdx.x * 8 + warp;
  if (col >= p.n) return;
  float acc[MT];
#pragma unroll
  for (int r = 0; r < MT; ++r) acc[r] = 0.f;
  const uint4* wrow = reinterpret_cast<const uint4*>(p.w + (long long)col * p.ldw);
  for (int j0 = lane; j0 < kv; j0 += 128) {
    uint4 wv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = j0 + 32 * u;
      wv[u] = j < kv ? __ldg(wrow + (j < p.w_head8 ? j : j + p.w_off8)) : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = j0 + 32 * u;
      if (j < kv) {
        const uint32_t ww[4] = {wv[u].x, wv[u].y, wv[u].z, wv[u].w};
#pragma unroll
        for (int r = 0; r < MT; ++r) {
          const uint4 av = gemv_sa[r * kv + j];
          const uint32_t aw[4] = {av.x, av.y, av.z, av.w};
          float s = acc[r];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            s = fmaf(__uint_as_float(aw[q] << 16), __uint_as_float(ww[q] << 16), s);
            s = fmaf(__uint_as_float(aw[q] & 0xffff0000u), __uint_as_float(ww[q] & 0xffff0000u), s);
          }
          acc[r] = s;
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < MT; ++r) {
    float s = acc[r];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0 && r < p.m) {
      const float v = apply_act<kAct>(s + (p.bias != nullptr ? __ldg(p.bias + col) : 0.f));
      if (kOutF32) {
        static_cast<float*>(p.out)[(long long)r * p.ldo + col] = v;
      } else {
        static_cast<__nv_bfloat16*>(p.out)[(long long)r * p.ldo + col] = __float2bfloat16_rn(v);
      }
    }
  }
}

using GemvFn = void (*)(const GemvParams);
#define EQXV_GEMV_ROW(MT, F32)                                                                                     \
  {                                                                                                                \
    gemv_kernel<MT, 0, F32>, gemv_kernel<MT, 1, F32>, gemv_kernel<MT, 2, F32>, gemv_kernel<MT, 3, F32>,            \
        gemv_kernel<MT, 4, F32>, gemv_kernel<MT, 5, F32>, gemv_kernel<MT, 6, F32>, gemv_kernel<MT, 7, F32>         \
  }
static GemvFn gemv_table(bool f32, int act) {
  static const GemvFn t[2][8] = {EQXV_GEMV_ROW(1, false), EQXV_GEMV_ROW(1, true)};
  return t[f32 ? 1 : 0][act];
}
constexpr int kGemvMaxSmem = 160 * 1024;

int gemv_init() {
  for (int f = 0; f < 2; ++f)
    for (int a = 0; a < 8; ++a)
      EQXV_CUDA(cudaFuncSetAttribute(gemv_table(f != 0, a), cudaFuncAttributeMaxDynamicSharedMemorySize, kGemvMaxSmem));
  return EQXV_OK;
}

bool gemv_applies(long long m, int n, int k) {
  static const bool off = getenv("EQXV_NO_GEMV") != nullptr;
  return !off && m == 1 && k % 8 == 0 && n >= 64 && (long long)k * 2 <= kGemvMaxSmem;
}

int launch_gemv(const void* a, long long lda, const void* w, long long ldw, int w_head, int w_off, const float* bias,
                void* out, long long ldo, long long m, int n, int k, int act, bool out_f32, cudaStream_t stream) {
  EQXV_CHECK_ARG(act >= 0 && act < 8, "gemm(few rows): unknown activation %d", act);
  EQXV_CHECK_ARG(lda % 8 == 0 && ldw % 8 == 0 && ((uintptr_t)a & 15) == 0 && ((uintptr_t)w & 15) == 0,
                 "gemm(few rows): operands must be 16-byte aligned with row pitches that are multiples of 8");
  GemvParams p{};
  p.a = static_cast<const __nv_bfloat16*>(a), p.w = static_cast<const __nv_bfloat16*>(w), p.bias = bias, p.out = out;
  p.lda = lda, p.ldw = ldw, p.ldo = ldo, p.m = (int)m, p.n = n, p.k = k;
  p.w_head8 = w_off ? w_head / 8 : k / 8, p.w_off8 = w_off / 8;
  EQXV_CUDA(launch_kernel(gemv_table(out_f32, act), dim3((unsigned)ceil_div(n, 8)), dim3(256), (size_t)k * 2, stream, p));
  EQXV_CUDA(cudaGetLastError());
  return EQXV_OK;
}

}  // namespace eqxv
