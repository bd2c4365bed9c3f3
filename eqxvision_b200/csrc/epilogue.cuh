// Epilogue arithmetic shared by the tcgen05 kernels (igemm.cu, bottleneck.cu): activations as template parameters,
// packed fp32x2 math, the 8-column bias / residual / activation / round step, per-column vector staging.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

#include "../../include/eqxv_b200.h"

namespace eqxv {

// The activation is a template parameter of the kernel: a per-element runtime switch inside the
// fully unrolled epilogue turned it into ~200 KB of branchy code and made every chunk I-cache bound
// (measured: 28 us per 128x64 chunk).
template <int kAct>
__device__ __forceinline__ float apply_act(float v) {
  if constexpr (kAct == EQXV_ACT_RELU) {
    return fmaxf(v, 0.f);
  } else if constexpr (kAct == EQXV_ACT_SILU) {
    // x * sigmoid(x) = h + h * tanh(h), h = x / 2: ONE special-function op per element instead of exp + reciprocal.
    // EfficientNet-B4 evaluates SiLU on 2.2 G elements per 128-image step; at 16 MUFU results per SM and clock the
    // exp/rcp form alone costs ~1 ms of a ~6 ms step and made the narrow expand GEMMs MUFU bound.
    const float h = 0.5f * v;
    float th;
    asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(h));
    return fmaf(h, th, h);
  } else if constexpr (kAct == EQXV_ACT_GELU_TANH) {
    const float u = 0.7978845608028654f * (v + 0.044715f * v * v * v);
    float th;
    asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(u));
    return 0.5f * v * (1.f + th);
  } else if constexpr (kAct == EQXV_ACT_HARDSWISH) {
    return v * fminf(fmaxf(v + 3.f, 0.f), 6.f) * (1.f / 6.f);
  } else if constexpr (kAct == EQXV_ACT_SIGMOID) {
    float th;
    asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(0.5f * v));
    return fmaf(0.5f, th, 0.5f);
  } else if constexpr (kAct == EQXV_ACT_HARDSIGMOID) {
    return fminf(fmaxf(v + 3.f, 0.f), 6.f) * (1.f / 6.f);
  } else if constexpr (kAct == EQXV_ACT_RELU6) {
    return fminf(fmaxf(v, 0.f), 6.f);
  } else {
    return v;
  }
}

// ---- packed fp32x2 epilogue math (sm_100 add.f32x2) -------------------------------------------------
__device__ __forceinline__ uint64_t f2_pack(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void f2_unpack(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t f2_add(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// bf16x2 (as u32) -> two fp32 (exact)
__device__ __forceinline__ uint64_t bf2_to_f2(uint32_t v) {
  return f2_pack(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u));
}
__device__ __forceinline__ uint32_t f2_to_bf2(uint64_t v) {
  float lo, hi;
  f2_unpack(v, lo, hi);
  const __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&t);
}

// 8 accumulator columns -> 8 bf16 outputs: + bias (+ residual) -> activation -> round. The common
// "none"/"relu" epilogues run on packed pairs (add.f32x2, max.bf16x2: half the FP32 issue slots);
// relu commutes with the bf16 rounding, so applying it after the conversion is exact.
template <int kAct, int kRes>
__device__ __forceinline__ uint4 epilogue8(const float* v, const float* bias_smem, const uint4 rv) {
  const float4 b0 = *reinterpret_cast<const float4*>(bias_smem);
  const float4 b1 = *reinterpret_cast<const float4*>(bias_smem + 4);
  const uint32_t r[4] = {rv.x, rv.y, rv.z, rv.w};
  uint32_t o[4];
  if constexpr (kAct == EQXV_ACT_NONE || (kAct == EQXV_ACT_RELU && kRes != 2)) {
    const uint64_t bb[4] = {f2_pack(b0.x, b0.y), f2_pack(b0.z, b0.w), f2_pack(b1.x, b1.y), f2_pack(b1.z, b1.w)};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint64_t x = f2_add(f2_pack(v[2 * q], v[2 * q + 1]), bb[q]);
      if constexpr (kRes != 0) x = f2_add(x, bf2_to_f2(r[q]));
      uint32_t y = f2_to_bf2(x);
      if constexpr (kAct == EQXV_ACT_RELU) {
        const __nv_bfloat162 z = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&y),
                                         __floats2bfloat162_rn(0.f, 0.f));
        y = *reinterpret_cast<const uint32_t*>(&z);
      }
      o[q] = y;
    }
  } else {
    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float x0 = v[2 * q] + bb[2 * q], x1 = v[2 * q + 1] + bb[2 * q + 1];
      const float r0 = __uint_as_float(r[q] << 16), r1 = __uint_as_float(r[q] & 0xffff0000u);
      if constexpr (kRes == 1) {
        x0 += r0;
        x1 += r1;
      }
      x0 = apply_act<kAct>(x0);
      x1 = apply_act<kAct>(x1);
      if constexpr (kRes == 2) {
        x0 += r0;
        x1 += r1;
      }
      const __nv_bfloat162 t = __floats2bfloat162_rn(x0, x1);
      o[q] = *reinterpret_cast<const uint32_t*>(&t);
    }
  }
  return make_uint4(o[0], o[1], o[2], o[3]);
}

// LayerNorm-consumer variant (kLN == 1): x = rstd * acc + (bias - mean * rstd * wsum), activation, round.
// All arithmetic on packed fp32 pairs (fma/mul.f32x2): the fc1 epilogue (LayerNorm + tanh-GELU on 128 x 256 outputs per
// tile and CTA, four warps) is as long as its mainloop, so its instruction count is what bounds the layer.
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t f2_mul(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
template <int kAct>
__device__ __forceinline__ uint4 epilogue8_ln(const float* v, const float* bias_smem, const float* wsum_smem,
                                              const float rstd, const float nmr) {
  const float4 b0 = *reinterpret_cast<const float4*>(bias_smem), b1 = *reinterpret_cast<const float4*>(bias_smem + 4);
  const float4 s0 = *reinterpret_cast<const float4*>(wsum_smem), s1 = *reinterpret_cast<const float4*>(wsum_smem + 4);
  const uint64_t bb[4] = {f2_pack(b0.x, b0.y), f2_pack(b0.z, b0.w), f2_pack(b1.x, b1.y), f2_pack(b1.z, b1.w)};
  const uint64_t ss[4] = {f2_pack(s0.x, s0.y), f2_pack(s0.z, s0.w), f2_pack(s1.x, s1.y), f2_pack(s1.z, s1.w)};
  const uint64_t rstd2 = f2_pack(rstd, rstd), nmr2 = f2_pack(nmr, nmr);
  uint32_t o[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint64_t x = f2_fma(f2_pack(v[2 * q], v[2 * q + 1]), rstd2, f2_fma(nmr2, ss[q], bb[q]));
    if constexpr (kAct == EQXV_ACT_GELU_TANH) {
      // 0.5 x (1 + tanh(0.79788 x (1 + 0.044715 x^2)))  =  h + h * tanh(u),  h = x / 2
      const uint64_t inner = f2_fma(f2_mul(x, x), f2_pack(0.044715f, 0.044715f), f2_pack(1.f, 1.f));
      const uint64_t u = f2_mul(f2_mul(x, f2_pack(0.7978845608028654f, 0.7978845608028654f)), inner);
      float u0, u1, t0, t1;
      f2_unpack(u, u0, u1);
      asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(u0));
      asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(u1));
      const uint64_t h = f2_mul(x, f2_pack(0.5f, 0.5f));
      x = f2_fma(h, f2_pack(t0, t1), h);
    } else if constexpr (kAct != EQXV_ACT_NONE) {
      float x0, x1;
      f2_unpack(x, x0, x1);
      x = f2_pack(apply_act<kAct>(x0), apply_act<kAct>(x1));
    }
    o[q] = f2_to_bf2(x);
  }
  return make_uint4(o[0], o[1], o[2], o[3]);
}

// Per-column fp32 vector (folded-BN shift / bias, filter column sums) -> shared memory, zero beyond `count`, once per
// CTA. Every thread issues up to four 16-byte loads BEFORE it stores anything: the persistent CTAs of the next kernel
// only become resident when the previous kernel's CTAs exit, so this prologue is NOT hidden by PDL, and the scalar loop
// it replaces paid one dependent L2 round trip per 288 columns (ncu on ViT-B/16 fc1, N = 3072, bias + column sums:
// 21 % of the kernel's stall samples sat on these two loops, profiles/r02_ncu_pair_vit.txt).
__device__ __forceinline__ void stage_columns(float* dst, const float* __restrict__ src, int count, int ncols_pad) {
  const int n4 = ncols_pad >> 2;   // ncols_pad is a multiple of 16
  const bool vec = src != nullptr && (reinterpret_cast<uintptr_t>(src) & 15) == 0;
  for (int base = 0; base < n4; base += 4 * (int)blockDim.x) {
    float4 v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c = (base + (int)threadIdx.x + k * (int)blockDim.x) * 4;
      v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (src != nullptr && c < count) {
        if (vec && c + 3 < count) {
          v[k] = __ldg(reinterpret_cast<const float4*>(src + c));
        } else {
          v[k].x = __ldg(src + c);
          if (c + 1 < count) v[k].y = __ldg(src + c + 1);
          if (c + 2 < count) v[k].z = __ldg(src + c + 2);
          if (c + 3 < count) v[k].w = __ldg(src + c + 3);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i4 = base + (int)threadIdx.x + k * (int)blockDim.x;
      if (i4 < n4) reinterpret_cast<float4*>(dst)[i4] = v[k];
    }
  }
}

}  // namespace eqxv
