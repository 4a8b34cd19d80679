// Packed fp32x2 epilogue arithmetic shared by the tensor-core kernels (gemm_tc.cu, dec_fused.cu).
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace tt {

// GELU(x) = x * Phi(x) evaluated as 0.5 x (1 + tanh(x (a + b x^2 + c x^4))) with (a, b, c) fitted to
// the erf form: |error| <= 2.6e-5 on the real line (the textbook tanh form is off by 4.7e-4), plus
// tanh.approx.f32's 2^-11 relative error -- both far below the bf16 rounding applied to the result.
// 8 issue slots + 1 MUFU per element: the exact erf costs > 20 and made fc1's epilogue the bottleneck.
// (evaluated two elements at a time with packed fp32x2 arithmetic: gelu_fast2 below)

// Packed fp32x2 arithmetic (sm_100 FFMA2/FMUL2/FADD2): one issue slot per two elements.  With two epilogue warps
// per scheduler the epilogue, not the tensor pipe, set the pace of the K=384 GEMMs (ncu: 14.5 warp instructions
// per output element, issue slots 50 % busy, tensor pipe 37 %; profiles/r1b_gemm_roles.md).
__device__ __forceinline__ uint64_t pk2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ uint64_t pk2u(uint32_t lo, uint32_t hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}
__device__ __forceinline__ void upk2u(uint64_t v, uint32_t& lo, uint32_t& hi) { asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v)); }
__device__ __forceinline__ void upk2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// gelu_fast on two elements at once
__device__ __forceinline__ uint64_t gelu_fast2(uint64_t x) {
  float a, b;
  upk2(mul2(x, x), a, b);
  const uint64_t x2 = pk2(fminf(a, 36.0f), fminf(b, 36.0f));
  uint64_t p = fma2(x2, pk2(-0.00035151678866f, -0.00035151678866f), pk2(0.037005646023f, 0.037005646023f));
  p = fma2(x2, p, pk2(0.797507884285f, 0.797507884285f));
  upk2(mul2(x, p), a, b);
  float ta, tb;
  asm("tanh.approx.f32 %0, %1;" : "=f"(ta) : "f"(a));
  asm("tanh.approx.f32 %0, %1;" : "=f"(tb) : "f"(b));
  const uint64_t hx = mul2(x, pk2(0.5f, 0.5f));
  return fma2(hx, pk2(ta, tb), hx);
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

}  // namespace tt
