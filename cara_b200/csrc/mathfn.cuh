// Scalar math shared by the kernels: exact-erf GELU (timm's nn.GELU, cara.py:84) and its derivative.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

namespace cara {

// erf by Abramowitz-Stegun 7.1.26 (|abs err| <= 1.5e-7): 2 MUFU + ~8 FMA, used in GEMM epilogues
// where the exact erff() would make the epilogue slower than the tensor-core main loop.
__device__ __forceinline__ float erf_fast(float x) {
  const float ax = fabsf(x);
  const float t = __fdividef(1.0f, fmaf(0.3275911f, ax, 1.0f));
  float y = fmaf(t, 1.061405429f, -1.453152027f);
  y = fmaf(t, y, 1.421413741f);
  y = fmaf(t, y, -0.284496736f);
  y = fmaf(t, y, 0.254829592f);
  y *= t;
  const float e = 1.0f - y * __expf(-ax * ax);
  return copysignf(e, x);
}
// The two epilogue forms below are the same A&S 7.1.26 erf with the algebra folded so that each costs 2 MUFU
// (rcp, ex2) and ~11 FP32 instructions:  with a = |u|, t = 1 / (1 + p a / sqrt 2), w = t poly(t), e = exp(-u^2 / 2)
//   erf(a / sqrt 2) = 1 - w e
//   GELU(u)  = u Phi(u)         = max(u, 0) - (a / 2) w e
//   GELU'(u) = Phi(u) + u phi(u) = [u >= 0] + sign(u) e (a / sqrt(2 pi) - w / 2)
__device__ __forceinline__ void gelu_terms(float a, float& w, float& e) {
  const float t = __fdividef(1.0f, fmaf(0.3275911f * 0.70710678118654752f, a, 1.0f));
  float y = fmaf(t, 1.061405429f, -1.453152027f);
  y = fmaf(t, y, 1.421413741f);
  y = fmaf(t, y, -0.284496736f);
  y = fmaf(t, y, 0.254829592f);
  w = y * t;
  const float z = a * 0.84932180028801904f;          // sqrt(log2(e) / 2): e = 2^(-z^2) = exp(-a^2 / 2)
  e = exp2f(-z * z);
}
// Forward epilogue form with ONE MUFU and 9 FP32 instructions (the fc1 epilogue is issue-bound):
//   GELU(u) = max(u, 0) - a Phi(-a),  a = |u|,  Phi(-a) = erfc(a / sqrt 2) / 2 = 2^q(a)
// q = degree-6 polynomial fitted (weighted least squares on Chebyshev nodes of [0, 6], weight a Phi(-a)) to
// log2(erfc(a / sqrt 2) / 2); a is clamped at 6 (Phi(-6) = 1e-9).  |abs err| <= 4e-7 against the exact-erf GELU over
// [-12, 12] in fp32 (tools/gelu_fit.py), the same class as the A&S form used for the derivative.
__device__ __forceinline__ float gelu_fast(float u) {
  const float a = fminf(fabsf(u), 6.0f);
  float q = fmaf(a, 3.042068784e-05f, -7.316191914e-04f);
  q = fmaf(a, q, 7.908979431e-03f);
  q = fmaf(a, q, -5.307019129e-02f);
  q = fmaf(a, q, -4.590840042e-01f);
  q = fmaf(a, q, -1.151080370e+00f);
  q = fmaf(a, q, -1.000007629e+00f);
  return fmaf(-a, exp2f(q), fmaxf(u, 0.0f));
}
__device__ __forceinline__ float gelu_grad_fast(float u) {
  const float a = fabsf(u);
  float w, e;
  gelu_terms(a, w, e);
  const float r = e * fmaf(a, 0.3989422804014327f, -0.5f * w);
  return u < 0.0f ? -r : 1.0f + r;
}
// Both GELU(u) and GELU'(u) for TWO elements at once in packed half arithmetic (fc1's training epilogue keeps the
// derivative for backward instead of the pre-activation, so the fc2 dX epilogue is a plain multiply):
//   p = Phi(-|u|) = 2^q(|u|)  (the same degree-6 fit as gelu_fast),   GELU(u)  = max(u, 0) - |u| p
//   Phi(u) = p + [u >= 0] (1 - 2p),   phi(u) = 2^(-u^2 log2(e)/2 - log2 sqrt(2 pi)),   GELU'(u) = Phi(u) + u phi(u)
// 4 MUFU.EX2.F16 + ~16 packed ops per PAIR.  Inputs are bf16 values (exact in half); half's 11-bit mantissa
// puts the error of p at <= 0.5 % where p matters, i.e. <= 1e-3 absolute on either output -- below the 2^-9 relative
// rounding of the bf16 outputs they are stored as (checked against the exact-erf forms in tests/test_kernels_gpu.py).
__device__ __forceinline__ __half2 ex2_h2(__half2 x) {       // MUFU.EX2.F16 per half (h2exp2 goes through fp32 and back)
  uint32_t r;
  asm("ex2.approx.f16x2 %0, %1;" : "=r"(r) : "r"(*reinterpret_cast<const uint32_t*>(&x)));
  return *reinterpret_cast<__half2*>(&r);
}
__device__ __forceinline__ void gelu_pair_h2(float ux, float uy, float2& g, float2& gp) {
  const __half2 u = __floats2half2_rn(ux, uy);
  const __half2 a = __hmin2(__habs2(u), __float2half2_rn(6.0f));
  __half2 q = __hfma2(a, __float2half2_rn(3.042068784e-05f), __float2half2_rn(-7.316191914e-04f));
  q = __hfma2(a, q, __float2half2_rn(7.908979431e-03f));
  q = __hfma2(a, q, __float2half2_rn(-5.307019129e-02f));
  q = __hfma2(a, q, __float2half2_rn(-4.590840042e-01f));
  q = __hfma2(a, q, __float2half2_rn(-1.151080370e+00f));
  q = __hfma2(a, q, __float2half2_rn(-1.000007629e+00f));
  const __half2 p = ex2_h2(q);
  const __half2 zero = __float2half2_rn(0.0f);
  const __half2 gh = __hfma2(__hneg2(a), p, __hmax2(u, zero));
  const __half2 ge = __hge2(u, zero);                                   // 1.0 where u >= 0
  const __half2 Phi = __hfma2(ge, __hfma2(__float2half2_rn(-2.0f), p, __float2half2_rn(1.0f)), p);
  const __half2 e = ex2_h2(__hfma2(__hmul2(u, u), __float2half2_rn(-0.72134752f), __float2half2_rn(-1.32574806f)));
  const __half2 gph = __hfma2(u, e, Phi);
  g = __half22float2(gh);
  gp = __half22float2(gph);
}
__device__ __forceinline__ float gelu_exact(float u) {
  return 0.5f * u * (1.0f + erff(u * 0.70710678118654752f));
}
__device__ __forceinline__ float gelu_grad_exact(float u) {
  const float cdf = 0.5f * (1.0f + erff(u * 0.70710678118654752f));
  const float pdf = 0.3989422804014327f * expf(-0.5f * u * u);
  return fmaf(u, pdf, cdf);
}

}  // namespace cara
