// Scalar math shared by the kernels: exact-erf GELU (timm's nn.GELU, cara.py:84) and its derivative.
#pragma once
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
__device__ __forceinline__ float gelu_exact(float u) {
  return 0.5f * u * (1.0f + erff(u * 0.70710678118654752f));
}
__device__ __forceinline__ float gelu_grad_exact(float u) {
  const float cdf = 0.5f * (1.0f + erff(u * 0.70710678118654752f));
  const float pdf = 0.3989422804014327f * expf(-0.5f * u * u);
  return fmaf(u, pdf, cdf);
}

}  // namespace cara
