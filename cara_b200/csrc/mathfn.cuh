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
__device__ __forceinline__ float gelu_fast(float u) {
  return 0.5f * u * (1.0f + erf_fast(u * 0.70710678118654752f));
}
// d/du [u * Phi(u)] = Phi(u) + u * phi(u)
__device__ __forceinline__ float gelu_grad_fast(float u) {
  const float cdf = 0.5f * (1.0f + erf_fast(u * 0.70710678118654752f));
  const float pdf = 0.3989422804014327f * __expf(-0.5f * u * u);
  return fmaf(u, pdf, cdf);
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
