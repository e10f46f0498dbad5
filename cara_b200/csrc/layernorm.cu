// LayerNorm forward / backward (timm Block.norm1/norm2, eps 1e-6) fused with the residual stream.
//
// HBM-bound: one warp per token row, 128-bit coalesced accesses, two-pass statistics held in
// registers, warp-shuffle reductions.  The residual stream x is fp32 (SURVEY hard part 2); the
// normalised activations handed to the tensor-core GEMMs are bf16 (or fp32 in fp32 mode).
//
//   fwd : x_out = x_in + rowscale[b] * delta      (residual add + DropPath of the previous branch)
//         h     = LN(x_out) * gamma + beta ; mean/rstd saved for backward
//   bwd : dx_out = dx_in + LN'(dh)                (gamma/beta are frozen: no parameter grads)
//         g_out  = bf16(rowscale[b] * dx_out)     (the next branch's incoming gradient, pre-cast)
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"

namespace cara {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float4 load_bf16x4(const __nv_bfloat16* p) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&u.x);
  const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
  const float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}
__device__ __forceinline__ void store_bf16x4(__nv_bfloat16* p, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&a);
  u.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = u;
}
__device__ __forceinline__ float4 load4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 load4(const __nv_bfloat16* p) { return load_bf16x4(p); }
__device__ __forceinline__ void store4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void store4(__nv_bfloat16* p, float4 v) { store_bf16x4(p, v); }

// NV = C / 128 float4 chunks per lane.
template <int NV, typename ActT>
__global__ void __launch_bounds__(256)
ln_fwd_kernel(const float* __restrict__ x_in, const ActT* __restrict__ delta, const float* __restrict__ rowscale,
              int rows_per_sample, float* __restrict__ x_out, const float* __restrict__ gamma,
              const float* __restrict__ beta, ActT* __restrict__ h, float* __restrict__ mean_out,
              float* __restrict__ rstd_out, int M, float eps) {
  constexpr int C = NV * 128;
  pdl_wait();
  pdl_trigger();
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const size_t base = static_cast<size_t>(row) * C;
  float4 v[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = load4(x_in + base + (i * 32 + lane) * 4);
  if (delta != nullptr) {
    const float rs = rowscale != nullptr ? rowscale[row / rows_per_sample] : 1.0f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float4 d = load4(delta + base + (i * 32 + lane) * 4);
      v[i].x = fmaf(rs, d.x, v[i].x); v[i].y = fmaf(rs, d.y, v[i].y);
      v[i].z = fmaf(rs, d.z, v[i].z); v[i].w = fmaf(rs, d.w, v[i].w);
    }
    if (x_out != nullptr) {
#pragma unroll
      for (int i = 0; i < NV; ++i) store4(x_out + base + (i * 32 + lane) * 4, v[i]);
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  const float mean = warp_sum(s) * (1.0f / C);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / C) + eps);
  if (h != nullptr) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c0 = (i * 32 + lane) * 4;
      const float4 g = load4(gamma + c0), b = load4(beta + c0);
      float4 o;
      o.x = (v[i].x - mean) * rstd * g.x + b.x; o.y = (v[i].y - mean) * rstd * g.y + b.y;
      o.z = (v[i].z - mean) * rstd * g.z + b.z; o.w = (v[i].w - mean) * rstd * g.w + b.w;
      store4(h + base + c0, o);
    }
  }
  if (lane == 0 && mean_out != nullptr) { mean_out[row] = mean; rstd_out[row] = rstd; }
}

template <int NV, typename ActT>
__global__ void __launch_bounds__(256)
ln_bwd_kernel(const ActT* __restrict__ dh, const float* __restrict__ x, const float* __restrict__ mean_in,
              const float* __restrict__ rstd_in, const float* __restrict__ gamma, const float* __restrict__ dx_in,
              float* __restrict__ dx_out, ActT* __restrict__ g_out, const float* __restrict__ rowscale,
              int rows_per_sample, int M) {
  constexpr int C = NV * 128;
  pdl_wait();
  pdl_trigger();
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const size_t base = static_cast<size_t>(row) * C;
  const float mean = mean_in[row], rstd = rstd_in[row];
  float4 dy[NV], xh[NV], rin[NV];
  float s1 = 0.f, s2 = 0.f;
  // the incoming residual gradient is requested together with dh and x (one memory round trip per row, not two:
  // it is only needed after the two warp reductions)
  if (dx_in != nullptr) {
#pragma unroll
    for (int i = 0; i < NV; ++i) rin[i] = load4(dx_in + base + (i * 32 + lane) * 4);
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c0 = (i * 32 + lane) * 4;
    const float4 d = load4(dh + base + c0), g = load4(gamma + c0), xv = load4(x + base + c0);
    dy[i] = make_float4(d.x * g.x, d.y * g.y, d.z * g.z, d.w * g.w);
    xh[i] = make_float4((xv.x - mean) * rstd, (xv.y - mean) * rstd, (xv.z - mean) * rstd, (xv.w - mean) * rstd);
    s1 += (dy[i].x + dy[i].y) + (dy[i].z + dy[i].w);
    s2 += (dy[i].x * xh[i].x + dy[i].y * xh[i].y) + (dy[i].z * xh[i].z + dy[i].w * xh[i].w);
  }
  const float m1 = warp_sum(s1) * (1.0f / C), m2 = warp_sum(s2) * (1.0f / C);
  const float rs = (g_out != nullptr && rowscale != nullptr) ? rowscale[row / rows_per_sample] : 1.0f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c0 = (i * 32 + lane) * 4;
    float4 o;
    o.x = rstd * (dy[i].x - m1 - xh[i].x * m2); o.y = rstd * (dy[i].y - m1 - xh[i].y * m2);
    o.z = rstd * (dy[i].z - m1 - xh[i].z * m2); o.w = rstd * (dy[i].w - m1 - xh[i].w * m2);
    if (dx_in != nullptr) { o.x += rin[i].x; o.y += rin[i].y; o.z += rin[i].z; o.w += rin[i].w; }
    store4(dx_out + base + c0, o);
    if (g_out != nullptr) store4(g_out + base + c0, make_float4(o.x * rs, o.y * rs, o.z * rs, o.w * rs));
  }
}

template <typename ActT>
static int ln_fwd_t(const LnFwdArgs& a, cudaStream_t st) {
  const int grid = (a.M + 7) / 8;
  cudaError_t err = cudaSuccess;
#define CARA_LN_FWD(NV)                                                                                       \
  case NV:                                                                                                    \
    err = launch_pdl_f<4>(ln_fwd_kernel<NV, ActT>, dim3(grid), dim3(256), 0, st, a.x_in,                           \
                     static_cast<const ActT*>(a.delta), a.rowscale, a.rows_per_sample, a.x_out, a.gamma, a.beta, \
                     static_cast<ActT*>(a.h), a.mean, a.rstd, a.M, a.eps);                                    \
    break;
  switch (a.C / 128) {
    CARA_LN_FWD(1) CARA_LN_FWD(2) CARA_LN_FWD(3) CARA_LN_FWD(4) CARA_LN_FWD(6) CARA_LN_FWD(8) CARA_LN_FWD(10)
    default: return -20;
  }
#undef CARA_LN_FWD
  return err == cudaSuccess ? 0 : -21;
}

template <typename ActT>
static int ln_bwd_t(const LnBwdArgs& a, cudaStream_t st) {
  const int grid = (a.M + 7) / 8;
  cudaError_t err = cudaSuccess;
#define CARA_LN_BWD(NV)                                                                                       \
  case NV:                                                                                                    \
    err = launch_pdl_f<4>(ln_bwd_kernel<NV, ActT>, dim3(grid), dim3(256), 0, st, static_cast<const ActT*>(a.dh),   \
                     a.x, a.mean, a.rstd, a.gamma, a.dx_in, a.dx_out, static_cast<ActT*>(a.g_out), a.rowscale, \
                     a.rows_per_sample, a.M);                                                                 \
    break;
  switch (a.C / 128) {
    CARA_LN_BWD(1) CARA_LN_BWD(2) CARA_LN_BWD(3) CARA_LN_BWD(4) CARA_LN_BWD(6) CARA_LN_BWD(8) CARA_LN_BWD(10)
    default: return -20;
  }
#undef CARA_LN_BWD
  return err == cudaSuccess ? 0 : -21;
}

int ln_fwd_launch(const LnFwdArgs& a, cudaStream_t st) {
  if (a.M <= 0 || a.C % 128 != 0) return -20;
  return a.act_fp32 ? ln_fwd_t<float>(a, st) : ln_fwd_t<__nv_bfloat16>(a, st);
}
int ln_bwd_launch(const LnBwdArgs& a, cudaStream_t st) {
  if (a.M <= 0 || a.C % 128 != 0) return -20;
  return a.act_fp32 ? ln_bwd_t<float>(a, st) : ln_bwd_t<__nv_bfloat16>(a, st);
}

}  // namespace cara
