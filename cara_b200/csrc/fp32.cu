// fp32 mode of the hot path (north_star: "fp32 mode: logits rel-err <= 1e-4" against the reference's PyTorch path).
//
// Everything here is plain SIMT fp32 -- no tensor cores, no bf16 anywhere -- and is meant for parity runs, not for
// throughput: the exact-erf GELU of timm's Mlp (cara.py:84) and the attention core of cp_attn (cara.py:44-48) with
// the whole head resident in shared memory.  The fp32 projections run on sgemm_kernel (misc.cu) and the fp32
// LayerNorm variants of layernorm.cu.
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "kernels.h"

namespace cara {
namespace {

__device__ __forceinline__ float warp_sum32(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max32(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__global__ void gelu_f32_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, long n) {
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const float u = x[i];
    y[i] = 0.5f * u * (1.0f + erff(u * 0.70710678118654752f));
  }
}
__global__ void gelu_f32_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, float* __restrict__ dx, long n) {
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const float u = x[i];
    const float cdf = 0.5f * (1.0f + erff(u * 0.70710678118654752f));
    const float pdf = 0.3989422804014327f * expf(-0.5f * u * u);
    dx[i] = dy[i] * fmaf(u, pdf, cdf);
  }
}

constexpr int F32_THREADS = 256, F32_WARPS = 8, F32_MAXT = 9;   // up to 9 * 32 = 288 keys per head

// one (sample, head) per CTA; tiles are [N][D + 1] floats (the +1 keeps lane-per-row accesses conflict-free)
__device__ __forceinline__ void load_tile(float* dst, const float* src, long pitch, int N, int D, int tid) {
  for (int i = tid; i < N * D; i += F32_THREADS) dst[(i / D) * (D + 1) + i % D] = src[static_cast<long>(i / D) * pitch + i % D];
}

// scores of one row `a` (broadcast from smem) against all rows of tile `t`: lane owns rows lane, lane + 32, ...
__device__ __forceinline__ void row_dots(const float* a, const float* t, int N, int D, int lane, float (&s)[F32_MAXT]) {
#pragma unroll
  for (int u = 0; u < F32_MAXT; ++u) {
    const int j = lane + 32 * u;
    float acc = 0.f;
    if (j < N) {
      const float* r = t + j * (D + 1);
      for (int d = 0; d < D; ++d) acc = fmaf(a[d], r[d], acc);
    }
    s[u] = acc;
  }
}
// out[d] (lane owns d = lane, lane + 32, lane + 64) = sum_j w_j t[j][d] with w_j held lane-wise as in row_dots
__device__ __forceinline__ void weighted_rows(const float (&w)[F32_MAXT], const float* t, int N, int D, int lane, float (&o)[3]) {
  o[0] = o[1] = o[2] = 0.f;
#pragma unroll
  for (int u = 0; u < F32_MAXT; ++u) {
    if (32 * u >= N) break;
    for (int jj = 0; jj < 32; ++jj) {
      const int j = 32 * u + jj;
      const float wj = __shfl_sync(0xffffffffu, w[u], jj);
      if (j < N) {
        const float* r = t + j * (D + 1);
        if (lane < D) o[0] = fmaf(wj, r[lane], o[0]);
        if (lane + 32 < D) o[1] = fmaf(wj, r[lane + 32], o[1]);
        if (lane + 64 < D) o[2] = fmaf(wj, r[lane + 64], o[2]);
      }
    }
  }
}

__global__ void __launch_bounds__(F32_THREADS)
attn_f32_fwd_kernel(const float* __restrict__ qkv, float* __restrict__ o, float* __restrict__ lse, int B, int N, int H,
                    int D, float scale) {
  extern __shared__ float sm[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const long pitch = 3L * H * D;
  const float* base = qkv + static_cast<long>(b) * N * pitch + h * D;
  float* Ks = sm;
  float* Vs = Ks + N * (D + 1);
  float* qs = Vs + N * (D + 1);                 // [warps][D]
  load_tile(Ks, base + H * D, pitch, N, D, tid);
  load_tile(Vs, base + 2 * H * D, pitch, N, D, tid);
  __syncthreads();
  float* q = qs + warp * D;
  for (int i = warp; i < N; i += F32_WARPS) {
    for (int d = lane; d < D; d += 32) q[d] = base[static_cast<long>(i) * pitch + d];
    __syncwarp();
    float s[F32_MAXT];
    row_dots(q, Ks, N, D, lane, s);
    float mx = -CUDART_INF_F;
#pragma unroll
    for (int u = 0; u < F32_MAXT; ++u) if (lane + 32 * u < N) { s[u] *= scale; mx = fmaxf(mx, s[u]); }
    mx = warp_max32(mx);
    float l = 0.f;
#pragma unroll
    for (int u = 0; u < F32_MAXT; ++u) { s[u] = lane + 32 * u < N ? expf(s[u] - mx) : 0.f; l += s[u]; }
    l = warp_sum32(l);
    const float inv = 1.0f / l;
    float acc[3];
    weighted_rows(s, Vs, N, D, lane, acc);
    float* orow = o + (static_cast<long>(b) * N + i) * H * D + h * D;
    if (lane < D) orow[lane] = acc[0] * inv;
    if (lane + 32 < D) orow[lane + 32] = acc[1] * inv;
    if (lane + 64 < D) orow[lane + 64] = acc[2] * inv;
    if (lane == 0 && lse != nullptr) lse[(static_cast<long>(b) * H + h) * N + i] = mx + logf(l);
    __syncwarp();
  }
}

__global__ void __launch_bounds__(F32_THREADS)
attn_f32_bwd_kernel(const float* __restrict__ qkv, const float* __restrict__ o, const float* __restrict__ lse,
                    const float* __restrict__ d_o, float* __restrict__ dqkv, int B, int N, int H, int D, float scale) {
  extern __shared__ float sm[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const int C = H * D;
  const long pitch = 3L * C;
  const float* base = qkv + static_cast<long>(b) * N * pitch + h * D;
  const float* dob = d_o + static_cast<long>(b) * N * C + h * D;
  const float* ob = o + static_cast<long>(b) * N * C + h * D;
  float* gb = dqkv + static_cast<long>(b) * N * pitch + h * D;
  const int T = N * (D + 1);
  float* Ta = sm;                               // phase 1: K, phase 2: Q
  float* Tb = Ta + T;                           // phase 1: V, phase 2: dO
  float* s_lse = Tb + T;
  float* s_dl = s_lse + N;
  float* ra = s_dl + N + warp * 2 * D;          // this warp's two broadcast rows
  float* rb = ra + D;
  for (int i = tid; i < N; i += F32_THREADS) {
    s_lse[i] = lse[(static_cast<long>(b) * H + h) * N + i];
    float acc = 0.f;
    for (int d = 0; d < D; ++d) acc = fmaf(dob[static_cast<long>(i) * C + d], ob[static_cast<long>(i) * C + d], acc);
    s_dl[i] = acc;
  }
  load_tile(Ta, base + C, pitch, N, D, tid);
  load_tile(Tb, base + 2 * C, pitch, N, D, tid);
  __syncthreads();
  // pass 1: warp owns query i -> dQ_i = scale * sum_j dS_ij K_j
  for (int i = warp; i < N; i += F32_WARPS) {
    for (int d = lane; d < D; d += 32) { ra[d] = base[static_cast<long>(i) * pitch + d]; rb[d] = dob[static_cast<long>(i) * C + d]; }
    __syncwarp();
    float s[F32_MAXT], dp[F32_MAXT];
    row_dots(ra, Ta, N, D, lane, s);
    row_dots(rb, Tb, N, D, lane, dp);
    const float li = s_lse[i], di = s_dl[i];
#pragma unroll
    for (int u = 0; u < F32_MAXT; ++u) {
      const float p = lane + 32 * u < N ? expf(s[u] * scale - li) : 0.f;
      s[u] = p * (dp[u] - di);
    }
    float acc[3];
    weighted_rows(s, Ta, N, D, lane, acc);
    float* row = gb + static_cast<long>(i) * pitch;
    if (lane < D) row[lane] = acc[0] * scale;
    if (lane + 32 < D) row[lane + 32] = acc[1] * scale;
    if (lane + 64 < D) row[lane + 64] = acc[2] * scale;
    __syncwarp();
  }
  __syncthreads();
  load_tile(Ta, base, pitch, N, D, tid);
  load_tile(Tb, dob, C, N, D, tid);
  __syncthreads();
  // pass 2: warp owns key j -> dK_j = scale * sum_i dS_ij Q_i,  dV_j = sum_i P_ij dO_i   (lanes over the queries)
  for (int j = warp; j < N; j += F32_WARPS) {
    for (int d = lane; d < D; d += 32) {
      ra[d] = base[static_cast<long>(j) * pitch + C + d];
      rb[d] = base[static_cast<long>(j) * pitch + 2 * C + d];
    }
    __syncwarp();
    float s[F32_MAXT], dp[F32_MAXT];
    row_dots(ra, Ta, N, D, lane, s);
    row_dots(rb, Tb, N, D, lane, dp);
#pragma unroll
    for (int u = 0; u < F32_MAXT; ++u) {
      const int i = lane + 32 * u;
      const float p = i < N ? expf(s[u] * scale - s_lse[i]) : 0.f;
      dp[u] = i < N ? p * (dp[u] - s_dl[i]) : 0.f;
      s[u] = p;
    }
    float acc[3];
    weighted_rows(dp, Ta, N, D, lane, acc);
    float* row = gb + static_cast<long>(j) * pitch + C;
    if (lane < D) row[lane] = acc[0] * scale;
    if (lane + 32 < D) row[lane + 32] = acc[1] * scale;
    if (lane + 64 < D) row[lane + 64] = acc[2] * scale;
    weighted_rows(s, Tb, N, D, lane, acc);
    row += C;
    if (lane < D) row[lane] = acc[0];
    if (lane + 32 < D) row[lane + 32] = acc[1];
    if (lane + 64 < D) row[lane + 64] = acc[2];
    __syncwarp();
  }
}

}  // namespace

int gelu_f32_launch(const float* dy, const float* x, float* out, long n, cudaStream_t st) {
  if (n <= 0) return -80;
  int grid = static_cast<int>((n + 255) / 256);
  if (grid > 148 * 16) grid = 148 * 16;
  if (dy == nullptr) gelu_f32_fwd_kernel<<<grid, 256, 0, st>>>(x, out, n);
  else gelu_f32_bwd_kernel<<<grid, 256, 0, st>>>(dy, x, out, n);
  return cudaGetLastError() == cudaSuccess ? 0 : -81;
}

int attn_f32_launch(const float* qkv, float* o, float* lse, const float* d_o, float* dqkv, int B, int N, int H, int D,
                    float scale, cudaStream_t st) {
  if (B <= 0 || H <= 0 || N <= 0 || N > 32 * F32_MAXT || D <= 0 || D > 96) return -82;
  const bool bwd = d_o != nullptr;
  const int smem = (2 * N * (D + 1) + 2 * N + 2 * F32_WARPS * D) * 4;
  if (smem > 227 * 1024) return -83;
  static int cfg_f = 0, cfg_b = 0;
  if (bwd) {
    if (cfg_b < smem) {
      if (cudaFuncSetAttribute(attn_f32_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return -84;
      cfg_b = smem;
    }
    attn_f32_bwd_kernel<<<B * H, F32_THREADS, smem, st>>>(qkv, o, lse, d_o, dqkv, B, N, H, D, scale);
  } else {
    if (cfg_f < smem) {
      if (cudaFuncSetAttribute(attn_f32_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return -84;
      cfg_f = smem;
    }
    attn_f32_fwd_kernel<<<B * H, F32_THREADS, smem, st>>>(qkv, o, lse, B, N, H, D, scale);
  }
  return cudaGetLastError() == cudaSuccess ? 0 : -85;
}

}  // namespace cara
