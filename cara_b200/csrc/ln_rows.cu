// LayerNorm fused with the rank-R row contraction of the CP adapter it feeds.
//
// The K = C row passes of skinny.cu re-read a [M, C] activation that a LayerNorm kernel has just written:
//   forward : h = LN(x) is the input of qkv (norm1) / fc1 (norm2)  ->  T = h A, Uhat_s = cs_s (.) T   (cara.py:35,81)
//   backward: g_out = rowscale * dx_out is the incoming gradient G of proj (norm2's backward) / fc2 (the next
//             block's norm1 backward)  ->  dU = G B, dThat = cs (.) dU, dcs += sum_m dU (.) T          (autograd of :57,:92)
// Here a LayerNorm CTA leaves the bf16 rows it emits in a 16-row shared tile (two rows per warp), all eight warps
// contract the tile with the (hi, lo) factor (resident in shared memory, [2Rp, C], loaded ONCE per CTA: the CTAs are
// persistent over equal row ranges) by mma.sync with K split four ways, and two warps add the quarters in a fixed
// order and write T / Uhat (dThat) from registers while the others are already loading the next tile's rows.
// Deterministic; no HBM traffic beyond the [M, Rp] / [M, 3Rp] outputs, so the stand-alone pass (77 MB re-read + a
// launch per projection) disappears.  Same arithmetic as rows_kernel: bf16 rows x (hi, lo) factor, fp32 accumulation,
// outputs in the [hi | lo | hi] block layout of cara_gemm_cp's adapter segment.
//
// One warp per token row, 128-bit coalesced accesses, warp-shuffle statistics; the grid is CTAs-per-SM x SMs.
// Shared memory per CTA (C = 768, Rp = 16): 48 KB factor + 24 KB tile + 3 KB partial sums; two CTAs per SM (registers).
// (Measured dead ends, profiles/r02_ln_rows.md: one CTA per 16 rows re-loading the factor; a ninth, contraction-only
// warp behind named barriers; every warp contracting its own two rows in a 2-of-16-live MMA tile.)
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"

namespace cara {
namespace {

constexpr int LR_THREADS = 256;

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ uint32_t sm_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void cpa16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cpa_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cpa_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void ldsm4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pk2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float4 ld_bf16x4(const __nv_bfloat16* p) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const float2 fa = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
  const float2 fb = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}
// 4 floats -> 4 bf16 (8 bytes): to global and into the swizzled shared row (16-byte chunk index XOR row & 7)
__device__ __forceinline__ void st_bf16x4_both(__nv_bfloat16* g, uint32_t srow, int r, int c0, float4 v) {
  uint2 u;
  u.x = pk2(v.x, v.y);
  u.y = pk2(v.z, v.w);
  if (g != nullptr) *reinterpret_cast<uint2*>(g) = u;
  const uint32_t chunk = static_cast<uint32_t>(c0 >> 3), half = static_cast<uint32_t>(c0 & 4) << 1;
  asm volatile("st.shared.v2.b32 [%0], {%1,%2};" ::"r"(srow + ((chunk ^ static_cast<uint32_t>(r & 7)) << 4) + half), "r"(u.x),
               "r"(u.y) : "memory");
}

constexpr int LR_ROWS = 16;             // rows per tile = one m16 MMA tile, two rows per warp

// Shared layout (bytes): factor [2*RP rows][C] bf16 swizzled | tile [16 rows][C] bf16 swizzled | part [KQ-1][16][RP] fp32
template <int NV, int RT>
struct LrSmem {
  static constexpr int C = NV * 128, RP = RT * 8, PITCH = C * 2;
  static constexpr int KQ = 8 / RT;                              // K split: warp = (column group of 8, K quarter)
  static constexpr int KSTEPS = C / 16 / KQ;
  static constexpr int F_BYTES = 2 * RP * PITCH, H_BYTES = LR_ROWS * PITCH, P_BYTES = (KQ - 1) * LR_ROWS * RP * 4;
  static constexpr int TOTAL = F_BYTES + H_BYTES + P_BYTES;
};

// all threads: request the transposed (hi, lo) factor [2RP, C] into shared memory
template <int NV, int RT>
__device__ __forceinline__ void load_factor(uint32_t fs, const __nv_bfloat16* __restrict__ Ft) {
  using S = LrSmem<NV, RT>;
  constexpr int CH = S::C / 8;                                   // 16-byte chunks per row
  for (int idx = threadIdx.x; idx < 2 * S::RP * CH; idx += LR_THREADS) {
    const int row = idx / CH, ch = idx - row * CH;
    cpa16(fs + row * S::PITCH + ((ch ^ (row & 7)) << 4), Ft + static_cast<size_t>(row) * S::C + ch * 8);
  }
  cpa_commit();
}

// All eight warps, between the tile's two CTA-wide barriers: warp (cg = warp % RT, kq = warp / RT) contracts K quarter kq
// of the 16-row tile with the hi and lo rows of rank columns [8cg, 8cg + 8) and folds the two; quarters 1.. go to `part`.
// out (quarter 0 only): rows g = lane / 4 and g + 8, rank columns 8cg + 2 (lane % 4) + {0, 1}.
template <int NV, int RT>
__device__ __forceinline__ void contract_tile(uint32_t fs, uint32_t hs, float* part, float (&out)[4]) {
  using S = LrSmem<NV, RT>;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cg = warp % RT, kq = warp / RT;
  float hi[4] = {0.f, 0.f, 0.f, 0.f}, lo[4] = {0.f, 0.f, 0.f, 0.f};
  const int arow = lane & 15, a_hi = lane >> 4;
  const int mid = lane >> 3;
  const int bn = (mid >> 1) * S::RP + cg * 8 + (lane & 7), b_hi = mid & 1;
  const uint32_t a_base = hs + arow * S::PITCH, b_base = fs + bn * S::PITCH;
#pragma unroll
  for (int kk = 0; kk < S::KSTEPS; ++kk) {
    const int ch0 = (kq * S::KSTEPS + kk) * 2;
    uint32_t af[4], bf[4];
    ldsm4(a_base + (((ch0 + a_hi) ^ (arow & 7)) << 4), af);
    ldsm4(b_base + (((ch0 + b_hi) ^ (bn & 7)) << 4), bf);
    mma16816(hi, af, bf[0], bf[1]);
    mma16816(lo, af, bf[2], bf[3]);
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) out[e] = hi[e] + lo[e];
  if (kq > 0) {
    const int g = lane >> 2, col = cg * 8 + 2 * (lane & 3);
    float* p = part + (kq - 1) * LR_ROWS * S::RP;
    *reinterpret_cast<float2*>(p + g * S::RP + col) = make_float2(out[0], out[1]);
    *reinterpret_cast<float2*>(p + (g + 8) * S::RP + col) = make_float2(out[2], out[3]);
  }
}
// quarter-0 warps, after the second barrier: add the other quarters in a fixed order
template <int NV, int RT>
__device__ __forceinline__ void gather_tile(const float* part, float (&out)[4]) {
  using S = LrSmem<NV, RT>;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, col = (warp % RT) * 8 + 2 * (lane & 3);
#pragma unroll
  for (int q = 0; q < S::KQ - 1; ++q) {
    const float* p = part + q * LR_ROWS * S::RP;
    const float2 a = *reinterpret_cast<const float2*>(p + g * S::RP + col);
    const float2 b = *reinterpret_cast<const float2*>(p + (g + 8) * S::RP + col);
    out[0] += a.x; out[1] += a.y; out[2] += b.x; out[3] += b.y;
  }
}

// v (two adjacent rank columns of one row) as the bf16 column blocks [hi | lo | hi] at row pointer p
template <int RP>
__device__ __forceinline__ void emit_split(__nv_bfloat16* p, int col, float v0, float v1) {
  const __nv_bfloat162 hi = __floats2bfloat162_rn(v0, v1);
  const float2 hf = __bfloat1622float2(hi);
  const uint32_t h = *reinterpret_cast<const uint32_t*>(&hi);
  *reinterpret_cast<uint32_t*>(p + col) = h;
  *reinterpret_cast<uint32_t*>(p + RP + col) = pk2(v0 - hf.x, v1 - hf.y);
  *reinterpret_cast<uint32_t*>(p + 2 * RP + col) = h;
}

struct LnRowsFwd {
  const float* x_in; const __nv_bfloat16* delta; const float* rowscale; int rows_per_sample;
  float* x_out; const float* gamma; const float* beta; __nv_bfloat16* h; float* mean; float* rstd;
  int M; float eps;
  const __nv_bfloat16* Ft; const float* scales; int slices; float* T; __nv_bfloat16* U; long ldu;
  int rows_per_cta;
};

// Every CTA owns the same number of consecutive rows (the grid is exactly CTAs-per-SM x SMs: all SMs stream the same
// number of bytes) and walks over them in tiles of 16.  Per tile: each warp normalises two rows into the shared tile;
// barrier; all warps contract the tile (K split); barrier; two warps gather and write T / Uhat while the others are
// already loading the next tile's rows.
template <int NV, int RT>
__global__ void __launch_bounds__(LR_THREADS, 2)
ln_fwd_rows_kernel(const LnRowsFwd a) {
  using S = LrSmem<NV, RT>;
  constexpr int C = S::C, RP = S::RP;
  extern __shared__ __align__(128) uint8_t lr_smem[];
  const uint32_t fs = sm_u32(lr_smem), hs = fs + S::F_BYTES;
  float* part = reinterpret_cast<float*>(lr_smem + S::F_BYTES + S::H_BYTES);
  pdl_wait();
  pdl_trigger();
  load_factor<NV, RT>(fs, a.Ft);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_begin = blockIdx.x * a.rows_per_cta;
  const int m_end = min(a.M, m_begin + a.rows_per_cta);
  // Both rows of this warp are requested before either is used, and the NEXT tile's rows are requested as soon as this
  // tile's registers are free -- before the barriers and the contraction: a row waits ~5 us in the loaded HBM queues,
  // and a CTA that has nothing in flight while it contracts leaves the memory system idle for that long.
  float4 v[2][NV];
  uint2 dl[2][NV];
  auto request = [&](int m0) {
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const int row = m0 + warp * 2 + rr;
      const size_t base = static_cast<size_t>(row < m_end ? row : m_begin) * C;   // (dead rows: any valid address)
#pragma unroll
      for (int i = 0; i < NV; ++i) v[rr][i] = *reinterpret_cast<const float4*>(a.x_in + base + (i * 32 + lane) * 4);
      if (a.delta != nullptr) {
#pragma unroll
        for (int i = 0; i < NV; ++i) dl[rr][i] = *reinterpret_cast<const uint2*>(a.delta + base + (i * 32 + lane) * 4);
      }
    }
  };
  request(m_begin);
#pragma unroll 1
  for (int m0 = m_begin; m0 < m_end; m0 += LR_ROWS) {
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const int r = warp * 2 + rr, row = m0 + r;
      if (row >= m_end) continue;                                // (rows past the end keep stale, never emitted data)
      const uint32_t srow = hs + r * S::PITCH;
      const size_t base = static_cast<size_t>(row) * C;
      if (a.delta != nullptr) {
        const float rs = a.rowscale != nullptr ? a.rowscale[row / a.rows_per_sample] : 1.0f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const float2 da = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&dl[rr][i].x));
          const float2 db = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&dl[rr][i].y));
          v[rr][i].x = fmaf(rs, da.x, v[rr][i].x); v[rr][i].y = fmaf(rs, da.y, v[rr][i].y);
          v[rr][i].z = fmaf(rs, db.x, v[rr][i].z); v[rr][i].w = fmaf(rs, db.y, v[rr][i].w);
        }
        if (a.x_out != nullptr) {
#pragma unroll
          for (int i = 0; i < NV; ++i) *reinterpret_cast<float4*>(a.x_out + base + (i * 32 + lane) * 4) = v[rr][i];
        }
      }
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) s += (v[rr][i].x + v[rr][i].y) + (v[rr][i].z + v[rr][i].w);
      const float mean = wsum(s) * (1.0f / C);
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const float e0 = v[rr][i].x - mean, e1 = v[rr][i].y - mean, e2 = v[rr][i].z - mean, e3 = v[rr][i].w - mean;
        q += (e0 * e0 + e1 * e1) + (e2 * e2 + e3 * e3);
      }
      const float rstd = rsqrtf(wsum(q) * (1.0f / C) + a.eps);
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c0 = (i * 32 + lane) * 4;
        const float4 g = *reinterpret_cast<const float4*>(a.gamma + c0), b = *reinterpret_cast<const float4*>(a.beta + c0);
        float4 o;
        o.x = (v[rr][i].x - mean) * rstd * g.x + b.x; o.y = (v[rr][i].y - mean) * rstd * g.y + b.y;
        o.z = (v[rr][i].z - mean) * rstd * g.z + b.z; o.w = (v[rr][i].w - mean) * rstd * g.w + b.w;
        st_bf16x4_both(a.h + base + c0, srow, r, c0, o);
      }
      if (lane == 0 && a.mean != nullptr) { a.mean[row] = mean; a.rstd[row] = rstd; }
    }
    if (m0 + LR_ROWS < m_end) request(m0 + LR_ROWS);
    if (m0 == m_begin) cpa_wait_all();                           // first tile: this thread's share of the factor has landed
    __syncthreads();                                             // tile (and factor) complete; previous gather done with `part`
    float acc[4];
    contract_tile<NV, RT>(fs, hs, part, acc);
    __syncthreads();                                             // tile free for the next rows, `part` complete
    if (warp < RT) {
      // T (fp32, kept for backward) and the per-slice scaled operand of the GEMM's adapter segment
      gather_tile<NV, RT>(part, acc);
      const int r_lo = m0 + (lane >> 2), r_hi = r_lo + 8, col = warp * 8 + 2 * (lane & 3);
      if (a.T != nullptr) {
        if (r_lo < m_end) *reinterpret_cast<float2*>(a.T + static_cast<size_t>(r_lo) * RP + col) = make_float2(acc[0], acc[1]);
        if (r_hi < m_end) *reinterpret_cast<float2*>(a.T + static_cast<size_t>(r_hi) * RP + col) = make_float2(acc[2], acc[3]);
      }
      for (int so = 0; so < a.slices; ++so) {
        const float2 sc = *reinterpret_cast<const float2*>(a.scales + so * RP + col);
        if (r_lo < m_end) emit_split<RP>(a.U + static_cast<size_t>(r_lo) * a.ldu + so * 3 * RP, col, sc.x * acc[0], sc.y * acc[1]);
        if (r_hi < m_end) emit_split<RP>(a.U + static_cast<size_t>(r_hi) * a.ldu + so * 3 * RP, col, sc.x * acc[2], sc.y * acc[3]);
      }
    }
  }
}

struct LnRowsBwd {
  const __nv_bfloat16* dh; const float* x; const float* mean; const float* rstd; const float* gamma;
  const float* dx_in; float* dx_out; __nv_bfloat16* g_out; const float* rowscale; int rows_per_sample; int M;
  const __nv_bfloat16* Ft; const float* scales; const float* T; __nv_bfloat16* dT; float* dc;
  int rows_per_cta;
};

template <int NV, int RT>
__global__ void __launch_bounds__(LR_THREADS, 2)
ln_bwd_rows_kernel(const LnRowsBwd a) {
  using S = LrSmem<NV, RT>;
  constexpr int C = S::C, RP = S::RP;
  extern __shared__ __align__(128) uint8_t lr_smem[];
  const uint32_t fs = sm_u32(lr_smem), hs = fs + S::F_BYTES;
  float* part = reinterpret_cast<float*>(lr_smem + S::F_BYTES + S::H_BYTES);
  pdl_wait();
  pdl_trigger();
  load_factor<NV, RT>(fs, a.Ft);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_begin = blockIdx.x * a.rows_per_cta;
  const int m_end = min(a.M, m_begin + a.rows_per_cta);
  float dcs0 = 0.f, dcs1 = 0.f;                                  // quarter-0 warps: this lane's share of dcs = sum_rows dU (.) T
  // One row in flight per warp (registers); the first row of the NEXT tile is requested before the barriers and the
  // contraction, so the CTA never sits there with nothing in flight.
  uint2 dhr[NV];
  float4 xv[NV], rin[NV];
  float mean = 0.f, rstd = 0.f;
  auto request = [&](int row) {
    const size_t base = static_cast<size_t>(row < m_end ? row : m_begin) * C;     // (dead rows: any valid address)
    if (a.dx_in != nullptr) {
#pragma unroll
      for (int i = 0; i < NV; ++i) rin[i] = *reinterpret_cast<const float4*>(a.dx_in + base + (i * 32 + lane) * 4);
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      dhr[i] = *reinterpret_cast<const uint2*>(a.dh + base + (i * 32 + lane) * 4);
      xv[i] = *reinterpret_cast<const float4*>(a.x + base + (i * 32 + lane) * 4);
    }
    mean = a.mean[row < m_end ? row : m_begin];
    rstd = a.rstd[row < m_end ? row : m_begin];
  };
  request(m_begin + warp * 2);
#pragma unroll 1
  for (int m0 = m_begin; m0 < m_end; m0 += LR_ROWS) {
#pragma unroll 1
    for (int rr = 0; rr < 2; ++rr) {
      const int r = warp * 2 + rr, row = m0 + r;
      if (row < m_end) {
        const uint32_t srow = hs + r * S::PITCH;
        const size_t base = static_cast<size_t>(row) * C;
        float4 dy[NV];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const int c0 = (i * 32 + lane) * 4;
          const float4 g = *reinterpret_cast<const float4*>(a.gamma + c0);
          const float2 da = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&dhr[i].x));
          const float2 db = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&dhr[i].y));
          dy[i] = make_float4(da.x * g.x, da.y * g.y, db.x * g.z, db.y * g.w);
          xv[i] = make_float4((xv[i].x - mean) * rstd, (xv[i].y - mean) * rstd, (xv[i].z - mean) * rstd, (xv[i].w - mean) * rstd);
          s1 += (dy[i].x + dy[i].y) + (dy[i].z + dy[i].w);
          s2 += (dy[i].x * xv[i].x + dy[i].y * xv[i].y) + (dy[i].z * xv[i].z + dy[i].w * xv[i].w);
        }
        const float m1 = wsum(s1) * (1.0f / C), m2 = wsum(s2) * (1.0f / C);
        const float rs = a.rowscale != nullptr ? a.rowscale[row / a.rows_per_sample] : 1.0f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const int c0 = (i * 32 + lane) * 4;
          float4 o;
          o.x = rstd * (dy[i].x - m1 - xv[i].x * m2); o.y = rstd * (dy[i].y - m1 - xv[i].y * m2);
          o.z = rstd * (dy[i].z - m1 - xv[i].z * m2); o.w = rstd * (dy[i].w - m1 - xv[i].w * m2);
          if (a.dx_in != nullptr) { o.x += rin[i].x; o.y += rin[i].y; o.z += rin[i].z; o.w += rin[i].w; }
          *reinterpret_cast<float4*>(a.dx_out + base + c0) = o;
          st_bf16x4_both(a.g_out + base + c0, srow, r, c0, make_float4(o.x * rs, o.y * rs, o.z * rs, o.w * rs));
        }
      }
      // the next row of this warp: the tile's second row, then the first row of the next tile (in flight across the barriers)
      const int next = rr == 0 ? row + 1 : m0 + LR_ROWS + warp * 2;
      if (next < m_end) request(next);
    }
    // quarter-0 warps: the projection's saved T rows of this lane, requested before the barriers
    const int r_lo = m0 + (lane >> 2), r_hi = r_lo + 8, col = (warp % RT) * 8 + 2 * (lane & 3);
    float2 t_lo = make_float2(0.f, 0.f), t_hi = make_float2(0.f, 0.f);
    if (warp < RT) {
      if (r_lo < m_end) t_lo = *reinterpret_cast<const float2*>(a.T + static_cast<size_t>(r_lo) * RP + col);
      if (r_hi < m_end) t_hi = *reinterpret_cast<const float2*>(a.T + static_cast<size_t>(r_hi) * RP + col);
    }
    if (m0 == m_begin) cpa_wait_all();
    __syncthreads();
    float acc[4];                                                // dU = G B of the tile
    contract_tile<NV, RT>(fs, hs, part, acc);
    __syncthreads();
    if (warp < RT) {
      gather_tile<NV, RT>(part, acc);
      const float2 sc = *reinterpret_cast<const float2*>(a.scales + col);
      if (r_lo < m_end) {
        emit_split<RP>(a.dT + static_cast<size_t>(r_lo) * 3 * RP, col, sc.x * acc[0], sc.y * acc[1]);
        dcs0 = fmaf(acc[0], t_lo.x, dcs0);
        dcs1 = fmaf(acc[1], t_lo.y, dcs1);
      }
      if (r_hi < m_end) {
        emit_split<RP>(a.dT + static_cast<size_t>(r_hi) * 3 * RP, col, sc.x * acc[2], sc.y * acc[3]);
        dcs0 = fmaf(acc[2], t_hi.x, dcs0);
        dcs1 = fmaf(acc[3], t_hi.y, dcs1);
      }
    }
  }
  if (warp < RT) {                                               // rows of the tile (lane / 4), then one atomic per CTA and value
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) {
      dcs0 += __shfl_xor_sync(0xffffffffu, dcs0, o);
      dcs1 += __shfl_xor_sync(0xffffffffu, dcs1, o);
    }
    if (lane < 4) {
      atomicAdd(a.dc + warp * 8 + 2 * lane, dcs0);
      atomicAdd(a.dc + warp * 8 + 2 * lane + 1, dcs1);
    }
  }
}

inline int lr_sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}
// `per_sm` CTAs on every SM, every CTA the same share of the rows (rounded up to a whole row)
inline int lr_rows_per_cta(int M, int per_sm, int* grid) {
  const int ctas = per_sm * lr_sm_count();
  int rows = (M + ctas - 1) / ctas;
  if (rows < LR_ROWS) rows = LR_ROWS;
  *grid = (M + rows - 1) / rows;
  return rows;
}

template <int NV, int RT>
int fwd_rows_t(LnRowsFwd a, cudaStream_t st) {
  constexpr int smem = LrSmem<NV, RT>::TOTAL;
  static bool done = false;
  if (!done) {
    if (cudaFuncSetAttribute(ln_fwd_rows_kernel<NV, RT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return -23;
    done = true;
  }
  int grid = 0;
  a.rows_per_cta = lr_rows_per_cta(a.M, 2, &grid);
  return launch_pdl_f<4>(ln_fwd_rows_kernel<NV, RT>, dim3(grid), dim3(LR_THREADS), smem, st, a) == cudaSuccess ? 0 : -21;
}
template <int NV, int RT>
int bwd_rows_t(LnRowsBwd a, cudaStream_t st) {
  constexpr int smem = LrSmem<NV, RT>::TOTAL;
  static bool done = false;
  if (!done) {
    if (cudaFuncSetAttribute(ln_bwd_rows_kernel<NV, RT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return -23;
    done = true;
  }
  int grid = 0;
  a.rows_per_cta = lr_rows_per_cta(a.M, 2, &grid);
  return launch_pdl_f<4>(ln_bwd_rows_kernel<NV, RT>, dim3(grid), dim3(LR_THREADS), smem, st, a) == cudaSuccess ? 0 : -21;
}

}  // namespace

// Shapes with at least two CTAs per SM: (C, Rp) in {(768, 16), (1024, 16)}.  1 = not covered (caller keeps the
// stand-alone LayerNorm + rows pass), 0 = launched, < 0 = error.
int ln_rows_supported(int C, int rp) { return (rp == 16 && (C == 768 || C == 1024)) ? 1 : 0; }

int ln_fwd_rows_launch(const LnFwdArgs& l, const __nv_bfloat16* Ft, const float* scales, int slices, int rp, float* T,
                       __nv_bfloat16* U, cudaStream_t st) {
  if (l.M <= 0 || l.act_fp32 || l.h == nullptr || Ft == nullptr || scales == nullptr || U == nullptr || slices < 1) return -20;
  if (l.delta == nullptr && l.x_out != nullptr) return -20;
  if (!ln_rows_supported(l.C, rp)) return 1;
  LnRowsFwd a{l.x_in, static_cast<const __nv_bfloat16*>(l.delta), l.rowscale, l.rows_per_sample, l.x_out, l.gamma, l.beta,
              static_cast<__nv_bfloat16*>(l.h), l.mean, l.rstd, l.M, l.eps, Ft, scales, slices, T, U,
              static_cast<long>(slices) * 3 * rp, 0};
  return l.C == 768 ? fwd_rows_t<6, 2>(a, st) : fwd_rows_t<8, 2>(a, st);
}

int ln_bwd_rows_launch(const LnBwdArgs& l, const __nv_bfloat16* Ft, const float* scales, int rp, const float* T,
                       __nv_bfloat16* dT, float* dc, cudaStream_t st) {
  if (l.M <= 0 || l.act_fp32 || l.g_out == nullptr || Ft == nullptr || scales == nullptr || T == nullptr || dT == nullptr ||
      dc == nullptr)
    return -20;
  if (!ln_rows_supported(l.C, rp)) return 1;
  LnRowsBwd a{static_cast<const __nv_bfloat16*>(l.dh), l.x, l.mean, l.rstd, l.gamma, l.dx_in, l.dx_out,
              static_cast<__nv_bfloat16*>(l.g_out), l.rowscale, l.rows_per_sample, l.M, Ft, scales, T, dT, dc, 0};
  return l.C == 768 ? bwd_rows_t<6, 2>(a, st) : bwd_rows_t<8, 2>(a, st);
}

}  // namespace cara
