// Attention core of cp_attn (cara.py:44-48) on the 5th-generation tensor cores: softmax(q k^T D^-1/2) v for
// D = 64 and N <= 256 tokens (ViT-B/L @224/16: N = 197), one CTA per (sample, head), two CTAs per SM.
//
//   TMA      : the head's Q, K, V rows are pulled straight out of the fused projection's [B*N, 3*H*D] output into
//              128-byte-swizzled shared tiles (npad = ceil16(N) rows each).
//   tcgen05  : S = Q K^T (128 query rows x npad keys per instruction group, both operands K-major from smem) lands
//              in tensor memory; the softmax warps read their own row (thread = TMEM lane = query), write the
//              probabilities back into TMEM as packed bf16 ON TOP of the scores they replace, and
//              O = P V runs with A = P from TMEM and B = V from smem as an MN-major operand -- V is used exactly as
//              it lies in memory ([key][d]), nothing is transposed and P never touches shared memory.
//   epilogue : O / l, bf16 (hi, lo) pair + base-2 log-sum-exp, one full 128-byte line per thread and tensor.
//
// TMEM columns (256 per CTA): scores [0, npad), probabilities [0, npad/2), output accumulator [192, 256).
#include "gemm_sm100.h"
#include "kernels.h"
#include "ptx.cuh"

#include <math_constants.h>

namespace cara {
namespace {

constexpr int TC_THREADS = 160;          // warps 0-3: softmax / epilogue (TMEM lane quarter = warp), warp 4: TMA + MMA
constexpr int TC_TMEM_COLS = 256;
constexpr int TC_O_COL = 192;
constexpr int TC_D = 64;

__global__ void __launch_bounds__(TC_THREADS, 2)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap map_qkv, const AttnArgs a, const int npad) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t tile0 = (raw + 1023u) & ~1023u;
  const uint32_t tile_bytes = static_cast<uint32_t>(npad) * 128u;
  const uint32_t sQ = tile0, sK = tile0 + tile_bytes, sV = tile0 + 2u * tile_bytes;
  // barriers live behind the larger of (three tiles) and (Q base + 256 rows): the second query tile's MMA reads
  // 128 rows starting at row 128 whatever npad is
  const uint32_t span = 3u * tile_bytes > 32768u ? 3u * tile_bytes : 32768u;
  const uint32_t bars = tile0 + span;
  const uint32_t bar_qk = bars, bar_v = bars + 8, bar_s = bars + 16, bar_p = bars + 24, bar_o = bars + 32,
                 bar_od = bars + 40, tmem_slot = bars + 48;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x / a.H, h = blockIdx.x % a.H;
  const int N = a.N, C = a.H * TC_D;
  const int mtiles = (N + 127) / 128;

  if (warp == 4) {
    if (lane == 0) {
      tma_prefetch_desc(&map_qkv);
      mbar_init(bar_qk, 1);
      mbar_init(bar_v, 1);
      mbar_init(bar_s, 1);
      mbar_init(bar_p, 128);
      mbar_init(bar_o, 1);
      mbar_init(bar_od, 128);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc<TC_TMEM_COLS>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 4) {
    if (lane == 0) {
      const int row0 = b * N;
      mbar_expect_tx(bar_qk, 2u * tile_bytes);
      tma_load_2d(sQ, &map_qkv, bar_qk, h * TC_D, row0);
      tma_load_2d(sK, &map_qkv, bar_qk, C + h * TC_D, row0);
      mbar_expect_tx(bar_v, tile_bytes);
      tma_load_2d(sV, &map_qkv, bar_v, 2 * C + h * TC_D, row0);
      const uint32_t idesc_s = umma_idesc_bf16_major(128, npad, 0, 0);
      constexpr uint32_t idesc_o = umma_idesc_bf16_major(128, TC_D, 0, 1);
      mbar_wait(bar_qk, 0);
      for (int i = 0; i < mtiles; ++i) {
        if (i > 0) mbar_wait(bar_od, static_cast<uint32_t>(i - 1) & 1u);   // previous O has been read out of TMEM
        tc_fence_after();
        const uint64_t dq = umma_desc_sw128(sQ + static_cast<uint32_t>(i) * 16384u);
        const uint64_t dk = umma_desc_sw128(sK);
#pragma unroll
        for (int k = 0; k < TC_D / 16; ++k) umma_bf16(tmem_base, dq + 2u * k, dk + 2u * k, idesc_s, k != 0 ? 1u : 0u);
        umma_commit(bar_s);
        mbar_wait(bar_p, static_cast<uint32_t>(i) & 1u);                    // probabilities are in TMEM
        tc_fence_after();
        if (i == 0) mbar_wait(bar_v, 0);
        for (int k = 0; k < npad / 16; ++k) {
          const uint64_t dv = umma_desc_mn_sw128(sV + static_cast<uint32_t>(k) * 2048u, tile_bytes, 1024u);
          umma_bf16_ts(tmem_base + TC_O_COL, tmem_base + 8u * k, dv, idesc_o, k != 0 ? 1u : 0u);
        }
        umma_commit(bar_o);
      }
    }
  } else {
    const int t = warp * 32 + lane;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    const float sl2 = a.scale * 1.4426950408889634f;
    for (int i = 0; i < mtiles; ++i) {
      const int row = i * 128 + t;
      const bool warp_live = i * 128 + warp * 32 < N;          // warp-uniform: any valid query row in this warp
      mbar_wait(bar_s, static_cast<uint32_t>(i) & 1u);
      tc_fence_after();
      float mx = -CUDART_INF_F, l = 0.f;
      if (warp_live) {
        // pass 1: row maximum over the valid keys
        int c0 = 0;
        for (; c0 + 32 <= npad; c0 += 32) {
          uint32_t r[32];
          tmem_ld32(lane_addr + c0, r);
          tmem_ld_wait();
          if (c0 + 32 <= N) {
#pragma unroll
            for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(r[j]));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (c0 + j < N) mx = fmaxf(mx, __uint_as_float(r[j]));
          }
        }
        if (c0 < npad) {
          uint32_t r[16];
          tmem_ld16(lane_addr + c0, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) if (c0 + j < N) mx = fmaxf(mx, __uint_as_float(r[j]));
        }
        // pass 2: p = 2^(s*sl2 - max*sl2), row sum, packed bf16 back into TMEM (column c/2 <- keys c, c+1).
        // Chunk c0 of P overwrites score columns [c0/2, c0/2+16) which this thread has already consumed.
        const float msc = mx * sl2;
        for (c0 = 0; c0 + 32 <= npad; c0 += 32) {
          uint32_t r[32], pk[16];
          tmem_ld32(lane_addr + c0, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float p0 = exp2f(fmaf(__uint_as_float(r[2 * j]), sl2, -msc));
            float p1 = exp2f(fmaf(__uint_as_float(r[2 * j + 1]), sl2, -msc));
            if (c0 + 2 * j >= N) p0 = 0.f;
            if (c0 + 2 * j + 1 >= N) p1 = 0.f;
            l += p0 + p1;
            pk[j] = pack_bf16(p0, p1);
          }
          tmem_st16(lane_addr + (c0 >> 1), pk);
        }
        if (c0 < npad) {
          uint32_t r[16], pk[8];
          tmem_ld16(lane_addr + c0, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float p0 = exp2f(fmaf(__uint_as_float(r[2 * j]), sl2, -msc));
            float p1 = exp2f(fmaf(__uint_as_float(r[2 * j + 1]), sl2, -msc));
            if (c0 + 2 * j >= N) p0 = 0.f;
            if (c0 + 2 * j + 1 >= N) p1 = 0.f;
            l += p0 + p1;
            pk[j] = pack_bf16(p0, p1);
          }
          tmem_st8(lane_addr + (c0 >> 1), pk);
        }
        tmem_st_wait();
      }
      tc_fence_before();
      mbar_arrive(bar_p);
      mbar_wait(bar_o, static_cast<uint32_t>(i) & 1u);
      tc_fence_after();
      uint32_t o[64];
      if (warp_live) {
        tmem_ld32(lane_addr + TC_O_COL, *reinterpret_cast<uint32_t(*)[32]>(&o[0]));
        tmem_ld32(lane_addr + TC_O_COL + 32, *reinterpret_cast<uint32_t(*)[32]>(&o[32]));
        tmem_ld_wait();
      }
      tc_fence_before();
      mbar_arrive(bar_od);
      if (warp_live && row < N) {
        const float inv = 1.0f / l;
        const long off = (static_cast<long>(b) * N + row) * C + h * TC_D;
        uint4* po = reinterpret_cast<uint4*>(a.o + off);
        uint4* pl = a.o_lo != nullptr ? reinterpret_cast<uint4*>(a.o_lo + off) : nullptr;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float v[8];
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(o[8 * j + e]) * inv;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            hi[e] = pack_bf16(v[2 * e], v[2 * e + 1]);
            const float2 hf = unpack_bf16(hi[e]);
            lo[e] = pack_bf16(v[2 * e] - hf.x, v[2 * e + 1] - hf.y);
          }
          po[j] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          if (pl != nullptr) pl[j] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
        if (a.lse != nullptr) a.lse[(static_cast<long>(b) * a.H + h) * N + row] = mx * sl2 + log2f(l);
      }
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc<TC_TMEM_COLS>(tmem_base);
  }
}

}  // namespace

// D = 64, N <= 256: tcgen05 path.  Returns 1 when the shape is not covered (caller falls back to the mma.sync kernel).
int attn_fwd_tc_launch(const AttnArgs& a, cudaStream_t st) {
  if (a.D != TC_D || a.N > 256 || a.N < 1) return 1;
  const int npad = ((a.N + 15) / 16) * 16;
  const long rows = static_cast<long>(a.B) * a.N;
  const long cols = 3L * a.H * TC_D;
  CUtensorMap map;
  if (make_map_bf16(&map, a.qkv, rows, cols, cols, npad) != 0) return -54;
  const int tile_bytes = npad * 128;
  const int span = 3 * tile_bytes > 32768 ? 3 * tile_bytes : 32768;
  const int smem = 1024 + span + 64;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 1024 + 3 * 256 * 128 + 64) !=
        cudaSuccess)
      return -52;
    configured = true;
  }
  attn_fwd_tc_kernel<<<a.B * a.H, TC_THREADS, smem, st>>>(map, a, npad);
  return cudaGetLastError() == cudaSuccess ? 0 : -53;
}

}  // namespace cara
