// Fused CP-adapted projection GEMM for sm_100a (tcgen05 + TMEM + TMA), persistent, warp-specialised.
//
//   D[M,N] = A0[M,K0] * B0[N,K0]^T  (+)  A1[M, slice*Rp : +K1] * B1[N mod slice_w, K1]^T   + epilogue
//
// Segment 0 is the frozen weight GEMM (x * W^T, or g * W for dX with a pre-transposed W).
// Segment 1 is the rank-R Canonical-Polyadic side chain of SURVEY Appendix A.1/A.2: A1 holds the
// already scaled low-rank activations  s * (x*A) (.) c_slice  (or s * (g*B) (.) c summed over slices for
// dX) and B1 the out-side factor, so the adapter term is accumulated INTO THE SAME TMEM ACCUMULATOR as
// the frozen product -- the delta weight of cara.py:27-35,52-57,76-81,88-92 is never materialised and
// no second [M,N] pass exists.
//
// Roles: warp 0 = TMA producer (one elected lane), warp 1 = tcgen05.mma issuer + TMEM owner, then the
// epilogue warps: 4 for the plain epilogue, 8 (two per TMEM lane group, each taking half of the 256
// columns) for the GELU / GELU' epilogues whose per-element math would otherwise outlast the K = 768 main
// loop.  Two 256-column fp32 accumulators in TMEM let the epilogue of tile i overlap the main loop of tile
// i+1.  Epilogue data path: tcgen05.ld -> registers (bias / GELU / GELU') -> 128-byte-swizzled shared
// staging (conflict-free 16-byte accesses) -> TMA tensor store, 32 rows x 64 columns per warp and step, so HBM
// only ever sees full 128-byte lines.  Every kind stages through 32 KB (plain: two 4 KB buffers for each of 4
// warps; GELU kinds: one buffer for each of 8 warps, the two GELU outputs going through it one after the other)
// which leaves room for a 4-stage main-loop ring in all of them.  The GELU' operand (saved pre-activation) is read
// with plain 16-byte loads: one full 128-byte line per thread and step.
#include "ptx.cuh"
#include "gemm_sm100.h"
#include "kernels.h"
#include "mathfn.cuh"

#include <stdlib.h>

namespace cara {

constexpr int BM = 128, BN = 256, BK = 64;      // CTA tile; BK*2B = one 128-byte swizzle row
constexpr int UK = 16;                          // tcgen05 kind::f16 K per instruction
constexpr int A_BYTES = BM * BK * 2;            // 16 KB
constexpr int TMEM_COLS = 512;                  // 2 accumulators x 256 fp32 columns
// PAIR = two CTAs of a cluster (one TPC) run tcgen05.mma.cta_group::2 on a 256 x 256 tile: each CTA stages its own
// 128 rows of A and HALF of the B tile, so L2 -> SM traffic per flop drops by a third and the ring gets 6 stages
// instead of 4.  (The single-CTA kernel measured ~13-15 TB/s of L2 -> SM operand traffic: the L2 bound.)
__host__ __device__ constexpr int b_rows(bool pair) { return pair ? BN / 2 : BN; }
__host__ __device__ constexpr int stage_bytes(bool pair) { return A_BYTES + b_rows(pair) * BK * 2; }
constexpr int EC = 64;                          // epilogue step: 64 output columns = one 128-B swizzle row
constexpr int EBUF = 32 * EC * 2;               // 4 KB staging buffer: 32 rows x 128 B; two per epilogue warp
// main-loop ring depth / epilogue warps per epilogue kind (smem budget 227 KB)
// (measured: the GELU kinds are bound by main-loop ring depth, not by epilogue math -- 3 stages + 8 epilogue
// warps ran at 288 us where 4 stages + 4 warps ... see profiles/)
__host__ __device__ constexpr int num_stages(int epi, bool pair) { return pair ? 6 : 4; }
__host__ __device__ constexpr int num_epi_warps(int epi) { return epi == EPI_NONE ? 4 : 8; }
__host__ __device__ constexpr int epi_bufs(int epi) { return epi == EPI_NONE ? 2 : 1; }   // staging buffers per warp
__host__ __device__ constexpr int num_threads(int epi) { return 64 + 32 * num_epi_warps(epi); }
__host__ __device__ constexpr int gemm_smem(int epi, bool pair) {
  int stages = num_stages(epi, pair);
  while (stages * stage_bytes(pair) + num_epi_warps(epi) * epi_bufs(epi) * EBUF + 1024 + 512 > 227 * 1024) --stages;
  return stages * stage_bytes(pair) + num_epi_warps(epi) * epi_bufs(epi) * EBUF + 1024 /*align slack*/ + 512 /*barriers*/;
}
__host__ __device__ constexpr int fitted_stages(int epi, bool pair) {
  int stages = num_stages(epi, pair);
  while (stages * stage_bytes(pair) + num_epi_warps(epi) * epi_bufs(epi) * EBUF + 1024 + 512 > 227 * 1024) --stages;
  return stages;
}

struct TileCoord {
  int m0, n0;
};

template <int EPI, bool PAIR>
__global__ void __launch_bounds__(num_threads(EPI), 1)
gemm_cp_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapB0,
               const __grid_constant__ CUtensorMap mapA1, const __grid_constant__ CUtensorMap mapB1,
               const __grid_constant__ CUtensorMap mapOut, const __grid_constant__ CUtensorMap mapAux,
               const GemmArgs p) {
  constexpr int STAGES = fitted_stages(EPI, PAIR);
  constexpr int EW = num_epi_warps(EPI);
  constexpr int STAGE_BYTES = stage_bytes(PAIR);
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;     // 0 = leader of the CTA pair
  const int unit = PAIR ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int nunits = PAIR ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  // n fastest: the CTAs resident at one time share a handful of A row-panels and all of B through L2
  auto tile_coord = [&](int t) {
    return TileCoord{(t / p.tiles_n) * (PAIR ? 2 * BM : BM) + static_cast<int>(rank) * BM, (t % p.tiles_n) * BN};
  };
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t tiles = (raw + 1023u) & ~1023u;           // SWIZZLE_128B atoms need 1024-B alignment
  constexpr int EBUFS = epi_bufs(EPI);
  const uint32_t stage_out = tiles + STAGES * STAGE_BYTES; // [EW warps][EBUFS buffers][4 KB]
  const uint32_t bars = stage_out + EW * EBUFS * EBUF;
  // barrier map (8 B each): full[S], empty[S], tfull[2], tempty[2], (unused)[EW warps][2], then the TMEM base word
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bars + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (2 * STAGES + 2 + a); };
  auto aux_bar = [&](int w, int b) { return bars + 8u * (2 * STAGES + 4 + w * 2 + b); };
  const uint32_t tmem_slot = bars + 8u * (2 * STAGES + 4 + 2 * EW);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.tiles_m * p.tiles_n;
  const int ext_kblocks = (p.ksteps_ext + 3) / 4;
  const int kblocks = p.kblocks_main + ext_kblocks;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapA0);
    tma_prefetch_desc(&mapB0);
    if (ext_kblocks) {
      tma_prefetch_desc(&mapA1);
      tma_prefetch_desc(&mapB1);
    }
    for (int s = 0; s < STAGES; ++s) {
      // pair: only the leader arrives (expect_tx of BOTH CTAs' bytes); the peer's TMA loads complete_tx on the
      // leader's barrier directly and the peer cannot run a phase ahead (its stage is freed by the leader's commit)
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), PAIR ? 2 * EW : EW);  // one arrive per epilogue warp (of both CTAs)
    }
    for (int w = 0; w < EW; ++w) {
      mbar_init(aux_bar(w, 0), 1);
      mbar_init(aux_bar(w, 1), 1);
    }
    tma_prefetch_desc(&mapOut);
    if (EPI != EPI_NONE) tma_prefetch_desc(&mapAux);
    fence_mbar_init();
  }
  if (warp == 1) {
    if (PAIR) tmem_alloc_pair<TMEM_COLS>(tmem_slot); else tmem_alloc<TMEM_COLS>(tmem_slot);
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();      // barriers + TMEM visible (to the peer CTA as well)
  pdl_wait();                                              // nothing above touches global memory
  pdl_trigger();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      const int b_off = PAIR ? static_cast<int>(rank) * (BN / 2) : 0;   // pair: this CTA stages half of the B rows
      auto load = [&](uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
        if (PAIR) tma_load_2d_pair(dst, m, bar, c0, c1); else tma_load_2d(dst, m, bar, c0, c1);
      };
      for (int t = unit; t < num_tiles && !(p.debug & 2); t += nunits) {
        const TileCoord tc = tile_coord(t);
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(empty_bar(s), ph ^ 1u);       // own stage free (pair: the leader's commit is multicast to both)
          const uint32_t sa = tiles + s * STAGE_BYTES;
          const uint32_t sb = sa + A_BYTES;
          if (!PAIR) mbar_expect_tx(full_bar(s), STAGE_BYTES);
          else if (rank == 0) mbar_expect_tx(full_bar(s), 2 * STAGE_BYTES);
          if (kb < p.kblocks_main) {
            load(sa, &mapA0, full_bar(s), kb * BK, tc.m0);
            load(sb, &mapB0, full_bar(s), kb * BK, tc.n0 + b_off);
          } else {
            const int e = kb - p.kblocks_main;
            const int slice = tc.n0 / p.ext_slice_w;
            load(sa, &mapA1, full_bar(s), slice * p.ext_rp + e * BK, tc.m0);
            load(sb, &mapB1, full_bar(s), e * BK, tc.n0 - slice * p.ext_slice_w + b_off);
          }
          if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(PAIR ? 2 * BM : BM, BN);
      int s = 0;
      uint32_t ph = 0;
      int as = 0;
      uint32_t aph = 0;
      for (int t = unit; t < num_tiles; t += nunits) {
        mbar_wait(tempty_bar(as), aph ^ 1u);      // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * BN);
        for (int kb = 0; kb < kblocks; ++kb) {
          if (!(p.debug & 2)) mbar_wait(full_bar(s), ph);   // TMA bytes have landed
          tc_fence_after();
          const uint32_t sa = tiles + s * STAGE_BYTES;
          const uint64_t da = umma_desc_sw128(sa);
          const uint64_t db = umma_desc_sw128(sa + A_BYTES);
          int ksteps = BK / UK;
          if (kb >= p.kblocks_main) {
            const int left = p.ksteps_ext - 4 * (kb - p.kblocks_main);
            ksteps = left < 4 ? left : 4;
          }
          for (int k = 0; k < ksteps; ++k) {
            // advance 16 bf16 = 32 B inside the 128-B swizzle row: +2 in the (addr>>4) field
            if (PAIR) umma_bf16_pair(d_tmem, da + 2u * k, db + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
            else umma_bf16(d_tmem, da + 2u * k, db + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          // when these MMAs retire: free the smem stage (in both CTAs of a pair) / publish the accumulator
          if (PAIR) umma_commit_pair(empty_bar(s)); else umma_commit(empty_bar(s));
          if (kb == kblocks - 1) {
            if (PAIR) umma_commit_pair(tfull_bar(as)); else umma_commit(tfull_bar(as));
          }
          if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
        if (++as == 2) { as = 0; aph ^= 1u; }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps
    const int lg = warp & 3;                      // TMEM lane group this warp may touch
    const int ew = warp - 2;                      // staging slot of this warp
    constexpr int NCH = (BN / EC) * 4 / EW;       // 64-column steps per warp and tile (4, or 2 with 8 warps)
    const int c_first = (ew >> 2) * NCH;          // with 8 warps the second four take the upper 128 columns
    const uint32_t my_stage = stage_out + ew * (EBUFS * EBUF);
    const uint32_t sw = static_cast<uint32_t>(lane & 7);          // 128-B swizzle phase of this thread's row
    const uint32_t row_off = static_cast<uint32_t>(lane) * 128u;
    int as = 0;
    uint32_t aph = 0;
    uint32_t q = 0;                               // running step counter of this warp (buffer selector)
    auto stage_row = [&](uint32_t buf, const uint32_t (&w)[32]) { // this thread's 64 bf16 -> its swizzled staging row
#pragma unroll
      for (int j = 0; j < 8; ++j)
        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(buf + row_off + ((static_cast<uint32_t>(j) ^ sw) << 4)),
                     "r"(w[4 * j]), "r"(w[4 * j + 1]), "r"(w[4 * j + 2]), "r"(w[4 * j + 3]) : "memory");
    };
    for (int t = unit; t < num_tiles; t += nunits) {
      const TileCoord tc = tile_coord(t);
      const int grow = tc.m0 + lg * 32 + lane;    // global row of this thread
      uint4 ux[8];                                // GELU': this thread's 64 saved pre-activations of the current step
      if (EPI == EPI_DGELU) {
        const uint4* src = reinterpret_cast<const uint4*>(p.aux + static_cast<size_t>(grow < p.M ? grow : 0) * p.ldaux +
                                                          tc.n0 + c_first * EC);
#pragma unroll
        for (int j = 0; j < 4; ++j) ld_global_nc_v8(src + 2 * j, ux[2 * j], ux[2 * j + 1]);
      }
      mbar_wait(tfull_bar(as), aph);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(lg * 32) << 16) + static_cast<uint32_t>(as * BN);
#pragma unroll 1
      for (int ci = 0; ci < NCH; ++ci, ++q) {
        const int c = c_first + ci;
        const uint32_t buf = EBUFS == 2 ? my_stage + (q & 1u) * EBUF : my_stage;
        uint32_t r[64];
        tmem_ld32(t_row + c * EC, *reinterpret_cast<uint32_t(*)[32]>(&r[0]));
        tmem_ld32(t_row + c * EC + 32, *reinterpret_cast<uint32_t(*)[32]>(&r[32]));
        tmem_ld_wait();
        if (ci == NCH - 1) {                                      // accumulator fully read: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (PAIR && rank != 0) mbar_arrive_remote(tempty_bar(as), 0); else mbar_arrive(tempty_bar(as));
          }
        }
        if (p.debug & 1) continue;
        const int n = tc.n0 + c * EC;
        if (p.bias != nullptr && n < p.N) {
          const float4* b4 = reinterpret_cast<const float4*>(p.bias + n);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float4 bv = __ldg(b4 + j);
            r[4 * j + 0] = __float_as_uint(__uint_as_float(r[4 * j + 0]) + bv.x);
            r[4 * j + 1] = __float_as_uint(__uint_as_float(r[4 * j + 1]) + bv.y);
            r[4 * j + 2] = __float_as_uint(__uint_as_float(r[4 * j + 2]) + bv.z);
            r[4 * j + 3] = __float_as_uint(__uint_as_float(r[4 * j + 3]) + bv.w);
          }
        }
        uint32_t w0[32];                                          // first (or only) output of this step, packed bf16
        if (EPI == EPI_DGELU) {
          // dX through GELU: multiply by gelu'(u), u = saved fc1 pre-activation
          const uint32_t* uw = reinterpret_cast<const uint32_t*>(ux);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float2 f = unpack_bf16(uw[j]);
            w0[j] = (p.debug & 8) ? pack_bf16(__uint_as_float(r[2 * j]) * f.x, __uint_as_float(r[2 * j + 1]) * f.y)
                                  : pack_bf16(__uint_as_float(r[2 * j]) * gelu_grad_fast(f.x), __uint_as_float(r[2 * j + 1]) * gelu_grad_fast(f.y));
          }
          if (ci + 1 < NCH) {                                     // next step's operand: in flight during the staging below
            const uint4* src = reinterpret_cast<const uint4*>(p.aux + static_cast<size_t>(grow < p.M ? grow : 0) * p.ldaux +
                                                              n + EC);
#pragma unroll
            for (int j = 0; j < 4; ++j) ld_global_nc_v8(src + 2 * j, ux[2 * j], ux[2 * j + 1]);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) w0[j] = pack_bf16(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
        }
        // earlier TMA stores must have finished READING the buffer this step writes
        if (lane == 0) {
          if (EBUFS == 2) tma_store_wait_read<1>(); else tma_store_wait_read<0>();
        }
        __syncwarp();
        if ((EPI != EPI_GELU || p.out != nullptr) && !(p.debug & 4)) {
          stage_row(buf, w0);
          fence_proxy_async();                                    // staging writes -> visible to the TMA engine
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&mapOut, buf, n, tc.m0 + lg * 32);
            tma_store_commit();
          }
        }
        if (EPI == EPI_GELU) {
          // fc1: `out` keeps the pre-activation for backward, `out2` gets GELU(u) for fc2.  GELU is applied to the
          // bf16-rounded pre-activation so forward and backward see the same u.  The math runs while the TMA engine
          // is still reading the pre-activation out of the (single) staging buffer.
          uint32_t w1[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float2 f = unpack_bf16(w0[j]);
            w1[j] = (p.debug & 8) ? pack_bf16(f.x + 1.0f, f.y + 1.0f) : pack_bf16(gelu_fast(f.x), gelu_fast(f.y));
          }
          if (lane == 0) tma_store_wait_read<0>();
          __syncwarp();
          stage_row(buf, w1);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&mapAux, buf, n, tc.m0 + lg * 32);
            tma_store_commit();
          }
        }
      }
      if (++as == 2) { as = 0; aph ^= 1u; }
    }
    if (lane == 0) tma_store_wait<0>();                           // all output bytes are in global memory
  }

  __syncwarp();                                            // reconverge the single-lane role warps
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();      // pair: the peer may still be reading / being signalled
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_pair<TMEM_COLS>(tmem_base); else tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

// ------------------------------------------------------------------------------ host side

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// Row-major bf16 [rows, cols] with row stride ld (elements); box = [box_rows, 64 cols], 128-B swizzle.
int make_map_bf16(CUtensorMap* map, const void* base, long rows, long cols, long ld, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (fn == nullptr) return -1;
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || ((ld * 2) & 15) != 0) return -2;
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(BK), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -3;
}

template <int EPI, bool PAIR>
static cudaError_t launch_epi(const CUtensorMap& a0, const CUtensorMap& b0, const CUtensorMap& a1,
                              const CUtensorMap& b1, const CUtensorMap& mo, const CUtensorMap& mx,
                              const GemmArgs& args, int grid, cudaStream_t st) {
  static bool attr_done = false;
  constexpr int smem = gemm_smem(EPI, PAIR);
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(gemm_cp_kernel<EPI, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(num_threads(EPI));
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = PAIR ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (pdl_mask() & 1) ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, gemm_cp_kernel<EPI, PAIR>, a0, b0, a1, b1, mo, mx, args);
}

template <bool PAIR>
static cudaError_t launch_kind(int epi, const CUtensorMap& a0, const CUtensorMap& b0, const CUtensorMap& a1,
                               const CUtensorMap& b1, const CUtensorMap& mo, const CUtensorMap& mx,
                               const GemmArgs& args, int grid, cudaStream_t st) {
  switch (epi) {
    case EPI_NONE: return launch_epi<EPI_NONE, PAIR>(a0, b0, a1, b1, mo, mx, args, grid, st);
    case EPI_GELU: return launch_epi<EPI_GELU, PAIR>(a0, b0, a1, b1, mo, mx, args, grid, st);
    default: return launch_epi<EPI_DGELU, PAIR>(a0, b0, a1, b1, mo, mx, args, grid, st);
  }
}

int gemm_cp_launch(const GemmDesc& d, cudaStream_t st) {
  if (d.M <= 0 || d.N <= 0 || d.K0 <= 0 || (d.K0 % 8) != 0 || (d.N % 64) != 0) return -10;
  if (d.out == nullptr && d.epi != EPI_GELU) return -16;
  CUtensorMap a0, b0, a1, b1, mo, mx;
  int rc;
  const bool pair = d.pair != 0;
  if ((rc = make_map_bf16(&a0, d.A0, d.M, d.K0, d.lda0, BM)) != 0) return rc * 10 - 1;
  if ((rc = make_map_bf16(&b0, d.B0, d.N, d.K0, d.ldb0, b_rows(pair))) != 0) return rc * 10 - 2;
  GemmArgs args{};
  args.M = d.M; args.N = d.N;
  args.kblocks_main = (d.K0 + BK - 1) / BK;
  args.ksteps_ext = 0;
  args.ext_slice_w = d.N; args.ext_rp = 0;
  if (d.A1 != nullptr) {
    if (d.K1 <= 0 || (d.K1 % 16) != 0 || d.ext_slices < 1 || (d.N % d.ext_slices) != 0) return -11;
    const int slice_w = d.N / d.ext_slices;
    if (d.ext_slices > 1 && (slice_w % BN) != 0) return -12;
    if ((rc = make_map_bf16(&a1, d.A1, d.M, static_cast<long>(d.K1) * d.ext_slices, d.lda1, BM)) != 0) return rc * 10 - 3;
    if ((rc = make_map_bf16(&b1, d.B1, slice_w, d.K1, d.ldb1, b_rows(pair))) != 0) return rc * 10 - 4;
    args.ksteps_ext = d.K1 / UK;
    args.ext_slice_w = slice_w;
    args.ext_rp = d.K1;
  } else {
    a1 = a0; b1 = b0;
  }
  args.bias = d.bias;
  args.out = d.out; args.ldo = d.ldo;
  args.out2 = d.out2; args.ldo2 = d.ldo2;
  args.aux = d.aux; args.ldaux = d.ldaux;
  {
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("CARA_GEMM_DEBUG"); dbg = e != nullptr ? atoi(e) : 0; }
    args.debug = dbg;
  }
  const int tile_m = pair ? 2 * BM : BM;
  args.tiles_m = (d.M + tile_m - 1) / tile_m;
  args.tiles_n = (d.N + BN - 1) / BN;
  const int num_tiles = args.tiles_m * args.tiles_n;
  int grid = d.num_sms > 0 ? d.num_sms : 148;
  if (pair) {
    grid &= ~1;
    if (grid > 2 * num_tiles) grid = 2 * num_tiles;
  } else if (grid > num_tiles) {
    grid = num_tiles;
  }
  // epilogue tensors: 32-row x 64-column boxes (one per epilogue warp and step)
  if (d.out != nullptr) {
    if ((rc = make_map_bf16(&mo, d.out, d.M, d.N, d.ldo, 32)) != 0) return rc * 10 - 5;
  } else {
    mo = a0;
  }
  mx = a0;
  switch (d.epi) {
    case EPI_NONE: break;
    case EPI_GELU:
      if (d.out2 == nullptr) return -13;
      if ((rc = make_map_bf16(&mx, d.out2, d.M, d.N, d.ldo2, 32)) != 0) return rc * 10 - 6;
      break;
    case EPI_DGELU:
      if (d.aux == nullptr) return -14;
      if ((rc = make_map_bf16(&mx, d.aux, d.M, d.N, d.ldaux, 32)) != 0) return rc * 10 - 7;
      break;
    default: return -15;
  }
  const cudaError_t e = pair ? launch_kind<true>(d.epi, a0, b0, a1, b1, mo, mx, args, grid, st)
                             : launch_kind<false>(d.epi, a0, b0, a1, b1, mo, mx, args, grid, st);
  return e == cudaSuccess ? 0 : -1000 - static_cast<int>(e);
}

}  // namespace cara
