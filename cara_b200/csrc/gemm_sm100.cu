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
// SIDE TILES (the rank-R row contraction, inside this kernel).  A1 itself is a product of the SAME A0 rows with a
// [K0, R] factor: forward T = x*A, Uhat_s = cs_s (.) T; backward dU_s = g_s*B, dThat = sum_s cs_s (.) dU_s,
// dcs_s = sum_m dU_s (.) T.  With side != 0 every 128-row panel gets one extra "side tile" ahead of its output tiles:
// the same TMA loads of the A0 panel (served from L2 -- the panel's output tiles are streaming it at that moment),
// tcgen05.mma 128 x 2Rp x 16 against the [hi ; lo] rows of the transposed factor into a few columns of the tile's TMEM
// accumulator buffer, and an epilogue that folds (hi, lo), scales, emits the [hi | lo | hi] bf16 operand rows (and T /
// the dcs partial sums) and then publishes the panel: each of the four epilogue warps st.releases the launch's generation
// number into its word of sync[2 + 4 * panel ..].
// The producer of an output tile acquires that flag (and crosses to the async proxy) right before it issues the
// TMA loads of the adapter segment, i.e. at the very end of its K loop, so the wait is normally already satisfied.
// Tiles are handed out round-robin in increasing order and a side tile precedes the output tiles that need it, so the
// dependency graph follows the tile order and cannot deadlock as long as all CTAs are resident (persistent grid of at
// most one CTA per SM); the grid size is chosen coprime to the tiles-per-panel count so that side tiles are spread
// over all CTAs.  The side tile of a panel is issued one full round of the grid AHEAD of its consumers (decode_tile)
// and stages two k-blocks per ring slot, so the flag is up long before anybody asks.  sync[0] is the generation of the previous launch and sync[1] an exit ticket: the last CTA to leave
// advances the generation, so the flags never need clearing -- also not between replays of a CUDA graph.
// This replaces a separate pass over x / g per projection (the rows kernel of round 1: 96 launches and 14 GB of HBM
// reads per ViT-B step).
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
constexpr int STAGE_BYTES = A_BYTES + BN * BK * 2;   // 48 KB
constexpr int STAGES = 4;
constexpr int TMEM_COLS = 512;                  // 2 accumulators x 256 fp32 columns
constexpr int EC = 64;                          // epilogue step: 64 output columns = one 128-B swizzle row
constexpr int EBUF = 32 * EC * 2;               // 4 KB staging buffer: 32 rows x 128 B; two per epilogue warp
constexpr int SIDE_RED = 4 * 32;                // dcs partial sums of this CTA (slices x rank)
__host__ __device__ constexpr bool light_epi(int epi) { return epi == EPI_NONE || epi == EPI_DELTA; }
__host__ __device__ constexpr int num_epi_warps(int epi) { return light_epi(epi) ? 4 : 8; }
__host__ __device__ constexpr int epi_bufs(int epi) { return light_epi(epi) ? 2 : 1; }   // staging buffers per warp
// SW: four extra warps that drain the side tiles (plain epilogue kind only: the 8-warp GELU kinds are at the register
// ceiling of 384 threads and keep the stand-alone rows pass)
__host__ __device__ constexpr int num_threads(int epi, bool sw = false) { return 64 + 32 * num_epi_warps(epi) + (sw ? 128 : 0); }
__host__ __device__ constexpr int gemm_smem(int epi) {
  return STAGES * STAGE_BYTES + num_epi_warps(epi) * epi_bufs(epi) * EBUF + 1024 /*align slack*/ + 512 /*barriers*/ +
         SIDE_RED * 4;
}
static_assert(gemm_smem(EPI_NONE) <= 227 * 1024 && gemm_smem(EPI_GELU) <= 227 * 1024, "shared memory budget");

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// The side tile of a panel has been written by (generic-proxy) stores of another CTA, one flag word per TMEM lane
// group (each epilogue warp publishes its own 32 rows): acquire the four flags, then order the TMA loads (async proxy)
// of this thread after them.  Watchdog as for the mbarriers: trap instead of hanging.
__device__ __forceinline__ bool panel_ready(const unsigned* flags, unsigned gen) {
  return ld_acquire_u32(flags) == gen && ld_acquire_u32(flags + 1) == gen && ld_acquire_u32(flags + 2) == gen &&
         ld_acquire_u32(flags + 3) == gen;
}
__device__ __forceinline__ void wait_panel(const unsigned* flags, unsigned gen) {
  if (!panel_ready(flags, gen)) {
    const long long t0 = clock64();
    while (!panel_ready(flags, gen)) {
      __nanosleep(64);
      if (clock64() - t0 > 4000000000LL) {
        printf("cara_b200: side-tile flag watchdog (block %d, want %u, have %u %u %u %u)\n", blockIdx.x, gen,
               ld_acquire_u32(flags), ld_acquire_u32(flags + 1), ld_acquire_u32(flags + 2), ld_acquire_u32(flags + 3));
        __trap();
      }
    }
  }
  asm volatile("fence.proxy.async.global;" ::: "memory");
}
// The two TMEM accumulator buffers.  An OUTPUT tile takes buffer `cur` and flips it; uses of a buffer are counted and give
// the mbarrier phase parities.  A SIDE tile takes no turn in that rotation: it borrows `cur` -- the buffer the next
// output tile of this CTA will use, already drained -- for its few columns, four dedicated warps drain it the moment
// its MMAs retire (sfull / sempty barriers), and the MMA warp then starts the next output tile in the same buffer
// while the output-tile epilogue warps are still busy with the previous tile in the other one.  (Round-2 history: with
// a turn in the rotation every side tile cost a whole un-overlapped epilogue; draining it from the output-tile
// epilogue warps between their steps pushed the 168-register GELU' epilogue into spills -- profiles/r02_side_tiles.md.)
struct AccState {
  int cur;
  uint32_t u0, u1;
  __device__ __forceinline__ uint32_t parity(int b) const { return (b ? u1 : u0) & 1u; }
  __device__ __forceinline__ void advance() {
    if (cur) ++u1; else ++u0;
    cur ^= 1;
  }
};

// Tile order.  Without side tiles: tile t = (panel t / tiles_n, column tile t % tiles_n).  With side tiles the sequence is
//   S_0 .. S_{la-1},  then for every panel p:  S_{p+la},  M_{p,0} .. M_{p,tiles_n-1}
// i.e. the side tile of a panel is issued `la` panels AHEAD of the output tiles that consume it, with la chosen by the
// host so that this is at least one full round of the persistent grid (la >= grid / tiles-per-panel).  A side tile
// queues behind the previous tile of its CTA and needs its own TMA round trips; issued in the same round as its
// consumers it finished after they wanted it and every output tile stalled behind the flag (measured: 2x).  One round
// of lookahead hides it completely, and the A0 panel it pulls from HBM is still in L2 a round later (one round touches
// ~10 MB of distinct operand bytes).
enum TileKind { TILE_SKIP = 0, TILE_SIDE = 1, TILE_MAIN = 2 };
struct TileInfo { int kind, panel, n0; };
__device__ __forceinline__ TileInfo decode_tile(int t, int tiles_m, int tiles_n, bool side, int la) {
  if (!side) return TileInfo{TILE_MAIN, t / tiles_n, (t % tiles_n) * BN};
  if (tiles_n == 0) return TileInfo{TILE_SIDE, t, 0};
  if (t < la) return TileInfo{t < tiles_m ? TILE_SIDE : TILE_SKIP, t, 0};
  const int u = t - la, tpp = tiles_n + 1;
  const int grp = u / tpp, pos = u - grp * tpp;
  if (pos == 0) return TileInfo{grp + la < tiles_m ? TILE_SIDE : TILE_SKIP, grp + la, 0};
  return TileInfo{TILE_MAIN, grp, (pos - 1) * BN};
}

// v[0 .. RP) per lane -> column sums over the 32 lanes; afterwards lane l holds the total of column
// (RP == 32 ? l : l >> 1) in v[0] (RP == 16: in both lanes of a pair).  RP/2 + ... + 1 (+ 1) shuffles.
template <int RP>
__device__ __forceinline__ void warp_colsum(float (&v)[RP], int lane) {
  int off = 16;
#pragma unroll
  for (int half = RP / 2; half >= 1; half >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const float send = up ? v[i] : v[i + half];
      const float keep = up ? v[i + half] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
    off >>= 1;
  }
  if (RP == 16) v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
}

// this thread's row of a low-rank activation as the bf16 column blocks [hi | lo | hi] (block pitch RP)
template <int RP>
__device__ __forceinline__ void emit_split_row(__nv_bfloat16* dst, const float (&v)[RP]) {
  uint32_t hi[RP / 2], lo[RP / 2];
#pragma unroll
  for (int i = 0; i < RP / 2; ++i) {
    hi[i] = pack_bf16(v[2 * i], v[2 * i + 1]);
    const float2 hf = unpack_bf16(hi[i]);
    lo[i] = pack_bf16(v[2 * i] - hf.x, v[2 * i + 1] - hf.y);
  }
  uint4* d = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int i = 0; i < RP / 8; ++i) {
    const uint4 h = make_uint4(hi[4 * i], hi[4 * i + 1], hi[4 * i + 2], hi[4 * i + 3]);
    d[i] = h;
    d[RP / 8 + i] = make_uint4(lo[4 * i], lo[4 * i + 1], lo[4 * i + 2], lo[4 * i + 3]);
    d[2 * (RP / 8) + i] = h;
  }
}

// Side-tile epilogue of one warp (32 rows = its TMEM lane group).  `t_row` = TMEM address of the warp's lanes at the
// accumulator buffer's first column; slice s occupies columns [s * 2RP, (s+1) * 2RP) = [hi-factor part | lo-factor part].
// `release` is called once, right after the last TMEM read.
template <int RP>
__device__ __noinline__ void side_epilogue(const GemmArgs& p, uint32_t t_row, int grow, int lane, float* side_red,
                                           uint32_t tempty) {
  // (not inlined: keeps the register allocation of the output tiles' epilogue loop as it was)
  auto release = [&]() {                        // accumulator buffer read: hand it back to the MMA warp
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(tempty);
  };
  const bool ok = grow < p.M;
  if (p.side == SIDE_FWD) {
    uint32_t r[2 * RP];
    tmem_ld32(t_row, *reinterpret_cast<uint32_t(*)[32]>(&r[0]));
    if constexpr (RP == 32) tmem_ld32(t_row + 32, *reinterpret_cast<uint32_t(*)[32]>(&r[32]));
    tmem_ld_wait();
    release();
    float T[RP];
#pragma unroll
    for (int i = 0; i < RP; ++i) T[i] = __uint_as_float(r[i]) + __uint_as_float(r[RP + i]);
    if (!ok) return;
    if (p.side_T != nullptr) {
      float4* t4 = reinterpret_cast<float4*>(p.side_T + static_cast<size_t>(grow) * RP);
#pragma unroll
      for (int i = 0; i < RP / 4; ++i) t4[i] = make_float4(T[4 * i], T[4 * i + 1], T[4 * i + 2], T[4 * i + 3]);
    }
    for (int s = 0; s < p.side_slices; ++s) {
      const float4* sc4 = reinterpret_cast<const float4*>(p.side_scales + s * RP);
      float u[RP];
#pragma unroll
      for (int i = 0; i < RP / 4; ++i) {
        const float4 c = __ldg(sc4 + i);
        u[4 * i] = c.x * T[4 * i]; u[4 * i + 1] = c.y * T[4 * i + 1];
        u[4 * i + 2] = c.z * T[4 * i + 2]; u[4 * i + 3] = c.w * T[4 * i + 3];
      }
      emit_split_row<RP>(p.side_U + static_cast<size_t>(grow) * p.side_ldu + s * 3 * RP, u);
    }
  } else {
    float T[RP], d[RP];
    {
      const float4* t4 = reinterpret_cast<const float4*>(p.side_T + static_cast<size_t>(ok ? grow : 0) * RP);
#pragma unroll
      for (int i = 0; i < RP / 4; ++i) {
        const float4 t = ok ? __ldg(t4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        T[4 * i] = t.x; T[4 * i + 1] = t.y; T[4 * i + 2] = t.z; T[4 * i + 3] = t.w;
      }
    }
#pragma unroll
    for (int i = 0; i < RP; ++i) d[i] = 0.f;
    for (int s = 0; s < p.side_slices; ++s) {
      uint32_t r[2 * RP];
      tmem_ld32(t_row + s * 2 * RP, *reinterpret_cast<uint32_t(*)[32]>(&r[0]));
      if constexpr (RP == 32) tmem_ld32(t_row + s * 2 * RP + 32, *reinterpret_cast<uint32_t(*)[32]>(&r[32]));
      tmem_ld_wait();
      if (s == p.side_slices - 1) release();
      const float4* sc4 = reinterpret_cast<const float4*>(p.side_scales + s * RP);
      float pr[RP];
#pragma unroll
      for (int i = 0; i < RP / 4; ++i) {
        const float4 c = __ldg(sc4 + i);
        const float cc[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float du = __uint_as_float(r[4 * i + e]) + __uint_as_float(r[RP + 4 * i + e]);
          d[4 * i + e] = fmaf(cc[e], du, d[4 * i + e]);
          pr[4 * i + e] = du * T[4 * i + e];          // rows >= M: the TMA zero fill makes du = 0
        }
      }
      warp_colsum<RP>(pr, lane);
      if (RP == 32) atomicAdd(side_red + s * RP + lane, pr[0]);
      else if ((lane & 1) == 0) atomicAdd(side_red + s * RP + (lane >> 1), pr[0]);
    }
    if (ok) emit_split_row<RP>(p.side_U + static_cast<size_t>(grow) * p.side_ldu, d);
  }
}

template <int EPI, bool SW>
__global__ void __launch_bounds__(num_threads(EPI, SW), 1)
gemm_cp_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapB0,
               const __grid_constant__ CUtensorMap mapA1, const __grid_constant__ CUtensorMap mapB1,
               const __grid_constant__ CUtensorMap mapOut, const __grid_constant__ CUtensorMap mapAux,
               const __grid_constant__ CUtensorMap mapP, const GemmArgs p) {
  constexpr int EW = num_epi_warps(EPI);
  const int unit = static_cast<int>(blockIdx.x);
  const int nunits = static_cast<int>(gridDim.x);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t tiles = (raw + 1023u) & ~1023u;           // SWIZZLE_128B atoms need 1024-B alignment
  constexpr int EBUFS = epi_bufs(EPI);
  const uint32_t stage_out = tiles + STAGES * STAGE_BYTES; // [EW warps][EBUFS buffers][4 KB]
  const uint32_t bars = stage_out + EW * EBUFS * EBUF;
  // barrier map (8 B each): full[S], empty[S], tfull[2], tempty[2], then the TMEM base word, the generation word and
  // (at +512) the dcs partial sums
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bars + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (2 * STAGES + 2 + a); };
  const uint32_t sfull_bar = bars + 8u * (2 * STAGES + 4), sempty_bar = bars + 8u * (2 * STAGES + 5);
  const uint32_t tmem_slot = bars + 8u * (2 * STAGES + 6);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));
  volatile uint32_t* gen_slot_ptr = tmem_slot_ptr + 1;
  float* side_red = reinterpret_cast<float*>(smem_raw + (bars + 512u - raw));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int side_tiles = p.side != SIDE_NONE ? 1 : 0;
  const int num_tiles = p.tiles_m * (p.tiles_n + side_tiles) + (side_tiles && p.tiles_n > 0 ? p.side_la : 0);
  const int side_fills = (p.kblocks_main + 1) / 2;         // a side tile stages TWO k-blocks per ring slot

  const int ext_kblocks = (p.ksteps_ext + 3) / 4;
  const int kblocks = p.kblocks_main + ext_kblocks;
  const int side_n = 2 * p.side_rp;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapA0);
    if (p.tiles_n > 0) tma_prefetch_desc(&mapB0);
    if (ext_kblocks) {
      tma_prefetch_desc(&mapA1);
      tma_prefetch_desc(&mapB1);
    }
    if (side_tiles) tma_prefetch_desc(&mapP);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), EW);               // one arrive per epilogue warp
    }
    mbar_init(sfull_bar, 1);
    mbar_init(sempty_bar, 4);                     // one arrive per side-drain warp
    if (p.tiles_n > 0) tma_prefetch_desc(&mapOut);
    if (EPI == EPI_GELU || EPI == EPI_DGELU) tma_prefetch_desc(&mapAux);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
  pdl_wait();                                              // nothing above touches global memory
  pdl_trigger();
  if (side_tiles) {
    if (threadIdx.x == 0) *gen_slot_ptr = ld_acquire_u32(p.sync) + 1u;
    for (int i = threadIdx.x; i < SIDE_RED; i += blockDim.x) side_red[i] = 0.f;
  }
  tc_fence_before();
  __syncthreads();                                         // barriers + TMEM base + generation visible
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const unsigned gen = side_tiles ? *gen_slot_ptr : 0u;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int t = unit; t < num_tiles; t += nunits) {
        const TileInfo ti = decode_tile(t, p.tiles_m, p.tiles_n, side_tiles != 0, p.side_la);
        if (ti.kind == TILE_SKIP) continue;
        const int panel = ti.panel;
        const int m0 = panel * BM;
        if (ti.kind == TILE_SIDE) {
          // ring slot layout of a side tile: [A0 k-block 2f | A0 k-block 2f+1 | P k-block 2f | P k-block 2f+1]
          // (half as many ring round trips as one k-block per slot: the loop is bound by TMA latency, not bytes)
          for (int f = 0; f < side_fills; ++f) {
            mbar_wait_park(empty_bar(s), ph ^ 1u);
            const uint32_t sa = tiles + s * STAGE_BYTES;
            const int nsub = p.kblocks_main - 2 * f < 2 ? 1 : 2;
            mbar_expect_tx(full_bar(s), nsub * (A_BYTES + side_n * BK * 2));
            for (int j = 0; j < nsub; ++j) {
              const int kb = 2 * f + j;
              tma_load_2d(sa + j * A_BYTES, &mapA0, full_bar(s), kb * BK, m0);
              tma_load_2d(sa + 2 * A_BYTES + j * side_n * BK * 2, &mapP, full_bar(s), (kb % p.side_kb_slice) * BK, 0);
            }
            if (++s == STAGES) { s = 0; ph ^= 1u; }
          }
          continue;
        }
        const int n0 = ti.n0;
        if (p.prefetch) {
          // (experiment, CARA_GEMM_PREFETCH=1, off: K = 768 shapes +3 %, K >= 2304 shapes -10 %: the extra L2 requests cost
          // more than the latency they hide)
          // L2 prefetch of the A0 panel of this CTA's NEXT output tile.  A panel is read first by whichever of its
          // tiles_n tiles gets there first -- a miss to HBM that all of them then wait for, and four ring slots do not
          // cover HBM latency.  The tiles_n CTAs that will work on that panel one round from now each pull a disjoint
          // 1/tiles_n of its k-blocks into L2 now (k-block index congruent to their column-tile index), so by then the
          // slot loads hit L2.
          const int t2 = t + nunits;
          if (t2 < num_tiles) {
            const TileInfo nx = decode_tile(t2, p.tiles_m, p.tiles_n, side_tiles != 0, p.side_la);
            if (nx.kind == TILE_MAIN)
              for (int kb = nx.n0 / BN; kb < p.kblocks_main; kb += p.tiles_n) tma_prefetch_2d(&mapA0, kb * BK, nx.panel * BM);
          }
        }
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait_park(empty_bar(s), ph ^ 1u);       // stage free
          const uint32_t sa = tiles + s * STAGE_BYTES;
          const uint32_t sb = sa + A_BYTES;
          mbar_expect_tx(full_bar(s), STAGE_BYTES);
          if (kb < p.kblocks_main) {
            tma_load_2d(sa, &mapA0, full_bar(s), kb * BK, m0);
            tma_load_2d(sb, &mapB0, full_bar(s), kb * BK, n0);
          } else {
            const int e = kb - p.kblocks_main;
            if (e == 0 && side_tiles && !(p.debug & 16)) wait_panel(p.sync + 2 + 4 * panel, gen);   // A1 rows of this panel are complete
            const int slice = n0 / p.ext_slice_w;
            tma_load_2d(sa, &mapA1, full_bar(s), slice * p.ext_rp + e * BK, m0);
            tma_load_2d(sb, &mapB1, full_bar(s), e * BK, n0 - slice * p.ext_slice_w);
          }
          if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN);
      const uint32_t idesc_side = umma_idesc_bf16(BM, side_n > 0 ? side_n : 16);
      int s = 0;
      uint32_t ph = 0;
      AccState acc{0, 0u, 0u};
      uint32_t side_count = 0;                    // side tiles issued so far by this CTA
      bool side_undrained = false;                // the last side tile's columns may still be unread
      for (int t = unit; t < num_tiles; t += nunits) {
        const TileInfo ti = decode_tile(t, p.tiles_m, p.tiles_n, side_tiles != 0, p.side_la);
        if (ti.kind == TILE_SKIP) continue;
        const bool is_side = ti.kind == TILE_SIDE;
        const int as = acc.cur;
        mbar_wait_park(tempty_bar(as), acc.parity(as) ^ 1u);   // the output-tile epilogue has drained this accumulator
        if (side_undrained) {                     // ... and the side warps the columns the last side tile borrowed
          mbar_wait_park(sempty_bar, (side_count - 1u) & 1u);
          side_undrained = false;
        }
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * BN);
        const int nkb = is_side ? side_fills : kblocks;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait_park(full_bar(s), ph);             // TMA bytes have landed
          tc_fence_after();
          const uint32_t sa = tiles + s * STAGE_BYTES;
          const uint64_t da = umma_desc_sw128(sa);
          const uint64_t db = umma_desc_sw128(sa + A_BYTES);
          if (is_side) {
            const int nsub = p.kblocks_main - 2 * kb < 2 ? 1 : 2;
            for (int j = 0; j < nsub; ++j) {
              const int kbj = 2 * kb + j;
              const int slice = kbj / p.side_kb_slice;
              const bool first = kbj - slice * p.side_kb_slice == 0;
              const uint32_t d_side = d_tmem + static_cast<uint32_t>(slice * side_n);
              const uint64_t dsa = umma_desc_sw128(sa + j * A_BYTES);
              const uint64_t dsb = umma_desc_sw128(sa + 2 * A_BYTES + j * side_n * BK * 2);
#pragma unroll
              for (int k = 0; k < BK / UK; ++k)
                umma_bf16(d_side, dsa + 2u * k, dsb + 2u * k, idesc_side, (first && k == 0) ? 0u : 1u);
            }
          } else {
            int ksteps = BK / UK;
            if (kb >= p.kblocks_main) {
              const int left = p.ksteps_ext - 4 * (kb - p.kblocks_main);
              ksteps = left < 4 ? left : 4;
            }
            for (int k = 0; k < ksteps; ++k) {
              // advance 16 bf16 = 32 B inside the 128-B swizzle row: +2 in the (addr>>4) field
              umma_bf16(d_tmem, da + 2u * k, db + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
            }
          }
          // when these MMAs retire: free the smem stage / publish the accumulator
          umma_commit(empty_bar(s));
          if (kb == nkb - 1) umma_commit(is_side ? sfull_bar : tfull_bar(as));
          if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
        if (is_side) {
          ++side_count;
          side_undrained = true;
        } else {
          acc.advance();
        }
      }
    }
  } else if (SW && warp >= 2 + EW) {
    // ------------------------------------------------------------------ side-drain warps (one per TMEM lane group)
    const int lg = warp & 3;
    int cur = 0;                                  // mirrors AccState::cur of the MMA warp: flips on every output tile
    uint32_t k = 0;                               // side tiles seen so far
    for (int t = unit; t < num_tiles; t += nunits) {
      const TileInfo ti = decode_tile(t, p.tiles_m, p.tiles_n, side_tiles != 0, p.side_la);
      if (ti.kind == TILE_SKIP) continue;
      if (ti.kind == TILE_MAIN) { cur ^= 1; continue; }
      mbar_wait_park(sfull_bar, k & 1u);
      ++k;
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(lg * 32) << 16) + static_cast<uint32_t>(cur * BN);
      const int grow_s = ti.panel * BM + lg * 32 + lane;
      if (p.side_rp == 16) side_epilogue<16>(p, t_row, grow_s, lane, side_red, sempty_bar);
      else side_epilogue<32>(p, t_row, grow_s, lane, side_red, sempty_bar);
      // publish this warp's 32 rows: the warp-level barrier orders the lanes' stores before lane 0's release,
      // which is cumulative (PTX memory model) -- only that one thread waits for the stores to become visible
      __syncwarp();
      if (lane == 0) st_release_u32(p.sync + 2 + 4 * ti.panel + lg, gen);
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
    AccState acc{0, 0u, 0u};
    uint32_t q = 0;                               // running step counter of this warp (buffer selector)
    auto stage_row = [&](uint32_t buf, const uint32_t (&w)[32]) { // this thread's 64 bf16 -> its swizzled staging row
#pragma unroll
      for (int j = 0; j < 8; ++j)
        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(buf + row_off + ((static_cast<uint32_t>(j) ^ sw) << 4)),
                     "r"(w[4 * j]), "r"(w[4 * j + 1]), "r"(w[4 * j + 2]), "r"(w[4 * j + 3]) : "memory");
    };
    for (int t = unit; t < num_tiles; t += nunits) {
      const TileInfo ti = decode_tile(t, p.tiles_m, p.tiles_n, side_tiles != 0, p.side_la);
      if (ti.kind != TILE_MAIN) continue;         // (side tiles belong to the side-drain warps)
      const int as = acc.cur;
      const int panel = ti.panel;
      const int m0 = panel * BM;
      const int grow = m0 + lg * 32 + lane;       // global row of this thread
      const int n0 = ti.n0;
      uint4 ux[8];                                // GELU': this thread's 64 saved pre-activations of the current step
      if (EPI == EPI_DGELU) {
        const uint4* src = reinterpret_cast<const uint4*>(p.aux + static_cast<size_t>(grow < p.M ? grow : 0) * p.ldaux +
                                                          n0 + c_first * EC);
#pragma unroll
        for (int j = 0; j < 4; ++j) ld_global_nc_v8(src + 2 * j, ux[2 * j], ux[2 * j + 1]);
      }
      // EPI_DELTA: this row's 64 O values (hi, lo) of the current step, requested one step ahead like ux above
      uint4 oh[EPI == EPI_DELTA ? 8 : 1], ol[EPI == EPI_DELTA ? 8 : 1];
      const int drow = grow < p.M ? grow : 0;
      const int d_b = EPI == EPI_DELTA ? drow / p.seq_n : 0;      // sample / token of this thread's row
      const int d_n = EPI == EPI_DELTA ? drow - d_b * p.seq_n : 0;
      auto load_o = [&](int col) {
        const uint4* sh = reinterpret_cast<const uint4*>(p.aux + static_cast<size_t>(drow) * p.ldaux + col);
        const uint4* sl = reinterpret_cast<const uint4*>(p.aux2 + static_cast<size_t>(drow) * p.ldaux2 + col);
#pragma unroll
        for (int j = 0; j < (EPI == EPI_DELTA ? 4 : 0); ++j) {
          ld_global_nc_v8(sh + 2 * j, oh[2 * j], oh[2 * j + 1]);
          ld_global_nc_v8(sl + 2 * j, ol[2 * j], ol[2 * j + 1]);
        }
      };
      if constexpr (EPI == EPI_DELTA) { if (n0 + c_first * EC < p.N) load_o(n0 + c_first * EC); }
      mbar_wait_park(tfull_bar(as), acc.parity(as));
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
          if (lane == 0) mbar_arrive(tempty_bar(as));
        }
        if (p.debug & 1) continue;
        const int n = n0 + c * EC;
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
          // dX through GELU: multiply by gelu'(u), which fc1's forward epilogue saved (aux)
          const uint32_t* uw = reinterpret_cast<const uint32_t*>(ux);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float2 f = unpack_bf16(uw[j]);
            w0[j] = pack_bf16(__uint_as_float(r[2 * j]) * f.x, __uint_as_float(r[2 * j + 1]) * f.y);
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
        if constexpr (EPI == EPI_DELTA) {
          // delta of (row, head n / 64) from the bf16-rounded outputs -- what the attention backward will read as dO
          const uint32_t* hw = reinterpret_cast<const uint32_t*>(oh);
          const uint32_t* lw = reinterpret_cast<const uint32_t*>(ol);
          float d0 = 0.f, d1 = 0.f;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float2 g = unpack_bf16(w0[j]), a = unpack_bf16(hw[j]), b = unpack_bf16(lw[j]);
            d0 = fmaf(g.x, a.x + b.x, d0);
            d1 = fmaf(g.y, a.y + b.y, d1);
          }
          if (grow < p.M && n < p.N)
            p.delta[(static_cast<size_t>(d_b) * (p.N / EC) + static_cast<size_t>(n / EC)) * p.seq_n + d_n] = d0 + d1;
          if (ci + 1 < NCH && n + EC < p.N) load_o(n + EC);      // in flight during the staging below
        }
        // earlier TMA stores must have finished READING the buffer this step writes
        if (lane == 0) {
          if (EBUFS == 2) tma_store_wait_read<1>(); else tma_store_wait_read<0>();
        }
        __syncwarp();
        uint32_t w1[EPI == EPI_GELU ? 32 : 1];                    // fc1: GELU(u) (second output)
        if constexpr (EPI == EPI_GELU) {
          // fc1: `out` (training only) keeps gelu'(u) for backward, `out2` gets GELU(u) for fc2; both are evaluated
          // on the bf16-rounded pre-activation u.
          if (p.out != nullptr) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float2 f = unpack_bf16(w0[j]);
              float2 g, gp;
              gelu_pair_h2(f.x, f.y, g, gp);
              w0[j] = pack_bf16(gp.x, gp.y);
              w1[j] = pack_bf16(g.x, g.y);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float2 f = unpack_bf16(w0[j]);
              w1[j] = pack_bf16(gelu_fast(f.x), gelu_fast(f.y));
            }
          }
        }
        if ((EPI != EPI_GELU || p.out != nullptr) && !(p.debug & 4)) {
          stage_row(buf, w0);
          fence_proxy_async();                                    // staging writes -> visible to the TMA engine
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&mapOut, buf, n, m0 + lg * 32);
            tma_store_commit();
          }
        }
        if constexpr (EPI == EPI_GELU) {
          // second output through the same (single) staging buffer
          if (lane == 0) tma_store_wait_read<0>();
          __syncwarp();
          stage_row(buf, w1);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&mapAux, buf, n, m0 + lg * 32);
            tma_store_commit();
          }
        }
      }
      acc.advance();
    }
    if (lane == 0) tma_store_wait<0>();                           // all output bytes are in global memory
  }

  __syncwarp();                                            // reconverge the single-lane role warps
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
  if (side_tiles) {
    if (p.side == SIDE_BWD && p.side_dc != nullptr) {      // this CTA's dcs partial sums: one atomic per element
      for (int i = threadIdx.x; i < p.side_slices * p.side_rp; i += blockDim.x) atomicAdd(p.side_dc + i, side_red[i]);
    }
    if (threadIdx.x == 0) {                                // exit ticket: the last CTA to leave advances the generation
      __threadfence();
      if (atomicAdd(p.sync + 1, 1u) == gridDim.x - 1) {
        p.sync[1] = 0u;
        __threadfence();
        st_release_u32(p.sync, gen);
      }
    }
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

struct Maps {
  CUtensorMap a0, b0, a1, b1, out, aux, p;
};

template <int EPI, bool SW>
static cudaError_t launch_epi(const Maps& m, const GemmArgs& args, int grid, cudaStream_t st) {
  static bool attr_done = false;
  constexpr int smem = gemm_smem(EPI);
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(gemm_cp_kernel<EPI, SW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(num_threads(EPI, SW));
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  // side tiles synchronise CTAs of ONE launch through global flags: never overlap such a launch with its neighbours
  cfg.numAttrs = ((pdl_mask() & 1) && args.side == SIDE_NONE) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, gemm_cp_kernel<EPI, SW>, m.a0, m.b0, m.a1, m.b1, m.out, m.aux, m.p, args);
}

static int gcd_int(int a, int b) { while (b) { const int t = a % b; a = b; b = t; } return a; }

int gemm_cp_launch(const GemmDesc& d, cudaStream_t st) {
  const bool side_only = d.side != SIDE_NONE && d.N == 0;
  if (d.M <= 0 || d.K0 <= 0 || (d.K0 % 8) != 0) return -10;
  if (!side_only && (d.N <= 0 || (d.N % 64) != 0)) return -10;
  if (!side_only && d.out == nullptr && d.epi != EPI_GELU) return -16;
  Maps m;
  int rc;
  if ((rc = make_map_bf16(&m.a0, d.A0, d.M, d.K0, d.lda0, BM)) != 0) return rc * 10 - 1;
  if (!side_only) {
    if ((rc = make_map_bf16(&m.b0, d.B0, d.N, d.K0, d.ldb0, BN)) != 0) return rc * 10 - 2;
  } else {
    m.b0 = m.a0;
  }
  GemmArgs args{};
  args.M = d.M; args.N = d.N;
  args.kblocks_main = (d.K0 + BK - 1) / BK;
  args.ksteps_ext = 0;
  args.ext_slice_w = d.N > 0 ? d.N : 1; args.ext_rp = 0;
  if (d.A1 != nullptr && !side_only) {
    if (d.K1 <= 0 || (d.K1 % 16) != 0 || d.ext_slices < 1 || (d.N % d.ext_slices) != 0) return -11;
    const int slice_w = d.N / d.ext_slices;
    if (d.ext_slices > 1 && (slice_w % BN) != 0) return -12;
    if ((rc = make_map_bf16(&m.a1, d.A1, d.M, static_cast<long>(d.K1) * d.ext_slices, d.lda1, BM)) != 0) return rc * 10 - 3;
    if ((rc = make_map_bf16(&m.b1, d.B1, slice_w, d.K1, d.ldb1, BN)) != 0) return rc * 10 - 4;
    args.ksteps_ext = d.K1 / UK;
    args.ext_slice_w = slice_w;
    args.ext_rp = d.K1;
  } else {
    m.a1 = m.a0; m.b1 = m.b0;
  }
  args.bias = d.bias;
  args.out = d.out; args.ldo = d.ldo;
  args.out2 = d.out2; args.ldo2 = d.ldo2;
  args.aux = d.aux; args.ldaux = d.ldaux;
  args.aux2 = d.aux2; args.ldaux2 = d.ldaux2;
  args.delta = d.delta; args.seq_n = d.seq_n;
  {
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("CARA_GEMM_DEBUG"); dbg = e != nullptr ? atoi(e) : 0; }
    args.debug = dbg;
  }
  {
    static int pf = -1;
    if (pf < 0) { const char* e = getenv("CARA_GEMM_PREFETCH"); pf = e != nullptr ? atoi(e) : 0; }
    args.prefetch = pf;
  }
  args.tiles_m = (d.M + BM - 1) / BM;
  args.tiles_n = side_only ? 0 : (d.N + BN - 1) / BN;
  // side tiles
  m.p = m.a0;
  args.side = SIDE_NONE;
  if (d.side != SIDE_NONE) {
    if (d.side != SIDE_FWD && d.side != SIDE_BWD) return -17;
    if (!side_only && d.epi != EPI_NONE) return -20;              // side-drain warps exist for the plain epilogue kind only
    if ((d.side_rp != 16 && d.side_rp != 32) || d.side_slices < 1 || d.side_slices > 4) return -17;
    if (d.P == nullptr || d.side_scales == nullptr || d.side_U == nullptr || d.sync == nullptr) return -17;
    if (d.side == SIDE_BWD && d.side_T == nullptr) return -17;
    const int kslices = d.side == SIDE_BWD ? d.side_slices : 1;
    if ((d.K0 % (BK * kslices)) != 0) return -18;
    if (kslices * 2 * d.side_rp > BN) return -18;                 // one accumulator buffer holds every slice
    if (args.tiles_m > kSyncPanels) return -19;
    if ((reinterpret_cast<uintptr_t>(d.side_U) & 15) != 0 || ((d.side_ldu * 2) & 15) != 0) return -17;
    if (d.side_T != nullptr && (reinterpret_cast<uintptr_t>(d.side_T) & 15) != 0) return -17;
    // with an adapter segment the side output must BE the segment's A operand (that is what the flags protect)
    if (args.ksteps_ext != 0 && (static_cast<const void*>(d.side_U) != static_cast<const void*>(d.A1))) return -17;
    if ((rc = make_map_bf16(&m.p, d.P, 2 * d.side_rp, d.K0 / kslices, d.ldp, 2 * d.side_rp)) != 0) return rc * 10 - 8;
    args.side = d.side;
    args.side_rp = d.side_rp;
    args.side_slices = d.side_slices;
    args.side_kb_slice = args.kblocks_main / kslices;
    args.side_scales = d.side_scales;
    args.side_T = d.side_T;
    args.side_U = d.side_U; args.side_ldu = d.side_ldu;
    args.side_dc = d.side_dc;
    args.sync = d.sync;
  }
  const int tpp = args.tiles_n + (args.side != SIDE_NONE ? 1 : 0);
  if (args.tiles_m * tpp <= 0) return -10;
  static int sm_count = 0;
  if (sm_count == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sm_count <= 0)
      sm_count = 148;
  }
  int grid = d.num_sms > 0 ? d.num_sms : sm_count;
  if (grid > sm_count) grid = sm_count;          // persistent CTAs, one per SM: every CTA of the grid must be resident
  if (grid > args.tiles_m * tpp) grid = args.tiles_m * tpp;
  args.side_la = 0;
  if (args.side != SIDE_NONE && tpp > 1) {
    // round-robin over a grid that shares a factor with tiles-per-panel would pin the (short) side tiles to a few CTAs
    // (tried instead: rotating the position a CTA takes inside each round -- slower than dropping a CTA or two)
    while (grid > 1 && gcd_int(grid, tpp) != 1) --grid;
    // side tiles run one full round of the grid ahead of their consumers (see decode_tile)
    static int la_extra = -1;
    if (la_extra < 0) { const char* e = getenv("CARA_SIDE_LA"); la_extra = e != nullptr ? atoi(e) : 1; }
    args.side_la = (grid + tpp - 1) / tpp + la_extra;
    if (args.side_la > args.tiles_m) args.side_la = args.tiles_m;
    if (args.side_la < 1) args.side_la = 1;
  }
  // epilogue tensors: 32-row x 64-column boxes (one per epilogue warp and step)
  if (d.out != nullptr && !side_only) {
    if ((rc = make_map_bf16(&m.out, d.out, d.M, d.N, d.ldo, 32)) != 0) return rc * 10 - 5;
  } else {
    m.out = m.a0;
  }
  m.aux = m.a0;
  if (!side_only) {
    switch (d.epi) {
      case EPI_NONE: break;
      case EPI_GELU:
        if (d.out2 == nullptr) return -13;
        if ((rc = make_map_bf16(&m.aux, d.out2, d.M, d.N, d.ldo2, 32)) != 0) return rc * 10 - 6;
        break;
      case EPI_DGELU:
        if (d.aux == nullptr) return -14;
        if ((rc = make_map_bf16(&m.aux, d.aux, d.M, d.N, d.ldaux, 32)) != 0) return rc * 10 - 7;
        break;
      case EPI_DELTA:
        if (d.aux == nullptr || d.aux2 == nullptr || d.delta == nullptr || d.seq_n <= 0 || d.M % d.seq_n != 0) return -21;
        if ((reinterpret_cast<uintptr_t>(d.aux) & 31) != 0 || (reinterpret_cast<uintptr_t>(d.aux2) & 31) != 0 ||
            ((d.ldaux * 2) & 31) != 0 || ((d.ldaux2 * 2) & 31) != 0)
          return -21;                                              // 256-bit loads of the O rows
        break;
      default: return -15;
    }
  }
  cudaError_t e;
  switch (side_only ? EPI_NONE : d.epi) {
    case EPI_NONE:
      e = args.side != SIDE_NONE ? launch_epi<EPI_NONE, true>(m, args, grid, st) : launch_epi<EPI_NONE, false>(m, args, grid, st);
      break;
    case EPI_GELU: e = launch_epi<EPI_GELU, false>(m, args, grid, st); break;
    case EPI_DELTA: e = launch_epi<EPI_DELTA, false>(m, args, grid, st); break;
    default: e = launch_epi<EPI_DGELU, false>(m, args, grid, st); break;
  }
  return e == cudaSuccess ? 0 : -1000 - static_cast<int>(e);
}

}  // namespace cara
