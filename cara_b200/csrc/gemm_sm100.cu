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
// Roles (192 threads): warp 0 = TMA producer (one elected lane), warp 1 = tcgen05.mma issuer + TMEM
// owner, warps 2..5 = epilogue.  Two 256-column fp32 accumulators in TMEM let the epilogue of tile i
// overlap the main loop of tile i+1.  Epilogue data path: tcgen05.ld -> registers (bias / GELU / GELU')
// -> 128-byte-swizzled shared staging (conflict-free 16-byte stores) -> TMA tensor store, 32 rows x 64
// columns per warp and step, double buffered; the GELU' operand arrives the same way through TMA loads
// prefetched one step ahead.  HBM therefore only ever sees full 128-byte lines.
#include "ptx.cuh"
#include "gemm_sm100.h"
#include "mathfn.cuh"

namespace cara {

constexpr int BM = 128, BN = 256, BK = 64;      // CTA tile; BK*2B = one 128-byte swizzle row
constexpr int UK = 16;                          // tcgen05 kind::f16 K per instruction
constexpr int A_BYTES = BM * BK * 2;            // 16 KB
constexpr int B_BYTES = BN * BK * 2;            // 32 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;  // 48 KB
constexpr int TMEM_COLS = 512;                  // 2 accumulators x 256 fp32 columns
constexpr int NUM_THREADS = 192;
constexpr int EC = 64;                          // epilogue step: 64 output columns = one 128-B swizzle row
constexpr int EBUF = 32 * EC * 2;               // 4 KB staging buffer: 32 rows x 128 B
// main-loop ring depth / staging tensors per epilogue kind (smem budget 227 KB)
__host__ __device__ constexpr int num_stages(int epi) { return epi == EPI_NONE ? 4 : 3; }
__host__ __device__ constexpr int num_stage_tensors(int epi) { return epi == EPI_NONE ? 1 : 2; }
__host__ __device__ constexpr int gemm_smem(int epi) {
  return num_stages(epi) * STAGE_BYTES + 4 * 2 * EBUF * num_stage_tensors(epi) + 1024 /*align slack*/ + 256 /*barriers*/;
}

struct TileCoord {
  int m0, n0;
};
__device__ __forceinline__ TileCoord tile_coord(int t, int tiles_n) {
  // n fastest: the CTAs resident at one time share a handful of A row-panels and all of B through L2
  return {(t / tiles_n) * BM, (t % tiles_n) * BN};
}

template <int EPI>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_cp_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapB0,
               const __grid_constant__ CUtensorMap mapA1, const __grid_constant__ CUtensorMap mapB1,
               const __grid_constant__ CUtensorMap mapOut, const __grid_constant__ CUtensorMap mapAux,
               const GemmArgs p) {
  constexpr int STAGES = num_stages(EPI);
  constexpr int NST = num_stage_tensors(EPI);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t tiles = (raw + 1023u) & ~1023u;           // SWIZZLE_128B atoms need 1024-B alignment
  const uint32_t stage_out = tiles + STAGES * STAGE_BYTES; // [4 warps][NST tensors][2 buffers][4 KB]
  const uint32_t bars = stage_out + 4 * 2 * EBUF * NST;
  // barrier map (8 B each): full[S], empty[S], tfull[2], tempty[2], aux[4 warps][2], then the TMEM base word
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bars + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (2 * STAGES + 2 + a); };
  auto aux_bar = [&](int w, int b) { return bars + 8u * (2 * STAGES + 4 + w * 2 + b); };
  const uint32_t tmem_slot = bars + 8u * (2 * STAGES + 12);
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
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 4);  // one arrive per epilogue warp
    }
    for (int w = 0; w < 4; ++w) {
      mbar_init(aux_bar(w, 0), 1);
      mbar_init(aux_bar(w, 1), 1);
    }
    tma_prefetch_desc(&mapOut);
    if (EPI != EPI_NONE) tma_prefetch_desc(&mapAux);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const TileCoord tc = tile_coord(t, p.tiles_n);
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(empty_bar(s), ph ^ 1u);
          const uint32_t sa = tiles + s * STAGE_BYTES;
          const uint32_t sb = sa + A_BYTES;
          mbar_expect_tx(full_bar(s), STAGE_BYTES);
          if (kb < p.kblocks_main) {
            tma_load_2d(sa, &mapA0, full_bar(s), kb * BK, tc.m0);
            tma_load_2d(sb, &mapB0, full_bar(s), kb * BK, tc.n0);
          } else {
            const int e = kb - p.kblocks_main;
            const int slice = tc.n0 / p.ext_slice_w;
            tma_load_2d(sa, &mapA1, full_bar(s), slice * p.ext_rp + e * BK, tc.m0);
            tma_load_2d(sb, &mapB1, full_bar(s), e * BK, tc.n0 - slice * p.ext_slice_w);
          }
          if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN);
      int s = 0;
      uint32_t ph = 0;
      int as = 0;
      uint32_t aph = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        mbar_wait(tempty_bar(as), aph ^ 1u);      // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * BN);
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(full_bar(s), ph);             // TMA bytes have landed
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
            umma_bf16(d_tmem, da + 2u * k, db + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(empty_bar(s));              // frees the smem stage when these MMAs retire
          if (kb == kblocks - 1) umma_commit(tfull_bar(as));
          if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
        if (++as == 2) { as = 0; aph ^= 1u; }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps 2..5
    const int lg = warp & 3;                      // TMEM lane group this warp may touch
    const int ew = warp - 2;                      // staging slot of this warp
    const uint32_t my_stage = stage_out + ew * (2 * EBUF * NST);   // tensor 0: [2][EBUF]; tensor 1: [2][EBUF]
    const uint32_t sw = static_cast<uint32_t>(lane & 7);          // 128-B swizzle phase of this thread's row
    const uint32_t row_off = static_cast<uint32_t>(lane) * 128u;
    int as = 0;
    uint32_t aph = 0;
    uint32_t q = 0;                               // running 64-column step counter (buffer / parity selector)
    if (EPI == EPI_DGELU && lane == 0 && static_cast<int>(blockIdx.x) < num_tiles) {
      const TileCoord t0 = tile_coord(blockIdx.x, p.tiles_n);
      mbar_expect_tx(aux_bar(ew, 0), EBUF);
      tma_load_2d(my_stage + 2 * EBUF, &mapAux, aux_bar(ew, 0), t0.n0, t0.m0 + lg * 32);
    }
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const TileCoord tc = tile_coord(t, p.tiles_n);
      mbar_wait(tfull_bar(as), aph);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(lg * 32) << 16) + static_cast<uint32_t>(as * BN);
#pragma unroll 1
      for (int c = 0; c < BN / EC; ++c, ++q) {
        const uint32_t b = q & 1u;
        const uint32_t so = my_stage + b * EBUF;                  // staging of the primary output
        const uint32_t sx = my_stage + 2 * EBUF + b * EBUF;       // second tensor: GELU output / GELU' operand
        // the TMA store that read this buffer two steps ago must have drained it
        if (lane == 0) tma_store_wait_read<1>();
        __syncwarp();
        if (EPI == EPI_DGELU) {
          if (lane == 0) {                                        // prefetch the next step's operand tile
            int nt = t, nc = c + 1;
            if (nc == BN / EC) { nt = t + gridDim.x; nc = 0; }
            if (nt < num_tiles) {
              const TileCoord tn = tile_coord(nt, p.tiles_n);
              mbar_expect_tx(aux_bar(ew, b ^ 1u), EBUF);
              tma_load_2d(my_stage + 2 * EBUF + (b ^ 1u) * EBUF, &mapAux, aux_bar(ew, b ^ 1u), tn.n0 + nc * EC,
                          tn.m0 + lg * 32);
            }
          }
          mbar_wait(aux_bar(ew, b), (q >> 1) & 1u);
        }
        uint32_t r[64];
        tmem_ld32(t_row + c * EC, *reinterpret_cast<uint32_t(*)[32]>(&r[0]));
        tmem_ld32(t_row + c * EC + 32, *reinterpret_cast<uint32_t(*)[32]>(&r[32]));
        tmem_ld_wait();
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
#pragma unroll
        for (int j = 0; j < 8; ++j) {                             // 8 x 16-byte chunks of this thread's row
          const uint32_t off = row_off + ((static_cast<uint32_t>(j) ^ sw) << 4);
          float v[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(r[8 * j + e]);
          if (EPI == EPI_DGELU) {
            // dX through GELU: multiply by gelu'(u), u = saved fc1 pre-activation (TMA-staged, same swizzle)
            uint32_t u0, u1, u2, u3;
            asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(u0), "=r"(u1), "=r"(u2), "=r"(u3) : "r"(sx + off));
            const uint32_t uw[4] = {u0, u1, u2, u3};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 f = unpack_bf16(uw[e]);
              v[2 * e + 0] *= gelu_grad_fast(f.x);
              v[2 * e + 1] *= gelu_grad_fast(f.y);
            }
          }
          const uint32_t o0 = pack_bf16(v[0], v[1]), o1 = pack_bf16(v[2], v[3]);
          const uint32_t o2 = pack_bf16(v[4], v[5]), o3 = pack_bf16(v[6], v[7]);
          if (p.out != nullptr)
            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(so + off), "r"(o0), "r"(o1), "r"(o2), "r"(o3) : "memory");
          if (EPI == EPI_GELU) {
            // fc1: `out` keeps the pre-activation for backward, `out2` gets GELU(u) for fc2.  GELU is applied
            // to the bf16-rounded pre-activation so forward and backward see the same u.
            const uint32_t ow[4] = {o0, o1, o2, o3};
            uint32_t gw[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 f = unpack_bf16(ow[e]);
              gw[e] = pack_bf16(gelu_fast(f.x), gelu_fast(f.y));
            }
            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(sx + off), "r"(gw[0]), "r"(gw[1]), "r"(gw[2]), "r"(gw[3]) : "memory");
          }
        }
        if (c == BN / EC - 1) {                                   // accumulator fully read: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar(as));
        }
        fence_proxy_async();                                      // staging writes -> visible to the TMA engine
        __syncwarp();
        if (lane == 0) {
          if (p.out != nullptr) tma_store_2d(&mapOut, so, n, tc.m0 + lg * 32);
          if (EPI == EPI_GELU) tma_store_2d(&mapAux, sx, n, tc.m0 + lg * 32);
          tma_store_commit();
        }
      }
      if (++as == 2) { as = 0; aph ^= 1u; }
    }
    if (lane == 0) tma_store_wait<0>();                           // all output bytes are in global memory
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
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
static int make_map_bf16(CUtensorMap* map, const void* base, long rows, long cols, long ld, int box_rows) {
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

template <int EPI>
static cudaError_t launch_epi(const CUtensorMap& a0, const CUtensorMap& b0, const CUtensorMap& a1,
                              const CUtensorMap& b1, const CUtensorMap& mo, const CUtensorMap& mx,
                              const GemmArgs& args, int grid, cudaStream_t st) {
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(gemm_cp_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, gemm_smem(EPI));
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  gemm_cp_kernel<EPI><<<grid, NUM_THREADS, gemm_smem(EPI), st>>>(a0, b0, a1, b1, mo, mx, args);
  return cudaGetLastError();
}

int gemm_cp_launch(const GemmDesc& d, cudaStream_t st) {
  if (d.M <= 0 || d.N <= 0 || d.K0 <= 0 || (d.K0 % 8) != 0 || (d.N % 64) != 0) return -10;
  if (d.out == nullptr && d.epi != EPI_GELU) return -16;
  CUtensorMap a0, b0, a1, b1, mo, mx;
  int rc;
  if ((rc = make_map_bf16(&a0, d.A0, d.M, d.K0, d.lda0, BM)) != 0) return rc * 10 - 1;
  if ((rc = make_map_bf16(&b0, d.B0, d.N, d.K0, d.ldb0, BN)) != 0) return rc * 10 - 2;
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
    if ((rc = make_map_bf16(&b1, d.B1, slice_w, d.K1, d.ldb1, BN)) != 0) return rc * 10 - 4;
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
  args.tiles_m = (d.M + BM - 1) / BM;
  args.tiles_n = (d.N + BN - 1) / BN;
  const int num_tiles = args.tiles_m * args.tiles_n;
  int grid = d.num_sms > 0 ? d.num_sms : 148;
  if (grid > num_tiles) grid = num_tiles;
  // epilogue tensors: 32-row x 64-column boxes (one per epilogue warp and step)
  if (d.out != nullptr) {
    if ((rc = make_map_bf16(&mo, d.out, d.M, d.N, d.ldo, 32)) != 0) return rc * 10 - 5;
  } else {
    mo = a0;
  }
  mx = a0;
  cudaError_t e;
  switch (d.epi) {
    case EPI_NONE: e = launch_epi<EPI_NONE>(a0, b0, a1, b1, mo, mx, args, grid, st); break;
    case EPI_GELU:
      if (d.out2 == nullptr) return -13;
      if ((rc = make_map_bf16(&mx, d.out2, d.M, d.N, d.ldo2, 32)) != 0) return rc * 10 - 6;
      e = launch_epi<EPI_GELU>(a0, b0, a1, b1, mo, mx, args, grid, st); break;
    case EPI_DGELU:
      if (d.aux == nullptr) return -14;
      if ((rc = make_map_bf16(&mx, d.aux, d.M, d.N, d.ldaux, 32)) != 0) return rc * 10 - 7;
      e = launch_epi<EPI_DGELU>(a0, b0, a1, b1, mo, mx, args, grid, st); break;
    default: return -15;
  }
  return e == cudaSuccess ? 0 : -1000 - static_cast<int>(e);
}

}  // namespace cara
