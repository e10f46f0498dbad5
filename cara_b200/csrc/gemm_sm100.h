// Internal interface of the tcgen05 fused-projection GEMM (see gemm_sm100.cu).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace cara {

enum GemmEpilogue { EPI_NONE = 0, EPI_GELU = 1, EPI_DGELU = 2, EPI_DELTA = 3 };
enum GemmSide { SIDE_NONE = 0, SIDE_FWD = 1, SIDE_BWD = 2 };

constexpr int kSyncPanels = 4096;                // M <= 524,288 rows
constexpr int kSyncWords = 2 + 4 * kSyncPanels;  // generation, exit ticket, one flag per (128-row panel, TMEM lane group)

// Device-side arguments (passed by value).
struct GemmArgs {
  int M, N;
  int kblocks_main;   // ceil(K0 / 64)
  int ksteps_ext;     // K1 / 16 (0: no adapter segment)
  int ext_slice_w;    // output columns per adapter slice (q|k|v, fc1 quarters)
  int ext_rp;         // A1 columns per slice (rank padded to 16)
  int tiles_m, tiles_n;
  const float* bias;  // [N] fp32 or null
  __nv_bfloat16* out; int ldo;
  __nv_bfloat16* out2; int ldo2;
  const __nv_bfloat16* aux; int ldaux;
  // EPI_DELTA (the dX GEMM of the attention output projection, head dim 64 = one epilogue step): aux / aux2 = the
  // attention output O as a bf16 (hi, lo) pair, delta[b, h, n] = sum_d out[m, h*64+d] (O_hi + O_lo)[m, h*64+d], m = b*seq_n + n
  const __nv_bfloat16* aux2; int ldaux2;
  float* delta; int seq_n;
  int prefetch;       // 1: L2-prefetch the A0 panel of the CTA's next tile (experiment CARA_GEMM_PREFETCH=1, default 0)
  int debug;          // experiments (CARA_GEMM_DEBUG): 1 = epilogue only drains TMEM, 4 = no output staging, 16 = no side-tile flag wait
  // rank-R side tiles (one per 128-row panel, ahead of the panel's output tiles; see gemm_sm100.cu)
  int side;             // GemmSide
  int side_rp;          // padded rank (16 / 32); the side MMA has N = 2 * side_rp columns ([hi ; lo] rows of the factor)
  int side_slices;      // SIDE_FWD: output slices of Uhat; SIDE_BWD: K-slices of A0 (one dU accumulator each)
  int side_kb_slice;    // k-blocks per K-slice (SIDE_FWD: kblocks_main)
  int side_la;          // panels of lookahead between a side tile and its consumers (>= one round of the grid)
  const float* side_scales;            // [side_slices, side_rp] fp32
  float* side_T;                       // [M, side_rp] fp32: SIDE_FWD out (may be null), SIDE_BWD in
  __nv_bfloat16* side_U; long side_ldu;// SIDE_FWD: Uhat [M, slices * 3rp]; SIDE_BWD: dThat [M, 3rp]  ([hi | lo | hi])
  float* side_dc;                      // SIDE_BWD: [side_slices, side_rp] fp32, accumulated
  unsigned* sync;                      // kSyncWords device words, zero-initialised once by the caller
};

// Host-side problem description.
struct GemmDesc {
  int M, N, K0;
  const __nv_bfloat16* A0; long lda0;   // [M, K0]
  const __nv_bfloat16* B0; long ldb0;   // [N, K0]
  // adapter segment (optional): A1 [M, ext_slices*K1], B1 [N/ext_slices, K1]
  int K1; int ext_slices;
  const __nv_bfloat16* A1; long lda1;
  const __nv_bfloat16* B1; long ldb1;
  const float* bias;
  __nv_bfloat16* out; int ldo;
  __nv_bfloat16* out2; int ldo2;
  const __nv_bfloat16* aux; int ldaux;
  const __nv_bfloat16* aux2; int ldaux2;   // EPI_DELTA
  float* delta; int seq_n;                // EPI_DELTA
  int epi;
  int num_sms;
  // side tiles (optional): P = [2rp, K0 / (SIDE_BWD ? side_slices : 1)] bf16, rows [hi ; lo] of the transposed factor.
  // With side != 0 and N == 0 only the side tiles run (the stand-alone rank-R row contraction).
  int side; int side_rp; int side_slices;
  const __nv_bfloat16* P; long ldp;
  const float* side_scales;
  float* side_T;
  __nv_bfloat16* side_U; long side_ldu;
  float* side_dc;
  unsigned* sync;
};

int gemm_cp_launch(const GemmDesc& d, cudaStream_t st);

// TMA descriptor of a row-major bf16 [rows, cols] matrix (row pitch ld elements): box = [box_rows, 64 columns],
// 128-byte swizzle.  0 on success.
int make_map_bf16(CUtensorMap* map, const void* base, long rows, long cols, long ld, int box_rows);

}  // namespace cara
