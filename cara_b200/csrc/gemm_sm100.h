// Internal interface of the tcgen05 fused-projection GEMM (see gemm_sm100.cu).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace cara {

enum GemmEpilogue { EPI_NONE = 0, EPI_GELU = 1, EPI_DGELU = 2 };

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
  int debug;          // experiments (CARA_GEMM_DEBUG): 1 = epilogue only drains TMEM, 2 = no TMA loads / no full-barrier waits
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
  int epi;
  int num_sms;
  int pair;   // 1: CTA pairs (cta_group::2, 256 x 256 tiles); 0: single-CTA 128 x 256 tiles
};

int gemm_cp_launch(const GemmDesc& d, cudaStream_t st);

// TMA descriptor of a row-major bf16 [rows, cols] matrix (row pitch ld elements): box = [box_rows, 64 columns],
// 128-byte swizzle.  0 on success.
int make_map_bf16(CUtensorMap* map, const void* base, long rows, long cols, long ld, int box_rows);

}  // namespace cara
