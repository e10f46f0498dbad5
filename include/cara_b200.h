/* cara_b200 -- C ABI of the B200-native CaRA hot path.
 *
 * The reference (BonnBytes/CaRA) is pure Python: its hot path is the PyTorch code of
 * src/cara/cara.py (cp_attn :15-60, cp_mlp :63-95) and the timm ViT it patches; it has no FFI of its
 * own.  These entry points are what a binding for that path would call instead of the ATen ops listed
 * in SURVEY.md section 2.1: plain pointers and sizes, borrowed device memory (the caller -- PyTorch --
 * owns every buffer; nothing is allocated here), an explicit CUDA stream, int status (0 = ok, <0 =
 * error, text via cara_last_error()).  INTEGRATION.md shows the ctypes binding.
 *
 * All matrices are row-major.  "bf16" = __nv_bfloat16 storage.  Strides (ld*) are in elements.
 */
#ifndef CARA_B200_H_
#define CARA_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

#define CARA_B200_ABI_VERSION 1
#if defined(__GNUC__)
#define CARA_API __attribute__((visibility("default")))
#else
#define CARA_API
#endif

enum cara_epilogue { CARA_EPI_NONE = 0, CARA_EPI_GELU = 1, CARA_EPI_DGELU = 2 };

CARA_API int cara_abi_version(void);
CARA_API const char* cara_last_error(void);
/* Select the CUDA device used by subsequent calls from this thread (one process per GPU). */
CARA_API int cara_set_device(int device);

/* Fused CP-adapted projection (replaces cara.py:25-42 qkv, :50-58 proj, :75-82 fc1, :87-93 fc2 and
 * their autograd dX):
 *   out[M,N] = A0[M,K0] * B0[N,K0]^T + bias[N]
 *            + A1[M, slice*K1 : (slice+1)*K1] * B1[N mod (N/ext_slices), K1]^T     (if A1 != NULL)
 * bf16 operands, fp32 accumulation in tensor memory, bf16 outputs.
 *   epi = CARA_EPI_GELU : out (may be NULL) = pre-activation, out2 = GELU(pre-activation)
 *   epi = CARA_EPI_DGELU: out = (.) * gelu'(aux[M,N])   (dX through the fc1 activation)
 * Requirements: K0 % 8 == 0, N % 32 == 0, K1 % 16 == 0, 16-byte aligned bases and row pitches. */
typedef struct cara_gemm_desc {
  int M, N, K0;
  const void* A0; long lda0;
  const void* B0; long ldb0;
  int K1, ext_slices;
  const void* A1; long lda1;
  const void* B1; long ldb1;
  const float* bias;
  void* out;  int ldo;
  void* out2; int ldo2;
  const void* aux; int ldaux;
  int epi;
  int num_sms; /* 0 = all */
} cara_gemm_desc;
CARA_API int cara_gemm_cp(const cara_gemm_desc* d, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CARA_B200_H_ */
