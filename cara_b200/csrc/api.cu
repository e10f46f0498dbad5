// C-ABI entry points (include/cara_b200.h).  Thin: validate, translate, launch.
#include "../../include/cara_b200.h"
#include "gemm_sm100.h"

#include <cstdio>
#include <cstring>

namespace {
thread_local char g_err[512] = "";
thread_local int g_device = -1;

int fail(int code, const char* what) {
  std::snprintf(g_err, sizeof(g_err), "%s (code %d, cuda: %s)", what, code,
                cudaGetErrorString(cudaPeekAtLastError()));
  return code;
}
}  // namespace

extern "C" {

int cara_abi_version(void) { return CARA_B200_ABI_VERSION; }
const char* cara_last_error(void) { return g_err; }

int cara_set_device(int device) {
  if (device == g_device) return 0;
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return fail(-1, "cudaSetDevice failed");
  g_device = device;
  return 0;
}

int cara_gemm_cp(const cara_gemm_desc* d, void* stream) {
  if (d == nullptr) return fail(-2, "cara_gemm_cp: null descriptor");
  cara::GemmDesc g{};
  g.M = d->M; g.N = d->N; g.K0 = d->K0;
  g.A0 = static_cast<const __nv_bfloat16*>(d->A0); g.lda0 = d->lda0;
  g.B0 = static_cast<const __nv_bfloat16*>(d->B0); g.ldb0 = d->ldb0;
  g.K1 = d->K1; g.ext_slices = d->ext_slices;
  g.A1 = static_cast<const __nv_bfloat16*>(d->A1); g.lda1 = d->lda1;
  g.B1 = static_cast<const __nv_bfloat16*>(d->B1); g.ldb1 = d->ldb1;
  g.bias = d->bias;
  g.out = static_cast<__nv_bfloat16*>(d->out); g.ldo = d->ldo;
  g.out2 = static_cast<__nv_bfloat16*>(d->out2); g.ldo2 = d->ldo2;
  g.aux = static_cast<const __nv_bfloat16*>(d->aux); g.ldaux = d->ldaux;
  g.epi = d->epi; g.num_sms = d->num_sms;
  int rc = cara::gemm_cp_launch(g, static_cast<cudaStream_t>(stream));
  if (rc != 0) return fail(rc, "cara_gemm_cp: launch failed / bad arguments");
  return 0;
}

}  // extern "C"
