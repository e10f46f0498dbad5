// GPU input pipeline for the entry point (reference image_classification/vtab.py:79-82):
//   transforms.Resize((224, 224), interpolation=3) -> ToTensor() -> Normalize(mean, std)
// on decoded uint8 HWC images.  Resize(interpolation=3) on a PIL image is Pillow's two-pass antialiased bicubic
// (libImaging/Resample.c, Pillow 12.2 -- an un-vendored dependency, restated): a horizontal pass then a vertical pass,
// each a fixed-point (22 fractional bits) weighted sum over the taps of every output pixel, rounded and clipped to
// uint8 BETWEEN the passes.  The integer tap tables (bounds, coefficients) are built on the host exactly as Pillow
// builds them (cara_b200/preprocess.py); the kernels below do the integer sums, so the resized image is bit-identical
// to Pillow's, and ToTensor / Normalize are the same three IEEE fp32 operations torchvision performs
// (u / 255, - mean, / std; explicit round-to-nearest intrinsics because the library is built with --use_fast_math).
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"

namespace cara {
namespace {

constexpr int PRECISION_BITS = 32 - 8 - 2;

__device__ __forceinline__ int clip8(int v) {
  v >>= PRECISION_BITS;                       // arithmetic shift, as Pillow's clip8 lookup index
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// src [B, H, W, 3] -> dst [B, H, OW, 3]; one block per (image, input row), threads over the output columns
__global__ void __launch_bounds__(256)
resize_h_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, const int* __restrict__ bounds,
                const int* __restrict__ kk, int ksize, int W, int OW) {
  const long row = blockIdx.x;                // b * H + y
  for (int xx = threadIdx.x; xx < OW; xx += blockDim.x) {
    const int xmin = __ldg(bounds + 2 * xx), xmax = __ldg(bounds + 2 * xx + 1);
    const int* k = kk + xx * ksize;
    const uint8_t* p = src + (row * W + xmin) * 3;
    int s0 = 1 << (PRECISION_BITS - 1), s1 = s0, s2 = s0;
    for (int x = 0; x < xmax; ++x) {
      const int w = __ldg(k + x);
      s0 += p[3 * x + 0] * w;
      s1 += p[3 * x + 1] * w;
      s2 += p[3 * x + 2] * w;
    }
    uint8_t* o = dst + (row * OW + xx) * 3;
    o[0] = static_cast<uint8_t>(clip8(s0));
    o[1] = static_cast<uint8_t>(clip8(s1));
    o[2] = static_cast<uint8_t>(clip8(s2));
  }
}

// src [B, H, OW, 3] uint8 -> out [B, 3, OH, OW] fp32 normalised (vertical pass when bounds != nullptr);
// one block per (image, output row): the taps of the row are uniform across the block
__global__ void __launch_bounds__(256)
resize_v_norm_kernel(const uint8_t* __restrict__ src, float* __restrict__ out, uint8_t* __restrict__ out_u8,
                     const int* __restrict__ bounds, const int* __restrict__ kk, int ksize, int H, int OH, int OW,
                     float m0, float m1, float m2, float d0, float d1, float d2) {
  const int yy = blockIdx.x % OH;
  const long b = blockIdx.x / OH;
  const int ymin = bounds != nullptr ? __ldg(bounds + 2 * yy) : yy;
  const int ymax = bounds != nullptr ? __ldg(bounds + 2 * yy + 1) : 1;
  const int* k = kk + yy * ksize;
  const long plane = static_cast<long>(OH) * OW;
  for (int xx = threadIdx.x; xx < OW; xx += blockDim.x) {
    const uint8_t* p = src + ((b * H + ymin) * OW + xx) * 3;
    int v0, v1, v2;
    if (bounds != nullptr) {
      int s0 = 1 << (PRECISION_BITS - 1), s1 = s0, s2 = s0;
      for (int y = 0; y < ymax; ++y) {
        const int w = __ldg(k + y);
        const uint8_t* q = p + static_cast<long>(y) * OW * 3;
        s0 += q[0] * w;
        s1 += q[1] * w;
        s2 += q[2] * w;
      }
      v0 = clip8(s0); v1 = clip8(s1); v2 = clip8(s2);
    } else {
      v0 = p[0]; v1 = p[1]; v2 = p[2];
    }
    const long pix = static_cast<long>(blockIdx.x) * OW + xx;
    if (out_u8 != nullptr) {                  // the resized uint8 image (parity checks against Pillow)
      uint8_t* o = out_u8 + pix * 3;
      o[0] = static_cast<uint8_t>(v0); o[1] = static_cast<uint8_t>(v1); o[2] = static_cast<uint8_t>(v2);
    }
    if (out != nullptr) {
      float* o = out + b * 3 * plane + static_cast<long>(yy) * OW + xx;
      o[0] = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(v0), 255.0f), m0), d0);
      o[plane] = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(v1), 255.0f), m1), d1);
      o[2 * plane] = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(v2), 255.0f), m2), d2);
    }
  }
}

}  // namespace

int resize_norm_launch(const uint8_t* src, int B, int H, int W, const int* xbounds, const int* xk, int xksize,
                       const int* ybounds, const int* yk, int yksize, uint8_t* tmp, float* out, uint8_t* out_u8,
                       int OH, int OW, const float* mean, const float* stdv, cudaStream_t st) {
  if (B <= 0 || H <= 0 || W <= 0 || OH <= 0 || OW <= 0) return -80;
  if ((xbounds == nullptr) != (W == OW) || (ybounds == nullptr) != (H == OH)) return -80;   // a pass is skipped only at equal size
  if (xbounds != nullptr && tmp == nullptr) return -80;
  if (static_cast<long>(B) * H > 2147483647L || static_cast<long>(B) * OH > 2147483647L) return -80;
  const uint8_t* cur = src;
  if (xbounds != nullptr) {
    resize_h_kernel<<<B * H, 256, 0, st>>>(src, tmp, xbounds, xk, xksize, W, OW);
    cur = tmp;
  }
  resize_v_norm_kernel<<<B * OH, 256, 0, st>>>(cur, out, out_u8, ybounds, yk, yksize, H, OH, OW, mean[0], mean[1], mean[2],
                                               stdv[0], stdv[1], stdv[2]);
  return cudaGetLastError() == cudaSuccess ? 0 : -81;
}

}  // namespace cara
