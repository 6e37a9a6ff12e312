// Fused AdamW over the flat parameter / gradient buffers.
// Same update as torch.optim.AdamW (decoupled weight decay, bias-corrected moments), which the
// reference builds at main_pretrain.py:312-320 with timm's no-decay rule expressed here as a
// per-element byte mask; the reference's ~184 small tensors x several kernels become one launch.
#pragma once
#include <math.h>

#include "common.cuh"

namespace mpmae {

__global__ void adamw_kernel(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m,
                             float *__restrict__ v, const uint8_t *__restrict__ decay, int64_t n, float lr, float b1,
                             float b2, float eps, float wd, float bc1, float bc2_sqrt, float ginv) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = g[i] * ginv;
    float pi = p[i];
    if (!decay || decay[i]) pi *= 1.f - lr * wd;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = pi - (lr / bc1) * (mi / denom);
  }
}

inline cudaError_t launch_adamw(float *p, const float *g, float *m, float *v, const uint8_t *decay, int64_t n, float lr,
                                float b1, float b2, float eps, float wd, int64_t step, float ginv, cudaStream_t st) {
  const float bc1 = 1.f - (float)pow((double)b1, (double)step);
  const float bc2 = 1.f - (float)pow((double)b2, (double)step);
  int64_t grid = cdiv64(n, 256);
  if (grid > 148 * 8) grid = 148 * 8;
  adamw_kernel<<<(unsigned)grid, 256, 0, st>>>(p, g, m, v, decay, n, lr, b1, b2, eps, wd, bc1, sqrtf(bc2), ginv);
  return cudaGetLastError();
}

}  // namespace mpmae
