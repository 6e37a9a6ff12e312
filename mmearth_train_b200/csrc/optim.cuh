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
                             float b2, float eps, float wd, float bc1, float bc2_sqrt, float ginv) { pdl_prologue();
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

// 4 elements per thread (16-byte loads / stores); n % 4 == 0 and 16-byte aligned buffers
__global__ void adamw_vec4_kernel(float4 *__restrict__ p, const float4 *__restrict__ g, float4 *__restrict__ m,
                                  float4 *__restrict__ v, const uchar4 *__restrict__ decay, int64_t n4, float lr, float b1,
                                  float b2, float eps, float wd, float bc1, float bc2_sqrt, float ginv) { pdl_prologue();
  const float keep = 1.f - lr * wd, step = lr / bc1;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 g4 = g[i];
    float4 p4 = p[i], m4 = m[i], v4 = v[i];
    const uchar4 d4 = decay ? decay[i] : make_uchar4(1, 1, 1, 1);
    auto upd = [&](float &pp, float &mm, float &vv, float gg, unsigned char dd) {
      gg *= ginv;
      if (dd) pp *= keep;
      mm = b1 * mm + (1.f - b1) * gg;
      vv = b2 * vv + (1.f - b2) * gg * gg;
      pp -= step * (mm / (sqrtf(vv) / bc2_sqrt + eps));
    };
    upd(p4.x, m4.x, v4.x, g4.x, d4.x); upd(p4.y, m4.y, v4.y, g4.y, d4.y);
    upd(p4.z, m4.z, v4.z, g4.z, d4.z); upd(p4.w, m4.w, v4.w, g4.w, d4.w);
    p[i] = p4; m[i] = m4; v[i] = v4;
  }
}

// The same update with the per-step scalars read from DEVICE memory, so a GradScaler-style loop (helpers.py:470-506) needs no
// host sync: state[0] = factor applied to the gradient (1 / loss scale, times the clipping coefficient), state[1] != 0 =
// non-finite gradient found -> the launch leaves parameters and moments untouched (GradScaler.step skips the step),
// state[2] = 1-based number of this step (bias corrections); a NEGATIVE host lr means "read the learning rate from state[3]"
// (a step captured in a CUDA graph cannot take per-iteration host scalars).
__global__ void adamw_vec4_dev_kernel(float4 *__restrict__ p, const float4 *__restrict__ g, float4 *__restrict__ m,
                                      float4 *__restrict__ v, const uchar4 *__restrict__ decay, int64_t n4, float lr_host, float b1,
                                      float b2, float eps, float wd, const float *__restrict__ state) { pdl_prologue();
  const float ginv = state[0];
  if (state[1] != 0.f) return;
  const float lr = lr_host < 0.f ? state[3] : lr_host;
  const double t = (double)state[2];
  const float bc1 = 1.f - (float)pow((double)b1, t);
  const float bc2_sqrt = sqrtf(1.f - (float)pow((double)b2, t));
  const float keep = 1.f - lr * wd, step = lr / bc1;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 g4 = g[i];
    float4 p4 = p[i], m4 = m[i], v4 = v[i];
    const uchar4 d4 = decay ? decay[i] : make_uchar4(1, 1, 1, 1);
    auto upd = [&](float &pp, float &mm, float &vv, float gg, unsigned char dd) {
      gg *= ginv;
      if (dd) pp *= keep;
      mm = b1 * mm + (1.f - b1) * gg;
      vv = b2 * vv + (1.f - b2) * gg * gg;
      pp -= step * (mm / (sqrtf(vv) / bc2_sqrt + eps));
    };
    upd(p4.x, m4.x, v4.x, g4.x, d4.x); upd(p4.y, m4.y, v4.y, g4.y, d4.y);
    upd(p4.z, m4.z, v4.z, g4.z, d4.z); upd(p4.w, m4.w, v4.w, g4.w, d4.w);
    p[i] = p4; m[i] = m4; v[i] = v4;
  }
}

inline cudaError_t launch_adamw_dev(float *p, const float *g, float *m, float *v, const uint8_t *decay, int64_t n, float lr,
                                    float b1, float b2, float eps, float wd, const float *state, cudaStream_t st) {
  const uintptr_t al = (uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v;
  if (n % 4 != 0 || (al & 15) != 0 || ((uintptr_t)decay & 3) != 0) return cudaErrorInvalidValue;
  int64_t g4 = cdiv64(n / 4, 256);
  if (g4 > 148 * 8) g4 = 148 * 8;
  pdl(adamw_vec4_dev_kernel, (unsigned)g4, 256, 0, st)(reinterpret_cast<float4 *>(p), reinterpret_cast<const float4 *>(g),
                                                     reinterpret_cast<float4 *>(m), reinterpret_cast<float4 *>(v),
                                                     reinterpret_cast<const uchar4 *>(decay), n / 4, lr, b1, b2, eps, wd, state);
  return cudaGetLastError();
}

inline cudaError_t launch_adamw(float *p, const float *g, float *m, float *v, const uint8_t *decay, int64_t n, float lr,
                                float b1, float b2, float eps, float wd, int64_t step, float ginv, cudaStream_t st) {
  const float bc1 = 1.f - (float)pow((double)b1, (double)step);
  const float bc2 = 1.f - (float)pow((double)b2, (double)step);
  int64_t grid = cdiv64(n, 256);
  if (grid > 148 * 8) grid = 148 * 8;
  const uintptr_t al = (uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v;
  if (n % 4 == 0 && (al & 15) == 0 && ((uintptr_t)decay & 3) == 0) {
    int64_t g4 = cdiv64(n / 4, 256);
    if (g4 > 148 * 8) g4 = 148 * 8;
    pdl(adamw_vec4_kernel, (unsigned)g4, 256, 0, st)(reinterpret_cast<float4 *>(p), reinterpret_cast<const float4 *>(g),
                                                   reinterpret_cast<float4 *>(m), reinterpret_cast<float4 *>(v),
                                                   reinterpret_cast<const uchar4 *>(decay), n / 4, lr, b1, b2, eps, wd, bc1,
                                                   sqrtf(bc2), ginv);
    return cudaGetLastError();
  }
  pdl(adamw_kernel, (unsigned)grid, 256, 0, st)(p, g, m, v, decay, n, lr, b1, b2, eps, wd, bc1, sqrtf(bc2), ginv);
  return cudaGetLastError();
}

}  // namespace mpmae
