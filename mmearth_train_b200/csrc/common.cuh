// Shared device helpers for the MP-MAE step kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mpmae {

constexpr int kWarp = 32;

__host__ __device__ __forceinline__ int cdiv(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ __forceinline__ int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// In-patch pixel order is a Z-order curve with the ROW bit as the least significant bit of each
// pair, so that the four children (dy, dx) of a stride-2 cell sit at 4*m + (dy + 2*dx): exactly
// MinkowskiEngine's kernel index for a 2x2 kernel (axis 0 fastest, kernel_region.hpp:199-221).
__host__ __device__ __forceinline__ int morton_encode(int py, int px) {
  int m = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) m |= (((py >> i) & 1) << (2 * i)) | (((px >> i) & 1) << (2 * i + 1));
  return m;
}
__host__ __device__ __forceinline__ void morton_decode(int m, int &py, int &px) {
  py = 0; px = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) { py |= ((m >> (2 * i)) & 1) << i; px |= ((m >> (2 * i + 1)) & 1) << i; }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// exact-erf GELU (torch.nn.GELU default, MinkowskiNonlinearity.py:113-114)
__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_grad_f(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
  const float pdf = 0.39894228040143268f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

// h = gelu(x) and d gelu / dx from one erf evaluation
__device__ __forceinline__ void gelu_both_f(float x, float &h, float &dgelu) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
  const float pdf = 0.39894228040143268f * __expf(-0.5f * x * x);
  h = x * cdf;
  dgelu = cdf + x * pdf;
}

// Geometry of the visible-patch row layout shared by all sparse kernels.
//   rows of stage with patch side P:  row = (n*V + slot)*P*P + morton(py, px)
//   slot_of[n*L + l] = slot of patch l (ascending patch index among visible ones) or -1
struct Geo {
  int B;  // samples
  int G;  // patch grid side (7)
  int L;  // G*G
  int V;  // visible patches per sample
};

}  // namespace mpmae
