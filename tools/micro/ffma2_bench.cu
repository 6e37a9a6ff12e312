// Microbenchmark: FP32 FMA issue rate on sm_100a, scalar FFMA vs packed FFMA2 (fma.rn.f32x2).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ffma2_bench tools/micro/ffma2_bench.cu && /tmp/ffma2_bench
#include <cuda_runtime.h>
#include <cstdio>

__device__ __forceinline__ unsigned long long pk(float a, float b) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}

template <int ILP>
__global__ void k_scalar(float *out, float w, int iters) {
  float acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 0.001f + i;
  float x = out[threadIdx.x & 31];
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = fmaf(acc[i], x, w);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
__global__ void k_packed(float *out, float w, int iters) {
  unsigned long long acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc[i] = pk(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f + i);
  const float xv = out[threadIdx.x & 31];
  const unsigned long long x = pk(xv, xv + 1e-3f), ww = pk(w, w);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = fma2(acc[i], x, ww);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < ILP; ++i) {
    float a, b;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(acc[i]));
    s += a + b;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  float *out;
  cudaMalloc(&out, 148 * 8 * 1024 * sizeof(float));
  cudaMemset(out, 0, 148 * 8 * 1024 * sizeof(float));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  for (int warps = 4; warps <= 32; warps *= 2) {
    for (int variant = 0; variant < 2; ++variant) {
      float ms = 0.f;
      for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        if (variant == 0) k_scalar<8><<<148, warps * 32>>>(out, 0.5f, iters);
        else k_packed<8><<<148, warps * 32>>>(out, 0.5f, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
      }
      const double fma = (double)148 * warps * 32 * 8 * iters * (variant ? 2 : 1);
      const double per_clk_sm = fma / (ms * 1e-3) / 148 / (clk_khz * 1e3);
      printf("warps/SM %2d %s: %.3f ms, %.1f GFMA/s, %.1f FMA/clk/SM (at %d MHz nominal)\n", warps, variant ? "FFMA2 " : "FFMA  ", ms,
             fma / ms * 1e-6, per_clk_sm, clk_khz / 1000);
    }
  }
  return 0;
}
