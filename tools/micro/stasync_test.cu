// Does st.async (STAS) to the CTA's own shared memory work, with / without a cluster launch, scalar / vector?
#include <cstdint>
#include <cstdio>
#include <cstdlib>
template <int VEC>
__device__ void body(float *out) {
  __shared__ __align__(16) float buf[128];
  __shared__ __align__(8) uint64_t bar;
  uint32_t a = (uint32_t)__cvta_generic_to_shared(buf + threadIdx.x * 4);
  uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar);
  if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
  __syncthreads();
  if (threadIdx.x == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 512;" ::"r"(b) : "memory");
  const float t = (float)threadIdx.x;
  if (VEC) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(a), "f"(t), "f"(t + .25f), "f"(t + .5f), "f"(t + .75f), "r"(b) : "memory");
  } else {
    for (int i = 0; i < 4; ++i)
      asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f32 [%0], %1, [%2];" ::"r"(a + 4 * i), "f"(t + .25f * i), "r"(b) : "memory");
  }
  uint32_t ok = 0;
  while (!ok) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0,1,0,p;}" : "=r"(ok) : "r"(b) : "memory");
  for (int i = 0; i < 4; ++i) out[threadIdx.x * 4 + i] = buf[threadIdx.x * 4 + i];
}
__global__ void k_plain_v(float *o) { body<1>(o); }
__global__ void k_plain_s(float *o) { body<0>(o); }
__global__ void __cluster_dims__(1, 1, 1) k_cluster_v(float *o) { body<1>(o); }
__global__ void __cluster_dims__(1, 1, 1) k_cluster_s(float *o) { body<0>(o); }
int main(int argc, char **argv) {
  const int only = argc > 1 ? atoi(argv[1]) : -1;
  float *d; cudaMalloc(&d, 512);
  float h[128];
  const char *names[4] = {"plain v4", "plain scalar", "cluster v4", "cluster scalar"};
  for (int v = 0; v < 4; ++v) {
    if (only >= 0 && v != only) continue;
    cudaMemset(d, 0, 512);
    if (v == 0) k_plain_v<<<1, 32>>>(d); else if (v == 1) k_plain_s<<<1, 32>>>(d); else if (v == 2) k_cluster_v<<<1, 32>>>(d); else k_cluster_s<<<1, 32>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", names[v], cudaGetErrorString(e)); return 1; }   // sticky: stop at the first failure
    cudaMemcpy(h, d, 512, cudaMemcpyDeviceToHost);
    int bad = 0; for (int i = 0; i < 128; ++i) bad += h[i] != (float)(i / 4) + .25f * (i % 4);
    printf("%s: ok, %d wrong\n", names[v], bad);
  }
}
