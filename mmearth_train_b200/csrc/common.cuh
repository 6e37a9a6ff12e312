// Shared device helpers for the MP-MAE step kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdlib>
#include <utility>

namespace mpmae {

constexpr int kWarp = 32;

// Programmatic dependent launch.  Every kernel starts with pdl_prologue(): `launch_dependents` lets the NEXT launch of the
// stream be scheduled while this grid is still running (its CTAs become resident as ours drain), `wait` then blocks until
// every prerequisite grid has completed and its memory is visible -- nothing of a kernel runs before that, so the stream
// semantics are unchanged; what overlaps is launch latency, CTA scheduling and the drain of the previous grid.  Without
// the launch attribute both instructions are no-ops.  pdl(...) is the launch side: the triple-chevron launch of `kernel`
// with (g, b, s, st) is spelled pdl(kernel, g, b, s, st)(args); the attribute is set when MPMAE_PDL=1.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_prologue() {
  pdl_trigger();
  pdl_wait();
}
inline bool pdl_enabled() {
  static const bool on = [] { const char *e = getenv("MPMAE_PDL"); return e ? atoi(e) != 0 : false; }();
  return on;
}
template <typename K>
struct PdlLaunch {
  K kernel;
  dim3 grid, block;
  size_t smem;
  cudaStream_t stream;
  template <typename... Args>
  cudaError_t operator()(Args &&...args) const {
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
  }
};
template <typename K>
inline PdlLaunch<K> pdl(K kernel, dim3 grid, dim3 block, size_t smem, cudaStream_t stream) {
  return PdlLaunch<K>{kernel, grid, block, smem, stream};
}

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
// (row, part) LayerNorm mapping: NP adjacent lanes (a power of two <= 32) share one row of C channels; lane `part` owns the
// float4 columns part + NP*j, j < F4 = C / (4 NP).  Consecutive lanes touch consecutive 16-byte pieces, so global
// accesses coalesce and shared-memory accesses are conflict free, and the row reductions are log2(NP) shuffles instead of
// a full warp per row (C = 40 would leave most of a warp idle).
__device__ __forceinline__ float group_sum(float v, int np) {
  for (int o = np >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// NP for a channel count (C % 4 == 0): the largest power of two dividing C/4, at most 32
__host__ __device__ __forceinline__ int ln_parts(int C) {
  int c4 = C >> 2, np = 1;
  while (np < 32 && (c4 & 1) == 0) { c4 >>= 1; np <<= 1; }
  return np;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// erf GELU (torch.nn.GELU default, MinkowskiNonlinearity.py:113-114).  The normal CDF is evaluated with
// Abramowitz & Stegun 7.1.26 (|error| < 7.5e-8 on Phi, below fp32 rounding of x * Phi(x); measured max abs error of
// gelu 4.7e-7 on [-8, 8] against float64, tighter than libdevice erff's 1.2e-6 because the negative tail is formed
// without cancellation).  One rcp + one ex2 + 11 FMA-pipe instructions, branch free; the exponential doubles as the
// normal pdf for the derivative.
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// cdf = Phi(x), e = exp(-x^2 / 2)
__device__ __forceinline__ void normal_cdf_f(float x, float &cdf, float &e) {
  const float u = fabsf(x) * 0.849321800f;                  // |x| sqrt(log2(e) / 2):  u^2 = x^2 log2(e) / 2
  const float t = rcp_approx(fmaf(0.272737481f, u, 1.0f));  // 1 / (1 + 0.3275911 |x| / sqrt 2)
  float p = fmaf(0.5f * 1.061405429f, t, 0.5f * -1.453152027f);
  p = fmaf(p, t, 0.5f * 1.421413741f);
  p = fmaf(p, t, 0.5f * -0.284496736f);
  p = fmaf(p, t, 0.5f * 0.254829592f);
  p *= t;
  e = ex2_approx(-(u * u));                           // exp(-x^2 / 2)
  const float q = p * e;                             // 0.5 erfc(|x| / sqrt 2)
  cdf = x < 0.f ? q : 1.0f - q;
}
// gelu alone: x Phi(x) = max(x, 0) - |x| q with q = 0.5 erfc(|x| / sqrt 2): no sign select, two instructions fewer than
// forming Phi first (the forward epilogue and the operand splitters evaluate this for every element of [R, 4C])
__device__ __forceinline__ float gelu_f(float x) {
  const float ax = fabsf(x);
  const float u = ax * 0.849321800f;
  const float t = rcp_approx(fmaf(0.272737481f, u, 1.0f));
  float p = fmaf(0.5f * 1.061405429f, t, 0.5f * -1.453152027f);
  p = fmaf(p, t, 0.5f * 1.421413741f);
  p = fmaf(p, t, 0.5f * -0.284496736f);
  p = fmaf(p, t, 0.5f * 0.254829592f);
  p *= t;
  const float q = p * ex2_approx(-(u * u));
  return fmaf(-ax, q, fmaxf(x, 0.f));
}
__device__ __forceinline__ float gelu_grad_f(float x) {
  float cdf, e;
  normal_cdf_f(x, cdf, e);
  return fmaf(x * 0.39894228040143268f, e, cdf);
}

// ---- the same two functions for TWO values at a time with packed fp32 math (fma.rn.f32x2 / mul.rn.f32x2, sm_100): the
// polynomial, the products and the final combination take one instruction per PAIR; only |x|, max(x, 0), the sign and the two
// MUFU evaluations (rcp, ex2) stay scalar.  The GEMM epilogues and operand splitters that evaluate GELU for every element of
// an [R, 4C] tensor are bound by issue slots (ncu source page, profiles/r2_*): ~9 instead of ~14 instructions per value.
__device__ __forceinline__ unsigned long long pk2(float a, float b) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void upk2(unsigned long long v, float &a, float &b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ unsigned long long fma2_f(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ unsigned long long mul2_f(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// nq = -0.5 erfc(|x| / sqrt 2) for both values (NEGATED: the callers fold the sign into an FMA), e = exp(-x^2 / 2), ax = |x|
__device__ __forceinline__ void neg_half_erfc2_f(float x0, float x1, unsigned long long &nq, unsigned long long &e,
                                                 unsigned long long &ax) {
  ax = pk2(fabsf(x0), fabsf(x1));
  const unsigned long long u = mul2_f(ax, pk2(0.849321800f, 0.849321800f));
  const unsigned long long nu = mul2_f(ax, pk2(-0.849321800f, -0.849321800f));
  const unsigned long long d = fma2_f(u, pk2(0.272737481f, 0.272737481f), pk2(1.0f, 1.0f));
  float d0, d1;
  upk2(d, d0, d1);
  const unsigned long long t = pk2(rcp_approx(d0), rcp_approx(d1));
  unsigned long long p = fma2_f(pk2(-0.5f * 1.061405429f, -0.5f * 1.061405429f), t, pk2(0.5f * 1.453152027f, 0.5f * 1.453152027f));
  p = fma2_f(p, t, pk2(-0.5f * 1.421413741f, -0.5f * 1.421413741f));
  p = fma2_f(p, t, pk2(0.5f * 0.284496736f, 0.5f * 0.284496736f));
  p = fma2_f(p, t, pk2(-0.5f * 0.254829592f, -0.5f * 0.254829592f));
  p = mul2_f(p, t);
  float a0, a1;
  upk2(mul2_f(u, nu), a0, a1);                           // -u^2
  e = pk2(ex2_approx(a0), ex2_approx(a1));
  nq = mul2_f(p, e);
}
__device__ __forceinline__ void gelu2_f(float x0, float x1, float &h0, float &h1) {
  unsigned long long nq, e, ax;
  neg_half_erfc2_f(x0, x1, nq, e, ax);
  upk2(fma2_f(ax, nq, pk2(fmaxf(x0, 0.f), fmaxf(x1, 0.f))), h0, h1);     // max(x, 0) - |x| q
}
// h = gelu(x) and d gelu / dx for two values
__device__ __forceinline__ void gelu_both2_f(float x0, float x1, float &h0, float &h1, float &g0, float &g1) {
  unsigned long long nq, e, ax;
  neg_half_erfc2_f(x0, x1, nq, e, ax);
  const unsigned long long x = pk2(x0, x1);
  const unsigned long long sg = pk2(copysignf(1.0f, x0), copysignf(1.0f, x1));
  // Phi = 0.5 + sign(x) (0.5 - q) = 0.5 + sign(x) (0.5 + nq)
  const unsigned long long cdf = fma2_f(sg, fma2_f(nq, pk2(1.0f, 1.0f), pk2(0.5f, 0.5f)), pk2(0.5f, 0.5f));
  upk2(mul2_f(x, cdf), h0, h1);
  upk2(fma2_f(mul2_f(x, e), pk2(0.39894228040143268f, 0.39894228040143268f), cdf), g0, g1);
}

__device__ __forceinline__ float4 gelu4_f(const float4 &x) {
  float4 h;
  gelu2_f(x.x, x.y, h.x, h.y);
  gelu2_f(x.z, x.w, h.z, h.w);
  return h;
}
__device__ __forceinline__ void gelu_both4_f(const float4 &x, float4 &h, float4 &g) {
  gelu_both2_f(x.x, x.y, h.x, h.y, g.x, g.y);
  gelu_both2_f(x.z, x.w, h.z, h.w, g.z, g.w);
}

// h = gelu(x) and d gelu / dx from one evaluation
__device__ __forceinline__ void gelu_both_f(float x, float &h, float &dgelu) {
  float cdf, e;
  normal_cdf_f(x, cdf, e);
  h = x * cdf;
  dgelu = fmaf(x * 0.39894228040143268f, e, cdf);
}

// 3xBF16 weight operand: [rows][ceil(cols / 32)][32 bf16 high parts | 32 bf16 remainders], zero padded -- one 128-byte
// group per 32 columns, which is one row of a tcgen05 operand tile (gemm_tc.cuh).  Sizes of that array:
__host__ __device__ inline int bf16_pair_cols(int cols) { return ((cols + 31) / 32) * 64; }            // bf16 elements per row
__host__ __device__ inline int64_t bf16_pair_floats(int64_t rows, int cols) { return rows * (bf16_pair_cols(cols) / 2); }

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
