// Mask / slot tables, row-wise LayerNorm, GRN statistics, weight folding, decoder entry and
// pooled image-head kernels.  All are HBM-bound row kernels: one warp per row, float4 where the
// channel count allows, warp-shuffle reductions.
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

namespace mpmae {

// ------------------------------------------------------------------------------------------------
// M0: mask and slot tables from the noise tensor (models/fcmae.py:214-231).
//   rank(l) = position of patch l in the ascending (stable) sort of noise; kept iff rank < V.
__global__ void mask_kernel(const float *__restrict__ noise, float *__restrict__ mask, int *__restrict__ slot_of,
                            int *__restrict__ vis_patch, int L, int V) { pdl_prologue();
  extern __shared__ float sm[];
  float *nz = sm;
  int *keep = reinterpret_cast<int *>(sm + L);
  const int n = blockIdx.x;
  for (int l = threadIdx.x; l < L; l += blockDim.x) nz[l] = noise[n * L + l];
  __syncthreads();
  for (int l = threadIdx.x; l < L; l += blockDim.x) {
    const float v = nz[l];
    int rank = 0;
    for (int j = 0; j < L; ++j) {
      const float u = nz[j];
      rank += (u < v) || (u == v && j < l);
    }
    keep[l] = rank < V;
    mask[n * L + l] = rank < V ? 0.f : 1.f;
  }
  __syncthreads();
  for (int l = threadIdx.x; l < L; l += blockDim.x) {
    int s = 0;
    for (int j = 0; j < l; ++j) s += keep[j];
    if (keep[l]) { slot_of[n * L + l] = s; vis_patch[n * V + s] = l; }
    else slot_of[n * L + l] = -1;
  }
}

// ------------------------------------------------------------------------------------------------
// Row LayerNorm without affine: xhat = (x - mean) * rstd ; the affine part is folded into the
// weights of the GEMM that consumes xhat (fold_kernel).  eps = 1e-6 (sparse_norm_layers.py:61-77).
// Generic flavour (warp per row) and the (row, part) flavour of common.cuh for C = 4 * NP * F4.
__global__ void ln_rows_fwd_generic_kernel(const float *__restrict__ x, float *__restrict__ xhat, float *__restrict__ rstd_out,
                                           int64_t R, int C, float eps) { pdl_prologue();
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= R) return;
  const int lane = threadIdx.x & 31;
  const float *xr = x + r * C;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += xr[c];
  const float mean = warp_sum(s) / (float)C;
  float v = 0.f;
  for (int c = lane; c < C; c += 32) { const float d = xr[c] - mean; v += d * d; }
  const float rstd = rsqrtf(warp_sum(v) / (float)C + eps);
  for (int c = lane; c < C; c += 32) xhat[r * C + c] = (xr[c] - mean) * rstd;
  if (lane == 0) rstd_out[r] = rstd;
}
template <int F4>
__global__ void __launch_bounds__(256) ln_rows_fwd_kernel(const float *__restrict__ x, float *__restrict__ xhat,
                                                          float *__restrict__ rstd_out, int64_t R, int C, int np, float eps) { pdl_prologue();
  const int64_t item = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t r = item / np;
  const int part = (int)(item - r * np);
  const bool ok = r < R;
  const float4 *xr = reinterpret_cast<const float4 *>(x + r * C) + part;
  float4 v[F4];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < F4; ++j) {
    v[j] = ok ? __ldg(xr + j * np) : make_float4(0.f, 0.f, 0.f, 0.f);
    s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
  }
  const float mean = group_sum(s, np) / (float)C;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < F4; ++j) {
    v[j].x -= mean; v[j].y -= mean; v[j].z -= mean; v[j].w -= mean;
    q += v[j].x * v[j].x + v[j].y * v[j].y + v[j].z * v[j].z + v[j].w * v[j].w;
  }
  const float rstd = rsqrtf(group_sum(q, np) / (float)C + eps);
  if (!ok) return;
  float4 *o = reinterpret_cast<float4 *>(xhat + r * C) + part;
#pragma unroll
  for (int j = 0; j < F4; ++j) o[j * np] = make_float4(v[j].x * rstd, v[j].y * rstd, v[j].z * rstd, v[j].w * rstd);
  if (part == 0) rstd_out[r] = rstd;
}
inline void launch_ln_rows_fwd(const float *x, float *xhat, float *rstd, int64_t R, int C, float eps, cudaStream_t st) {
  const int np = ln_parts(C), f4 = (C >> 2) / np;
  const unsigned grid = (unsigned)cdiv64(R * np, 256);
  if ((C & 3) == 0 && f4 >= 1 && f4 <= 6) {
    switch (f4) {
      case 1: pdl(ln_rows_fwd_kernel<1>, grid, 256, 0, st)(x, xhat, rstd, R, C, np, eps); return;
      case 2: pdl(ln_rows_fwd_kernel<2>, grid, 256, 0, st)(x, xhat, rstd, R, C, np, eps); return;
      case 3: pdl(ln_rows_fwd_kernel<3>, grid, 256, 0, st)(x, xhat, rstd, R, C, np, eps); return;
      case 4: pdl(ln_rows_fwd_kernel<4>, grid, 256, 0, st)(x, xhat, rstd, R, C, np, eps); return;
      case 5: pdl(ln_rows_fwd_kernel<5>, grid, 256, 0, st)(x, xhat, rstd, R, C, np, eps); return;
      default: pdl(ln_rows_fwd_kernel<6>, grid, 256, 0, st)(x, xhat, rstd, R, C, np, eps); return;
    }
  }
  pdl(ln_rows_fwd_generic_kernel, (unsigned)cdiv64(R, 8), 256, 0, st)(x, xhat, rstd, R, C, eps);
}

// dx = rstd * (dxhat - mean_C(dxhat) - xhat * mean_C(dxhat * xhat))  (+ add)
__global__ void ln_rows_bwd_generic_kernel(const float *__restrict__ dxhat, const float *__restrict__ xhat,
                                           const float *__restrict__ rstd, const float *__restrict__ add,
                                           float *__restrict__ dx, int64_t R, int C) { pdl_prologue();
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= R) return;
  const int lane = threadIdx.x & 31;
  float s1 = 0.f, s2 = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float g = dxhat[r * C + c];
    s1 += g; s2 += g * xhat[r * C + c];
  }
  s1 = warp_sum(s1) / (float)C; s2 = warp_sum(s2) / (float)C;
  const float rs = rstd[r];
  for (int c = lane; c < C; c += 32) {
    float v = rs * (dxhat[r * C + c] - s1 - xhat[r * C + c] * s2);
    if (add) v += add[r * C + c];
    dx[r * C + c] = v;
  }
}
template <int F4>
__global__ void __launch_bounds__(256) ln_rows_bwd_kernel(const float *__restrict__ dxhat, const float *__restrict__ xhat,
                                                          const float *__restrict__ rstd, const float *__restrict__ add,
                                                          float *__restrict__ dx, int64_t R, int C, int np) { pdl_prologue();
  const int64_t item = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t r = item / np;
  const int part = (int)(item - r * np);
  const bool ok = r < R;
  const float4 *gr = reinterpret_cast<const float4 *>(dxhat + r * C) + part;
  const float4 *hr = reinterpret_cast<const float4 *>(xhat + r * C) + part;
  float4 g[F4], h[F4];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int j = 0; j < F4; ++j) {
    g[j] = ok ? __ldg(gr + j * np) : make_float4(0.f, 0.f, 0.f, 0.f);
    h[j] = ok ? __ldg(hr + j * np) : make_float4(0.f, 0.f, 0.f, 0.f);
    s1 += (g[j].x + g[j].y) + (g[j].z + g[j].w);
    s2 += g[j].x * h[j].x + g[j].y * h[j].y + g[j].z * h[j].z + g[j].w * h[j].w;
  }
  s1 = group_sum(s1, np) / (float)C;
  s2 = group_sum(s2, np) / (float)C;
  if (!ok) return;
  const float rs = __ldg(rstd + r);
  float4 *o = reinterpret_cast<float4 *>(dx + r * C) + part;
  const float4 *ar = add ? reinterpret_cast<const float4 *>(add + r * C) + part : nullptr;
#pragma unroll
  for (int j = 0; j < F4; ++j) {
    float4 v;
    v.x = rs * (g[j].x - s1 - h[j].x * s2); v.y = rs * (g[j].y - s1 - h[j].y * s2);
    v.z = rs * (g[j].z - s1 - h[j].z * s2); v.w = rs * (g[j].w - s1 - h[j].w * s2);
    if (ar) { const float4 a4 = __ldg(ar + j * np); v.x += a4.x; v.y += a4.y; v.z += a4.z; v.w += a4.w; }
    o[j * np] = v;
  }
}
inline void launch_ln_rows_bwd(const float *dxhat, const float *xhat, const float *rstd, const float *add, float *dx,
                               int64_t R, int C, cudaStream_t st) {
  const int np = ln_parts(C), f4 = (C >> 2) / np;
  const unsigned grid = (unsigned)cdiv64(R * np, 256);
  if ((C & 3) == 0 && f4 >= 1 && f4 <= 6) {
    switch (f4) {
      case 1: pdl(ln_rows_bwd_kernel<1>, grid, 256, 0, st)(dxhat, xhat, rstd, add, dx, R, C, np); return;
      case 2: pdl(ln_rows_bwd_kernel<2>, grid, 256, 0, st)(dxhat, xhat, rstd, add, dx, R, C, np); return;
      case 3: pdl(ln_rows_bwd_kernel<3>, grid, 256, 0, st)(dxhat, xhat, rstd, add, dx, R, C, np); return;
      case 4: pdl(ln_rows_bwd_kernel<4>, grid, 256, 0, st)(dxhat, xhat, rstd, add, dx, R, C, np); return;
      case 5: pdl(ln_rows_bwd_kernel<5>, grid, 256, 0, st)(dxhat, xhat, rstd, add, dx, R, C, np); return;
      default: pdl(ln_rows_bwd_kernel<6>, grid, 256, 0, st)(dxhat, xhat, rstd, add, dx, R, C, np); return;
    }
  }
  pdl(ln_rows_bwd_generic_kernel, (unsigned)cdiv64(R, 8), 256, 0, st)(dxhat, xhat, rstd, add, dx, R, C);
}

// ------------------------------------------------------------------------------------------------
// GRN (models/sparse_norm_layers.py:24-33 batch-global, eps 1e-6; models/norm_layers.py:41-44 per
// sample, eps 1e-4).  gsq[g, d] = sum over the group's rows of h^2 comes from the pw1 epilogue.
//   Gx = sqrt(gsq) ; Nx = Gx / (mean_d Gx + eps) ; scale s = 1 + gamma * Nx
__global__ void grn_scale_kernel(const float *__restrict__ gsq, const float *__restrict__ gamma,
                                 float *__restrict__ nx, float *__restrict__ scale, float *__restrict__ denom,
                                 int D, float eps) { pdl_prologue();
  __shared__ float red[32];
  const int g = blockIdx.x;
  float s = 0.f;
  for (int d = threadIdx.x; d < D; d += blockDim.x) s += sqrtf(gsq[(int64_t)g * D + d]);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) red[0] = t;
  }
  __syncthreads();
  const float den = red[0] / (float)D + eps;
  if (threadIdx.x == 0) denom[g] = den;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    const float n = sqrtf(gsq[(int64_t)g * D + d]) / den;
    nx[(int64_t)g * D + d] = n;
    scale[(int64_t)g * D + d] = 1.f + gamma[d] * n;
  }
}

// g = h * scale[group(r), d] + beta[d]
__global__ void grn_apply_kernel(const float *__restrict__ h, const float *__restrict__ scale,
                                 const float *__restrict__ beta, float *__restrict__ g, int64_t R, int D,
                                 int group_rows) { pdl_prologue();
  const int64_t n4 = R * (D >> 2);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / (D >> 2);
    const int d = (int)(i - r * (D >> 2)) * 4;
    const int64_t grp = r / group_rows;
    const float4 hv = *reinterpret_cast<const float4 *>(h + r * D + d);
    const float4 sv = *reinterpret_cast<const float4 *>(scale + grp * D + d);
    const float4 bv = *reinterpret_cast<const float4 *>(beta + d);
    float4 o;
    o.x = fmaf(hv.x, sv.x, bv.x); o.y = fmaf(hv.y, sv.y, bv.y);
    o.z = fmaf(hv.z, sv.z, bv.z); o.w = fmaf(hv.w, sv.w, bv.w);
    *reinterpret_cast<float4 *>(g + r * D + d) = o;
  }
}

// Backward of the GRN statistic.  ds[g, d] = sum_rows dg * h (from the dg epilogue).
//   dgamma[d] += sum_g Nx * ds ; dNx = gamma * ds ; dGx = dNx/den - (sum_j dNx_j Gx_j) / (D den^2)
//   kg[g, d] = dGx / Gx   (so that dh += kg * h)
__device__ __forceinline__ void grn_bwd_scale_body(const float *ds, const float *__restrict__ nx,
                                                   const float *__restrict__ denom, const float *__restrict__ gamma,
                                                   float *__restrict__ dgamma, float *__restrict__ kg, int D, int g) {
  __shared__ float red[32];
  const float den = denom[g];
  float s = 0.f;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    const int64_t i = (int64_t)g * D + d;
    s += gamma[d] * __ldcg(ds + i) * (nx[i] * den);  // dNx_j * Gx_j
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) red[0] = t;
  }
  __syncthreads();
  const float cross = red[0] / ((float)D * den * den);
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    const int64_t i = (int64_t)g * D + d;
    const float dsv = __ldcg(ds + i);
    const float dnx = gamma[d] * dsv;
    const float gx = nx[i] * den;
    const float dgx = dnx / den - cross;
    kg[i] = gx > 0.f ? dgx / gx : 0.f;
    atomicAdd(&dgamma[d], nx[i] * dsv);
  }
}
__global__ void grn_bwd_scale_kernel(const float *__restrict__ ds, const float *__restrict__ nx,
                                     const float *__restrict__ denom, const float *__restrict__ gamma,
                                     float *__restrict__ dgamma, float *__restrict__ kg, int D) { pdl_prologue();
  grn_bwd_scale_body(ds, nx, denom, gamma, dgamma, kg, D, blockIdx.x);
}

// h' = gelu(a) * scale[d]: materialises the pw2 operand for the GEMM backends without operand-splitter warps (fp32 SIMT,
// single-pass TF32); the split backends apply it on the way into the tensor core instead (GemmArgs::a_gelu)
__global__ void gelu_scale_rows_kernel(const float *__restrict__ a, const float *__restrict__ scale, float *__restrict__ out,
                                       int64_t R, int D) { pdl_prologue();
  const int64_t n4 = R * (D >> 2);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const int d = (int)(i % (D >> 2)) * 4;
    const float4 av = *reinterpret_cast<const float4 *>(a + i * 4);
    const float4 sv = scale ? *reinterpret_cast<const float4 *>(scale + d) : make_float4(1.f, 1.f, 1.f, 1.f);
    const float4 hv = gelu4_f(av);
    *reinterpret_cast<float4 *>(out + i * 4) = make_float4(hv.x * sv.x, hv.y * sv.y, hv.z * sv.z, hv.w * sv.w);
  }
}

// da = (dg * scale[g, d] + kg[g, d] * h) * gelu'(a)   (in place over dg allowed)
__global__ void grn_gelu_bwd_kernel(const float *dg, const float *__restrict__ h, const float *__restrict__ a,
                                    const float *__restrict__ scale, const float *__restrict__ kg, float *da,
                                    int64_t R, int D, int group_rows) { pdl_prologue();
  const int64_t n4 = R * (D >> 2);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / (D >> 2);
    const int d = (int)(i - r * (D >> 2)) * 4;
    const int64_t grp = r / group_rows;
    const float4 gv = *reinterpret_cast<const float4 *>(dg + r * D + d);
    const float4 hv = *reinterpret_cast<const float4 *>(h + r * D + d);
    const float4 av = *reinterpret_cast<const float4 *>(a + r * D + d);
    const float4 sv = *reinterpret_cast<const float4 *>(scale + grp * D + d);
    const float4 kv = *reinterpret_cast<const float4 *>(kg + grp * D + d);
    float4 o;
    o.x = (gv.x * sv.x + kv.x * hv.x) * gelu_grad_f(av.x);
    o.y = (gv.y * sv.y + kv.y * hv.y) * gelu_grad_f(av.y);
    o.z = (gv.z * sv.z + kv.z * hv.z) * gelu_grad_f(av.z);
    o.w = (gv.w * sv.w + kv.w * hv.w) * gelu_grad_f(av.w);
    *reinterpret_cast<float4 *>(da + r * D + d) = o;
  }
}

// ------------------------------------------------------------------------------------------------
// Weight folding.  A LayerNorm affine (scale w, shift b over the K axis) in front of a linear map
// y = (xhat*w + b) . W^T + bias  is absorbed into  Wf[n,k] = W[n,k]*w[k%SL],  bf[n] = bias[n] +
// sum_k W[n,k]*b[k%SL].  Source W may be stored [N,K] (nn.Linear) or [K,N] (ME conv kernels) via
// strides.  Emits Wf [N,K] and/or WfT [K,N].  One warp per output row n.
struct FoldArgs {
  const float *W; int64_t s_n, s_k;
  const float *scale_k;  // [SL] or null
  const float *shift_k;  // [SL] or null
  const float *scale_n;  // [N] or null (row scale, e.g. per-modality loss seeds)
  const float *bias;     // [N] or null
  float *Wf;             // [N, K] or null
  float *WfT;            // [K, N] or null
  float *bf;             // [N] or null
  int N, K, SL;
  // 3xTF32: when set, Wf / WfT receive the TF32-representable high part and these the remainder (w - hi)
  float *Wf_lo;
  float *WfT_lo;
  // 3xBF16 (b16 != 0): Wf / WfT receive the full fp32 value and Wf_lo / WfT_lo are reinterpreted as packed bf16 pairs:
  // hi part [N*K] followed by lo part [N*K] (same footprint as one fp32 array)
  int b16;
  // fused batch-global GRN statistic (sparse blocks): when gsq is set the K-axis scale is computed here,
  //   s[k] = 1 + gamma[k] * sqrt(gsq[k]) / (mean_k sqrt(gsq) + eps),  and nx / scale / denom are written for the backward
  const float *gsq, *gamma;
  float *nx_out, *scale_out, *denom_out;
  float grn_eps;
};
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }
// element (row, col) of the interleaved bf16 pair array (common.cuh: bf16_pair_cols); col may lie in the zero padding
__device__ __forceinline__ void store_bf16_pair(float *pair_array, int row, int col, int cols, float wf) {
  uint16_t *b = reinterpret_cast<uint16_t *>(pair_array) + (int64_t)row * bf16_pair_cols(cols) + (col >> 5) * 64 + (col & 31);
  const __nv_bfloat16 h = __float2bfloat16_rn(wf);
  const __nv_bfloat16 l = __float2bfloat16_rn(wf - __bfloat162float(h));
  b[0] = __bfloat16_as_ushort(h);
  b[32] = __bfloat16_as_ushort(l);
}
// 32 x 32 tiles through shared memory: the source is read along its contiguous axis and both orientations (Wf [N, K],
// WfT [K, N]) leave with coalesced stores.  grid = (K tiles, N tiles), block = (32, 8).  The folded bias of a row block is
// produced by the CTAs of the first K tile.
__device__ __forceinline__ void fold_tile(const FoldArgs &p, int bx, int by) {
  __shared__ float tile[32][33];
  __shared__ float red[8];
  __shared__ float sks[32];
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * 32 + tx;
  const int k0 = bx * 32, n0 = by * 32;
  if (p.gsq) {   // every CTA recomputes the (tiny) GRN denominator; K = SL here
    float s = 0.f;
    for (int d = tid; d < p.K; d += 256) s += sqrtf(p.gsq[d]);
    s = warp_sum(s);
    if (tx == 0) red[ty] = s;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i];
    const float den = t / (float)p.K + p.grn_eps;
    if (tid < 32) {
      const int k = k0 + tid;
      float sk = 1.f;
      if (k < p.K) {
        const float nx = sqrtf(p.gsq[k]) / den;
        sk = 1.f + p.gamma[k] * nx;
        if (by == 0) { p.nx_out[k] = nx; p.scale_out[k] = sk; }
      }
      sks[tid] = sk;
    }
    if (bx == 0 && by == 0 && tid == 0) p.denom_out[0] = den;
  } else if (tid < 32) {
    const int k = k0 + tid;
    sks[tid] = (p.scale_k && k < p.K) ? p.scale_k[k % p.SL] : 1.f;
  }
  // source tile -> tile[n][k]
  if (p.s_k == 1) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int n = n0 + ty + 8 * i, k = k0 + tx;
      tile[ty + 8 * i][tx] = (n < p.N && k < p.K) ? p.W[n * p.s_n + k] : 0.f;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int n = n0 + tx, k = k0 + ty + 8 * i;
      tile[tx][ty + 8 * i] = (n < p.N && k < p.K) ? p.W[n * p.s_n + k * p.s_k] : 0.f;
    }
  }
  __syncthreads();
  const bool split = (p.Wf_lo != nullptr || p.WfT_lo != nullptr) && !p.b16;
  if (p.Wf) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int nl = ty + 8 * i, n = n0 + nl, k = k0 + tx;
      if (n < p.N && k < p.K) {
        const float wf = tile[nl][tx] * sks[tx] * (p.scale_n ? p.scale_n[n] : 1.f);
        const float hi = tf32_hi(wf);
        p.Wf[(int64_t)n * p.K + k] = split ? hi : wf;
        if (p.Wf_lo) {
          if (p.b16) store_bf16_pair(p.Wf_lo, n, k, p.K, wf);
          else p.Wf_lo[(int64_t)n * p.K + k] = wf - hi;
        }
      } else if (n < p.N && p.Wf_lo && p.b16) {
        store_bf16_pair(p.Wf_lo, n, k, p.K, 0.f);     // zero padding of the last 32-column group (the tile covers it)
      }
    }
  }
  if (p.WfT) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int kl = ty + 8 * i, k = k0 + kl, n = n0 + tx;
      if (n < p.N && k < p.K) {
        const float wf = tile[tx][kl] * sks[kl] * (p.scale_n ? p.scale_n[n] : 1.f);
        const float hi = tf32_hi(wf);
        p.WfT[(int64_t)k * p.N + n] = split ? hi : wf;
        if (p.WfT_lo) {
          if (p.b16) store_bf16_pair(p.WfT_lo, k, n, p.N, wf);
          else p.WfT_lo[(int64_t)k * p.N + n] = wf - hi;
        }
      } else if (k < p.K && p.WfT_lo && p.b16) {
        store_bf16_pair(p.WfT_lo, k, n, p.N, 0.f);
      }
    }
  }
  if (p.bf && bx == 0) {   // bf[n] = bias[n] + sum_k W[n, k] * shift[k % SL]: warp ty owns rows ty, ty+8, ...
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int n = n0 + ty + 8 * i;
      if (n >= p.N) continue;
      float acc = 0.f;
      if (p.shift_k)
        for (int k = tx; k < p.K; k += 32) acc = fmaf(p.W[n * p.s_n + k * p.s_k], p.shift_k[k % p.SL], acc);
      acc = warp_sum(acc);
      if (tx == 0) p.bf[n] = (p.bias ? p.bias[n] : 0.f) + acc;
    }
  }
}
__global__ void __launch_bounds__(256) fold_kernel(FoldArgs p) { pdl_prologue(); fold_tile(p, blockIdx.x, blockIdx.y); }
inline void launch_fold(const FoldArgs &a, cudaStream_t st) {
  pdl(fold_kernel, dim3(cdiv(a.K, 32), cdiv(a.N, 32)), dim3(32, 8), 0, st)(a);
}
// Many folds in one launch (all the parameter-only folds of a forward pass): jobs and the prefix sum of their tile
// counts live in device memory; a CTA finds its job by scanning the (short) prefix array.
__global__ void __launch_bounds__(256) fold_batch_kernel(const FoldArgs *__restrict__ jobs, const int *__restrict__ tile_start,
                                                         int njobs) { pdl_prologue();
  __shared__ FoldArgs job;
  __shared__ int local;
  if (threadIdx.x == 0 && threadIdx.y == 0) {
    int j = 0;
    while (j + 1 < njobs && (int)blockIdx.x >= tile_start[j + 1]) ++j;
    job = jobs[j];
    local = (int)blockIdx.x - tile_start[j];
  }
  __syncthreads();
  const int kt = cdiv(job.K, 32);
  fold_tile(job, local % kt, local / kt);
}

// Chain rule back through the fold:  given dWf [N,K] and dbf [N]
//   dW[n,k]    += dWf[n,k]*scale[k%SL] + dbf[n]*shift[k%SL]
//   dscale[j]  += sum_{n, k%SL==j} dWf[n,k]*W[n,k]
//   dshift[j]  += sum_{n, k%SL==j} dbf[n]*W[n,k]
//   dbias[n]   += dbf[n]
// One CTA per k; threads stride n.
struct UnfoldArgs {
  const float *W; int64_t s_n, s_k;
  const float *scale_k, *shift_k;
  const float *dWf, *dbf;
  float *dW;       // same strides as W
  float *dscale, *dshift, *dbias;
  int N, K, SL;
};
__device__ __forceinline__ void unfold_column(const UnfoldArgs &p, int k) {
  __shared__ float r1[32], r2[32];
  const int j = k % p.SL;
  const float sc = p.scale_k ? p.scale_k[j] : 1.f;
  const float sh = p.shift_k ? p.shift_k[j] : 0.f;
  float a1 = 0.f, a2 = 0.f;
  for (int n = threadIdx.x; n < p.N; n += blockDim.x) {
    const float w = p.W[n * p.s_n + k * p.s_k];
    const float g = p.dWf[(int64_t)n * p.K + k];
    const float gb = p.dbf ? p.dbf[n] : 0.f;
    atomicAdd(&p.dW[n * p.s_n + k * p.s_k], g * sc + gb * sh);
    a1 = fmaf(g, w, a1);
    a2 = fmaf(gb, w, a2);
    if (k == 0 && p.dbias) atomicAdd(&p.dbias[n], gb);
  }
  a1 = warp_sum(a1); a2 = warp_sum(a2);
  if ((threadIdx.x & 31) == 0) { r1[threadIdx.x >> 5] = a1; r2[threadIdx.x >> 5] = a2; }
  __syncthreads();
  if (threadIdx.x < 32) {
    float t1 = threadIdx.x < (blockDim.x >> 5) ? r1[threadIdx.x] : 0.f;
    float t2 = threadIdx.x < (blockDim.x >> 5) ? r2[threadIdx.x] : 0.f;
    t1 = warp_sum(t1); t2 = warp_sum(t2);
    if (threadIdx.x == 0) {
      if (p.dscale) atomicAdd(&p.dscale[j], t1);
      if (p.dshift) atomicAdd(&p.dshift[j], t2);
    }
  }
}
__global__ void unfold_kernel(UnfoldArgs p) { pdl_prologue(); unfold_column(p, blockIdx.x); }
// The un-fold of a sparse block's pw2 (grid = K = 4C columns) followed, in the LAST CTA to finish, by the backward of the
// batch-global GRN statistic, which needs the complete dscale vector (= A_c = sum_r dg * h).
struct GrnBwdArgs {
  const float *nx, *denom, *gamma;
  float *dgamma, *kg;
  unsigned int *counter;   // zero before the launch; left at zero
};
__global__ void unfold_grn_kernel(UnfoldArgs p, GrnBwdArgs q) { pdl_prologue();
  __shared__ unsigned int last;
  unfold_column(p, blockIdx.x);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = (atomicAdd(q.counter, 1u) == gridDim.x - 1) ? 1u : 0u;
  __syncthreads();
  if (!last) return;
  __threadfence();
  grn_bwd_scale_body(p.dscale, q.nx, q.denom, q.gamma, q.dgamma, q.kg, p.K, 0);
  if (threadIdx.x == 0) *q.counter = 0u;
}
// Several un-folds in one launch (the deferred ones of a backward part): job table + prefix sum of the K's in device memory
__global__ void unfold_batch_kernel(const UnfoldArgs *__restrict__ jobs, const int *__restrict__ k_start, int njobs) { pdl_prologue();
  __shared__ UnfoldArgs job;
  __shared__ int local;
  if (threadIdx.x == 0) {
    int j = 0;
    while (j + 1 < njobs && (int)blockIdx.x >= k_start[j + 1]) ++j;
    job = jobs[j];
    local = (int)blockIdx.x - k_start[j];
  }
  __syncthreads();
  unfold_column(job, local);
}

// out[r, c] = x[r, c] * cs[c] (the per-column loss seeds applied to the raw prediction gradient); fill: v everywhere
__global__ void scale_cols_kernel(const float *__restrict__ x, const float *__restrict__ cs, float *__restrict__ out, int64_t R,
                                  int C) { pdl_prologue();
  const int64_t n = R * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = x[i] * cs[i % C];
}
__global__ void fill_kernel(float *__restrict__ x, float v, int64_t n) { pdl_prologue();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) x[i] = v;
}

// ------------------------------------------------------------------------------------------------
// Decoder entry (models/fcmae.py:251-255): dense cell rows from projected visible rows + mask token.
//   xd[n*L + l, :] = slot>=0 ? z[n*V+slot, :] : token
__global__ void scatter_token_kernel(const float *__restrict__ z, const float *__restrict__ token,
                                     const int *__restrict__ slot_of, float *__restrict__ xd, int64_t cells, int L,
                                     int V, int C) { pdl_prologue();
  const int C4 = C >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cells * C4; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t cell = i / C4;
    const int c = (int)(i - cell * C4) * 4;
    const int slot = slot_of[cell];
    const int64_t n = cell / L;
    const float4 v = slot >= 0 ? *reinterpret_cast<const float4 *>(z + (n * V + slot) * C + c)
                               : *reinterpret_cast<const float4 *>(token + c);
    *reinterpret_cast<float4 *>(xd + cell * C + c) = v;
  }
}
// backward: dz[n*V+slot] = dxd[cell] for visible cells; dtoken += sum over masked cells.
// blockDim.x = C/4: a thread owns one float4 column group for every cell its CTA visits (token sums in registers).
__global__ void gather_token_bwd_kernel(const float *__restrict__ dxd, const int *__restrict__ slot_of,
                                        float *__restrict__ dz, float *__restrict__ dtoken, int64_t cells, int L, int V,
                                        int C) { pdl_prologue();
  const int C4 = C >> 2, t = threadIdx.x;
  if (t >= C4) return;
  const float4 *src = reinterpret_cast<const float4 *>(dxd);
  float4 *dst = reinterpret_cast<float4 *>(dz);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t cell = blockIdx.x; cell < cells; cell += gridDim.x) {
    const int slot = __ldg(slot_of + cell);
    const int64_t n = cell / L;
    const float4 v = __ldg(src + cell * C4 + t);
    if (slot >= 0) dst[(n * V + slot) * C4 + t] = v;
    else { acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
  }
  atomicAdd(&dtoken[4 * t], acc.x); atomicAdd(&dtoken[4 * t + 1], acc.y);
  atomicAdd(&dtoken[4 * t + 2], acc.z); atomicAdd(&dtoken[4 * t + 3], acc.w);
}

// ------------------------------------------------------------------------------------------------
// Image-level heads: channel LayerNorm of every decoder cell (norm_layers.py:26-31, eps 1e-6) then
// mean over the L cells (fcmae.py:259-262).  One CTA per sample, warp per cell; a lane owns the float4 columns
// lane + 32 j of every row its warp visits, so the per-channel sums over cells live in registers and only the
// per-warp partials meet in shared memory.  C % 128 == 0, C <= 1024.
constexpr int kPoolMaxF4 = 8;
__global__ void __launch_bounds__(256) pool_ln_fwd_kernel(const float *__restrict__ d, const float *__restrict__ w,
                                                          const float *__restrict__ b, float *__restrict__ pooled,
                                                          float *__restrict__ rstd_out, int L, int C, float eps) { pdl_prologue();
  extern __shared__ float part[];  // [nw][C]
  const int n = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int f4 = C >> 7;           // float4 per lane
  float4 acc[kPoolMaxF4];
#pragma unroll
  for (int j = 0; j < kPoolMaxF4; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int l = warp; l < L; l += nw) {
    const float4 *xr = reinterpret_cast<const float4 *>(d + ((int64_t)n * L + l) * C) + lane;
    float4 v[kPoolMaxF4];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < kPoolMaxF4; ++j)
      if (j < f4) { v[j] = __ldg(xr + 32 * j); s += (v[j].x + v[j].y) + (v[j].z + v[j].w); }
    const float mean = warp_sum(s) / (float)C;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < kPoolMaxF4; ++j)
      if (j < f4) {
        v[j].x -= mean; v[j].y -= mean; v[j].z -= mean; v[j].w -= mean;
        q += v[j].x * v[j].x + v[j].y * v[j].y + v[j].z * v[j].z + v[j].w * v[j].w;
      }
    const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
    if (lane == 0) rstd_out[(int64_t)n * L + l] = rstd;
#pragma unroll
    for (int j = 0; j < kPoolMaxF4; ++j)
      if (j < f4) { acc[j].x += v[j].x * rstd; acc[j].y += v[j].y * rstd; acc[j].z += v[j].z * rstd; acc[j].w += v[j].w * rstd; }
  }
#pragma unroll
  for (int j = 0; j < kPoolMaxF4; ++j)
    if (j < f4) reinterpret_cast<float4 *>(part + (size_t)warp * C)[lane + 32 * j] = acc[j];
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float t = 0.f;
    for (int k = 0; k < nw; ++k) t += part[(size_t)k * C + c];
    pooled[(int64_t)n * C + c] = w[c] * (t / (float)L) + b[c];
  }
}
// backward: dpooled [B,C] -> ddec[n,l,:] += LNbwd(dn = w*dpooled/L) ; dw += dpooled * mean_l(nhat) ; db += dpooled
__global__ void __launch_bounds__(256) pool_ln_bwd_kernel(const float *__restrict__ d, const float *__restrict__ rstd_in,
                                                          const float *__restrict__ w, const float *__restrict__ dpooled,
                                                          float *__restrict__ ddec, float *__restrict__ dw,
                                                          float *__restrict__ db, int L, int C, float eps) { pdl_prologue();
  extern __shared__ float part[];  // [nw][C] sum_l nhat
  const int n = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int f4 = C >> 7;
  float4 acc[kPoolMaxF4], g[kPoolMaxF4];
  float s1 = 0.f;
#pragma unroll
  for (int j = 0; j < kPoolMaxF4; ++j) {
    acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    g[j] = acc[j];
    if (j < f4) {   // g = w * dpooled / L is the same for every cell of the sample
      const float4 wv = __ldg(reinterpret_cast<const float4 *>(w) + lane + 32 * j);
      const float4 dv = __ldg(reinterpret_cast<const float4 *>(dpooled + (int64_t)n * C) + lane + 32 * j);
      g[j] = make_float4(wv.x * dv.x / (float)L, wv.y * dv.y / (float)L, wv.z * dv.z / (float)L, wv.w * dv.w / (float)L);
      s1 += (g[j].x + g[j].y) + (g[j].z + g[j].w);
    }
  }
  s1 = warp_sum(s1) / (float)C;
  for (int l = warp; l < L; l += nw) {
    const int64_t row = (int64_t)n * L + l;
    const float4 *xr = reinterpret_cast<const float4 *>(d + row * C) + lane;
    float4 v[kPoolMaxF4];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < kPoolMaxF4; ++j)
      if (j < f4) { v[j] = __ldg(xr + 32 * j); s += (v[j].x + v[j].y) + (v[j].z + v[j].w); }
    const float mean = warp_sum(s) / (float)C;
    const float rstd = __ldg(rstd_in + row);
    float s2 = 0.f;
#pragma unroll
    for (int j = 0; j < kPoolMaxF4; ++j)
      if (j < f4) {
        v[j].x = (v[j].x - mean) * rstd; v[j].y = (v[j].y - mean) * rstd;
        v[j].z = (v[j].z - mean) * rstd; v[j].w = (v[j].w - mean) * rstd;
        s2 += g[j].x * v[j].x + g[j].y * v[j].y + g[j].z * v[j].z + g[j].w * v[j].w;
        acc[j].x += v[j].x; acc[j].y += v[j].y; acc[j].z += v[j].z; acc[j].w += v[j].w;
      }
    s2 = warp_sum(s2) / (float)C;
    float4 *o = reinterpret_cast<float4 *>(ddec + row * C) + lane;
#pragma unroll
    for (int j = 0; j < kPoolMaxF4; ++j)
      if (j < f4) {
        float4 t = o[32 * j];
        t.x += rstd * (g[j].x - s1 - v[j].x * s2); t.y += rstd * (g[j].y - s1 - v[j].y * s2);
        t.z += rstd * (g[j].z - s1 - v[j].z * s2); t.w += rstd * (g[j].w - s1 - v[j].w * s2);
        o[32 * j] = t;
      }
  }
#pragma unroll
  for (int j = 0; j < kPoolMaxF4; ++j)
    if (j < f4) reinterpret_cast<float4 *>(part + (size_t)warp * C)[lane + 32 * j] = acc[j];
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float t = 0.f;
    for (int k = 0; k < nw; ++k) t += part[(size_t)k * C + c];
    const float dp = dpooled[(int64_t)n * C + c];
    atomicAdd(&dw[c], dp * t / (float)L);
    atomicAdd(&db[c], dp);
  }
}

// ------------------------------------------------------------------------------------------------
// Dense encoder features [B, C, G, G] from the stage-3 rows (SparseTensor.dense(), zeros at masked cells)
__global__ void densify_kernel(const float *__restrict__ x3, const int *__restrict__ slot_of, float *__restrict__ out,
                               int B, int L, int V, int C) { pdl_prologue();
  const int64_t total = (int64_t)B * C * L;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int l = (int)(i % L);
    const int c = (int)((i / L) % C);
    const int64_t n = i / ((int64_t)L * C);
    const int slot = slot_of[n * L + l];
    out[i] = slot >= 0 ? x3[(n * V + slot) * C + c] : 0.f;
  }
}

// column sums of a [R, C] matrix into out[C] (+=): bias gradients that have no GEMM to ride on.  The matrix is read as
// a flat float4 stream: blockDim.x = (C/4 float4 column groups of this chunk) x (rows per pass), so consecutive threads
// read consecutive addresses whatever C is (C = 40 is 10 float4 per row), and a thread always owns the same 4 columns.
// gridDim.y walks column chunks of 2048; gridDim.x strides row passes.
__global__ void __launch_bounds__(512) colsum_kernel(const float *__restrict__ x, const float *__restrict__ rs,
                                                     float *__restrict__ out, int64_t R, int C) { pdl_prologue();
  __shared__ float4 red[512];
  const int c4_total = C >> 2;
  const int c4_0 = blockIdx.y * 512;
  const int c4n = min(512, c4_total - c4_0);          // float4 groups in this chunk
  const int rpp = blockDim.x / c4n;                   // rows per pass
  const int tr = threadIdx.x / c4n, tc = threadIdx.x - tr * c4n;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (tr < rpp) {
    const float4 *xp = reinterpret_cast<const float4 *>(x) + c4_0 + tc;
    const int64_t step = (int64_t)gridDim.x * rpp;
    int64_t r = (int64_t)blockIdx.x * rpp + tr;
    float4 acc1 = acc, acc2 = acc, acc3 = acc;
    for (; r + 3 * step < R; r += 4 * step) {   // four independent loads in flight per thread: the pass is latency bound
      const float4 v0 = __ldg(xp + r * c4_total), v1 = __ldg(xp + (r + step) * c4_total);
      const float4 v2 = __ldg(xp + (r + 2 * step) * c4_total), v3 = __ldg(xp + (r + 3 * step) * c4_total);
      acc.x += v0.x; acc.y += v0.y; acc.z += v0.z; acc.w += v0.w;
      acc1.x += v1.x; acc1.y += v1.y; acc1.z += v1.z; acc1.w += v1.w;
      acc2.x += v2.x; acc2.y += v2.y; acc2.z += v2.z; acc2.w += v2.w;
      acc3.x += v3.x; acc3.y += v3.y; acc3.z += v3.z; acc3.w += v3.w;
    }
    for (; r < R; r += step) {
      const float4 v = __ldg(xp + r * c4_total);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    acc.x += (acc1.x + acc2.x) + acc3.x; acc.y += (acc1.y + acc2.y) + acc3.y;
    acc.z += (acc1.z + acc2.z) + acc3.z; acc.w += (acc1.w + acc2.w) + acc3.w;
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x < c4n) {
    float4 t = red[threadIdx.x];
    for (int i = 1; i < rpp; ++i) {
      const float4 u = red[i * c4n + threadIdx.x];
      t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
    }
    const int c = (c4_0 + threadIdx.x) * 4;
    if (rs) { t.x *= rs[c]; t.y *= rs[c + 1]; t.z *= rs[c + 2]; t.w *= rs[c + 3]; }
    atomicAdd(&out[c], t.x); atomicAdd(&out[c + 1], t.y); atomicAdd(&out[c + 2], t.z); atomicAdd(&out[c + 3], t.w);
  }
}
// launch helper: C % 4 == 0, x 16-byte aligned
inline void launch_colsum(const float *x, const float *rs, float *out, int64_t R, int C, cudaStream_t st) {
  const int c4 = C >> 2;
  const int chunks = cdiv(c4, 512);
  const int c4n = c4 < 512 ? c4 : 512;
  int threads = (512 / c4n) * c4n;
  threads = ((threads + 31) / 32) * 32;
  if (threads > 512) threads = 512;
  const int rpp = threads / c4n > 0 ? threads / c4n : 1;
  int64_t gx = cdiv64(R, (int64_t)rpp * 8);            // >= 8 passes per CTA
  if (gx > 148 * 4 / chunks) gx = chunks >= 4 ? 148 : 148 * 4 / chunks;
  if (gx < 1) gx = 1;
  pdl(colsum_kernel, dim3((unsigned)gx, chunks), threads, 0, st)(x, rs, out, R, C);
}

// Small ragged products of the image-level heads (256 x 878 x 512 at cfg2; N = 878 is not a multiple of 4, so they do not
// go through the tensor-core path).  32 x 32 output tiles (hundreds of CTAs even at this size), 2 x 2 outputs per thread.
//   NT: out[b, k] = sum_n A[b, n] * cs[n] * W[n, k]          (dX of the heads;   A [Brows, Nred], W [Nred, Kout])
//   NN: out[m, n] = bias[n] + sum_k A[m, k] * W[n, k]        (the heads forward; A [M, K],      W [N, K])
__global__ void __launch_bounds__(256) small_gemm_nt_kernel(const float *__restrict__ A, const float *__restrict__ W,
                                                            const float *__restrict__ cs, float *__restrict__ out, int Brows,
                                                            int Kout, int Nred) { pdl_prologue();
  __shared__ float As[32][33];   // [n][b]
  __shared__ float Ws[32][33];   // [n][k]
  const int b0 = blockIdx.y * 32, k0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;   // tx: 2 columns, ty: 2 rows
  float acc[2][2] = {};
  float ra[4], rw[4];   // next tile's operands, fetched while the current tile is being multiplied
  auto fetch = [&](int n0) {
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int i = threadIdx.x + t * 256, r = i >> 5, q = i & 31;
      ra[t] = (b0 + r < Brows && n0 + q < Nred) ? A[(int64_t)(b0 + r) * Nred + n0 + q] * cs[n0 + q] : 0.f;
      rw[t] = (n0 + r < Nred && k0 + q < Kout) ? W[(int64_t)(n0 + r) * Kout + k0 + q] : 0.f;
    }
  };
  fetch(0);
  for (int n0 = 0; n0 < Nred; n0 += 32) {
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int i = threadIdx.x + t * 256, r = i >> 5, q = i & 31;
      As[q][r] = ra[t];
      Ws[r][q] = rw[t];
    }
    __syncthreads();
    if (n0 + 32 < Nred) fetch(n0 + 32);
#pragma unroll 8
    for (int n = 0; n < 32; ++n) {
      const float a0 = As[n][ty * 2], a1 = As[n][ty * 2 + 1];
      const float w0 = Ws[n][tx * 2], w1 = Ws[n][tx * 2 + 1];
      acc[0][0] = fmaf(a0, w0, acc[0][0]); acc[0][1] = fmaf(a0, w1, acc[0][1]);
      acc[1][0] = fmaf(a1, w0, acc[1][0]); acc[1][1] = fmaf(a1, w1, acc[1][1]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int b = b0 + ty * 2 + i, k = k0 + tx * 2 + j;
      if (b < Brows && k < Kout) out[(int64_t)b * Kout + k] = acc[i][j];
    }
}
__global__ void __launch_bounds__(256) small_gemm_nn_kernel(const float *__restrict__ A, const float *__restrict__ W,
                                                            const float *__restrict__ bias, float *__restrict__ out, int M,
                                                            int N, int K) { pdl_prologue();
  __shared__ float As[32][33];   // [k][m]
  __shared__ float Ws[32][33];   // [k][n]
  const int m0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[2][2] = {};
  float ra[4], rw[4];
  auto fetch = [&](int k0) {   // q runs along k: coalesced reads of both row-major operands
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int i = threadIdx.x + t * 256, r = i >> 5, q = i & 31;
      ra[t] = (m0 + r < M && k0 + q < K) ? A[(int64_t)(m0 + r) * K + k0 + q] : 0.f;
      rw[t] = (n0 + r < N && k0 + q < K) ? W[(int64_t)(n0 + r) * K + k0 + q] : 0.f;
    }
  };
  fetch(0);
  for (int k0 = 0; k0 < K; k0 += 32) {
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int i = threadIdx.x + t * 256, r = i >> 5, q = i & 31;
      As[q][r] = ra[t];
      Ws[q][r] = rw[t];
    }
    __syncthreads();
    if (k0 + 32 < K) fetch(k0 + 32);
#pragma unroll 8
    for (int k = 0; k < 32; ++k) {
      const float a0 = As[k][ty * 2], a1 = As[k][ty * 2 + 1];
      const float w0 = Ws[k][tx * 2], w1 = Ws[k][tx * 2 + 1];
      acc[0][0] = fmaf(a0, w0, acc[0][0]); acc[0][1] = fmaf(a0, w1, acc[0][1]);
      acc[1][0] = fmaf(a1, w0, acc[1][0]); acc[1][1] = fmaf(a1, w1, acc[1][1]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int m = m0 + ty * 2 + i, n = n0 + tx * 2 + j;
      if (m < M && n < N) out[(int64_t)m * N + n] = acc[i][j] + (bias ? bias[n] : 0.f);
    }
}

}  // namespace mpmae
