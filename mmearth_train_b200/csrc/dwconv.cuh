// Depthwise 7x7 convolution over VISIBLE patches only (+ bias + LayerNorm), its transposed stencil
// for dX, and the weight/bias gradient.
//
// Replaces MinkowskiEngine's per-kernel-offset launches with atomics
//   forward  detail::matmulDwconv   MinkowskiEngine/src/depthwise_convolution_kernel.cu:27-52
//   backward detail::matmulDwconv2  MinkowskiEngine/src/depthwise_convolution_kernel.cu:69-122
// and the coordinate hash map / kernel map (src/coordinate_map_gpu.cu:1479-1547): the mask is
// patch aligned, so a neighbour's row is a closed-form function of the per-sample slot table.
//
// A CTA owns one visible patch (P >= 2: "patch mode", the (P+6)^2 halo window is staged in shared
// memory with zeros for masked / out-of-image pixels) or one whole sample (P == 1: "grid mode",
// the GxG cell grid is staged).  One warp per output pixel, lanes over channels, so the LayerNorm
// reduction is a warp shuffle.
#pragma once
#include "common.cuh"

namespace mpmae {

struct DwArgs {
  const float *x;      // [R, C]
  const float *w;      // tap (kh, kw), channel c at  kh*w_skh + kw*w_skw + c*w_sc
  int w_skh, w_skw, w_sc;
  const float *bias;   // [C] or null
  const float *resid;  // [R, C] or null, added to the result
  float *out;          // [R, C]
  float *rstd;         // [R] (do_ln)
  const int *slot_of;  // [B*L] or null (all cells visible, V == L)
  Geo geo;
  int P, C;
  int flip;            // use tap (6-kh, 6-kw): transposed stencil
  int do_ln;           // write (u - mean) * rstd and rstd
  float eps;
  int w_in_smem;
  float *colsum_out;   // optional [C]: += column sums of the result rows (the bias gradient of whatever consumes them);
                       // honoured by the pipelined kernels, the dispatcher runs colsum_kernel after the others
};

// row of stage-grid pixel (gy, gx) of sample n, or -1
__device__ __forceinline__ int64_t sparse_row(const int *slot_of, const Geo &g, int n, int P, int gy, int gx) {
  const int side = g.G * P;
  if (gy < 0 || gx < 0 || gy >= side || gx >= side) return -1;
  const int qy = gy / P, qx = gx / P;
  const int l = qy * g.G + qx;
  const int slot = slot_of ? slot_of[n * g.L + l] : l;
  if (slot < 0) return -1;
  return ((int64_t)n * g.V + slot) * (P * P) + morton_encode(gy - qy * P, gx - qx * P);
}

template <int NW>
__global__ void __launch_bounds__(NW * 32) dwconv_fwd_kernel(DwArgs p) { pdl_prologue();
  extern __shared__ __align__(16) float smem[];
  const int C = p.C, P = p.P, G = p.geo.G;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool grid_mode = (P == 1);
  const int Wd = grid_mode ? G : P + 6;
  const int npix = Wd * Wd;
  float *win = smem;
  float *urow = win + (size_t)npix * C;
  float *wsm = urow + NW * C;

  int n, ph = 0, pw = 0;
  int64_t out_base;
  if (grid_mode) {
    n = blockIdx.x;
    out_base = (int64_t)n * p.geo.V;
  } else {
    n = blockIdx.x / p.geo.V;
    const int slot = blockIdx.x % p.geo.V;
    // patch index of this slot: scan the slot table (L <= a few hundred, warp-uniform)
    int l = slot;
    if (p.slot_of) {
      l = -1;
      for (int q = 0; q < p.geo.L; ++q)
        if (p.slot_of[n * p.geo.L + q] == slot) { l = q; break; }
    }
    ph = l / G; pw = l % G;
    out_base = (int64_t)blockIdx.x * P * P;
  }

  const int C4 = C >> 2;
  for (int i = tid; i < npix * C4; i += NW * 32) {
    const int wp = i / C4, c4 = i - wp * C4;
    const int wy = wp / Wd, wx = wp - wy * Wd;
    const int gy = grid_mode ? wy : ph * P + wy - 3;
    const int gx = grid_mode ? wx : pw * P + wx - 3;
    const int64_t r = sparse_row(p.slot_of, p.geo, n, P, gy, gx);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r >= 0) v = *reinterpret_cast<const float4 *>(p.x + r * C + c4 * 4);
    *reinterpret_cast<float4 *>(win + (size_t)wp * C + c4 * 4) = v;
  }
  if (p.w_in_smem) {
    for (int i = tid; i < 49 * C; i += NW * 32) {
      const int t = i / C, c = i - t * C;
      int kh = t / 7, kw = t - kh * 7;
      if (p.flip) { kh = 6 - kh; kw = 6 - kw; }
      wsm[i] = p.w[kh * p.w_skh + kw * p.w_skw + c * p.w_sc];
    }
  }
  __syncthreads();

  const int n_out = grid_mode ? p.geo.L : P * P;
  float *my_u = urow + warp * C;
  for (int o = warp; o < n_out; o += NW) {
    int cy, cx;
    int64_t orow;
    if (grid_mode) {
      const int slot = p.slot_of ? p.slot_of[n * p.geo.L + o] : o;
      if (slot < 0) continue;
      cy = o / G; cx = o - cy * G;
      orow = out_base + slot;
    } else {
      int py, px;
      morton_decode(o, py, px);
      cy = py + 3; cx = px + 3;
      orow = out_base + o;
    }
    float lsum = 0.f;
    for (int c = lane; c < C; c += 32) {
      float acc = p.bias ? p.bias[c] : 0.f;
#pragma unroll
      for (int kh = 0; kh < 7; ++kh) {
        const int wy = cy + kh - 3;
        if (grid_mode && (wy < 0 || wy >= G)) continue;
#pragma unroll
        for (int kw = 0; kw < 7; ++kw) {
          const int wx = cx + kw - 3;
          if (grid_mode && (wx < 0 || wx >= G)) continue;
          float wv;
          if (p.w_in_smem) {
            wv = wsm[(kh * 7 + kw) * C + c];
          } else {
            const int a = p.flip ? 6 - kh : kh, b = p.flip ? 6 - kw : kw;
            wv = __ldg(p.w + a * p.w_skh + b * p.w_skw + c * p.w_sc);
          }
          acc = fmaf(win[(size_t)(wy * Wd + wx) * C + c], wv, acc);
        }
      }
      if (p.resid) acc += p.resid[orow * C + c];
      if (p.do_ln) { my_u[c] = acc; lsum += acc; }
      else p.out[orow * C + c] = acc;
    }
    if (p.do_ln) {
      const float mean = warp_sum(lsum) / (float)C;
      float lvar = 0.f;
      for (int c = lane; c < C; c += 32) { const float d = my_u[c] - mean; lvar += d * d; }
      const float rstd = rsqrtf(warp_sum(lvar) / (float)C + p.eps);
      for (int c = lane; c < C; c += 32) p.out[orow * C + c] = (my_u[c] - mean) * rstd;
      if (lane == 0) p.rstd[orow] = rstd;
      __syncwarp();
    }
  }
}

inline size_t dwconv_fwd_smem(int P, int G, int C, int NW, bool w_in_smem) {
  const int Wd = (P == 1) ? G : P + 6;
  return ((size_t)Wd * Wd * C + (size_t)NW * C + (w_in_smem ? 49 * (size_t)C : 0)) * sizeof(float);
}

inline cudaError_t launch_dwconv_fwd(DwArgs p, cudaStream_t st) {
  constexpr int NW = 8;
  size_t sm = dwconv_fwd_smem(p.P, p.geo.G, p.C, NW, true);
  p.w_in_smem = 1;
  if (sm > 200 * 1024) { p.w_in_smem = 0; sm = dwconv_fwd_smem(p.P, p.geo.G, p.C, NW, false); }
  if (sm > 227 * 1024) return cudaErrorInvalidValue;
  static size_t configured = 0;
  if (sm > configured) {
    cudaError_t e = cudaFuncSetAttribute(dwconv_fwd_kernel<NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024));
    if (e != cudaSuccess) return e;
    configured = 227 * 1024;
  }
  const unsigned grid = (p.P == 1) ? p.geo.B : p.geo.B * p.geo.V;
  pdl(dwconv_fwd_kernel<NW>, grid, NW * 32, sm, st)(p);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// dW[tap, c] += sum_out du[out, c] * x[out + off(tap), c] ; db[c] += sum_out du[out, c]
struct DwWgradArgs {
  const float *x;     // [R, C] forward input of the depthwise conv
  const float *du;    // [R, C] gradient at its output
  float *dw;          // parameter-layout gradient, same strides as the forward weight
  int w_skh, w_skw, w_sc;
  float *dbias;       // [C]
  const int *slot_of;
  Geo geo;
  int P, C, CC;       // CC = channel chunk per CTA (blockIdx.y)
};

template <int NW>
__global__ void __launch_bounds__(NW * 32) dwconv_wgrad_kernel(DwWgradArgs p) { pdl_prologue();
  extern __shared__ __align__(16) float smem[];
  const int C = p.C, CC = p.CC, P = p.P, G = p.geo.G;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool grid_mode = (P == 1);
  const int Wd = grid_mode ? G : P + 6;
  const int npix = Wd * Wd;
  const int n_out = grid_mode ? p.geo.L : P * P;
  const int c0 = blockIdx.y * CC;
  float *xwin = smem;                          // [npix][CC]
  float *dus = xwin + (size_t)npix * CC;       // [n_out][CC]
  float *dws = dus + (size_t)n_out * CC;       // [50][CC]  (tap 49 = bias)
  for (int i = tid; i < 50 * CC; i += NW * 32) dws[i] = 0.f;

  const int units = grid_mode ? p.geo.B : p.geo.B * p.geo.V;
  const int CC4 = CC >> 2;
  for (int u = blockIdx.x; u < units; u += gridDim.x) {
    int n, ph = 0, pw = 0;
    int64_t out_base;
    if (grid_mode) {
      n = u; out_base = (int64_t)n * p.geo.V;
    } else {
      n = u / p.geo.V;
      const int slot = u % p.geo.V;
      int l = slot;
      if (p.slot_of) {
        l = -1;
        for (int q = 0; q < p.geo.L; ++q)
          if (p.slot_of[n * p.geo.L + q] == slot) { l = q; break; }
      }
      ph = l / G; pw = l % G;
      out_base = (int64_t)u * P * P;
    }
    __syncthreads();  // previous unit's readers are done
    for (int i = tid; i < npix * CC4; i += NW * 32) {
      const int wp = i / CC4, c4 = i - wp * CC4;
      const int wy = wp / Wd, wx = wp - wy * Wd;
      const int gy = grid_mode ? wy : ph * P + wy - 3;
      const int gx = grid_mode ? wx : pw * P + wx - 3;
      const int64_t r = sparse_row(p.slot_of, p.geo, n, P, gy, gx);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r >= 0) v = *reinterpret_cast<const float4 *>(p.x + r * C + c0 + c4 * 4);
      *reinterpret_cast<float4 *>(xwin + (size_t)wp * CC + c4 * 4) = v;
    }
    for (int i = tid; i < n_out * CC4; i += NW * 32) {
      const int o = i / CC4, c4 = i - o * CC4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (grid_mode) {
        const int slot = p.slot_of ? p.slot_of[n * p.geo.L + o] : o;
        if (slot >= 0) v = *reinterpret_cast<const float4 *>(p.du + (out_base + slot) * C + c0 + c4 * 4);
      } else {
        v = *reinterpret_cast<const float4 *>(p.du + (out_base + o) * C + c0 + c4 * 4);
      }
      *reinterpret_cast<float4 *>(dus + (size_t)o * CC + c4 * 4) = v;
    }
    __syncthreads();
    for (int t = warp; t < 50; t += NW) {
      const int kh = t / 7, kw = t - kh * 7;
      for (int c = lane; c < CC; c += 32) {
        float acc = 0.f;
        for (int o = 0; o < n_out; ++o) {
          const float d = dus[(size_t)o * CC + c];
          if (t == 49) { acc += d; continue; }
          int cy, cx;
          if (grid_mode) { cy = o / G; cx = o - cy * G; }
          else { int py, px; morton_decode(o, py, px); cy = py + 3; cx = px + 3; }
          const int wy = cy + kh - 3, wx = cx + kw - 3;
          if (grid_mode && (wy < 0 || wy >= G || wx < 0 || wx >= G)) continue;
          acc = fmaf(d, xwin[(size_t)(wy * Wd + wx) * CC + c], acc);
        }
        dws[t * CC + c] += acc;
      }
    }
  }
  __syncthreads();
  for (int i = tid; i < 49 * CC; i += NW * 32) {
    const int t = i / CC, c = i - t * CC;
    const int kh = t / 7, kw = t - kh * 7;
    atomicAdd(&p.dw[kh * p.w_skh + kw * p.w_skw + (c0 + c) * p.w_sc], dws[i]);
  }
  if (p.dbias)
    for (int c = tid; c < CC; c += NW * 32) atomicAdd(&p.dbias[c0 + c], dws[49 * CC + c]);
}

inline int dwconv_pick_chunk(int C) {
  const int cands[] = {128, 96, 80, 64, 48, 40, 32, 16, 8, 4};
  if (C <= 128) return C;
  for (int c : cands)
    if (C % c == 0) return c;
  return 4;
}

inline cudaError_t launch_dwconv_wgrad(DwWgradArgs p, cudaStream_t st) {
  constexpr int NW = 8;
  p.CC = dwconv_pick_chunk(p.C);
  const int Wd = (p.P == 1) ? p.geo.G : p.P + 6;
  const int n_out = (p.P == 1) ? p.geo.L : p.P * p.P;
  const size_t sm = ((size_t)Wd * Wd + n_out + 50) * p.CC * sizeof(float);
  if (sm > 227 * 1024) return cudaErrorInvalidValue;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(dwconv_wgrad_kernel<NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024));
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const int units = (p.P == 1) ? p.geo.B : p.geo.B * p.geo.V;
  const int chunks = p.C / p.CC;
  int gx = (148 * 4) / chunks;
  if (gx < 1) gx = 1;
  if (gx > units) gx = units;
  pdl(dwconv_wgrad_kernel<NW>, dim3(gx, chunks), NW * 32, sm, st)(p);
  return cudaGetLastError();
}

}  // namespace mpmae
