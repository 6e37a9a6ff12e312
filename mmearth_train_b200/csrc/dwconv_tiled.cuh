// Register-tiled depthwise 7x7 kernels (forward + LayerNorm, transposed stencil for dX, weight gradient).
//
// Same maths as dwconv.cuh (which stays as the generic fallback); what changes is the mapping:
//   * "patch" flavour (P = 8, 4, 2): one CTA per visible patch; the (P+6)^2 halo window is staged in shared
//     memory ([pixel][C], zeros for masked / out-of-image pixels); ONE THREAD = one channel x a strip of TR
//     output rows, its 49 taps live in registers and the inputs slide through a (P+6)-wide register row, so the
//     inner loop is pure FMA (TR*P*49 FMAs for (TR+6)*(P+6) shared loads) instead of 2 shared loads per FMA.
//   * "grid" flavour (P = 1, the 7x7 stage-3 / decoder maps): one CTA per sample, one thread per channel holding
//     the channel's whole 7x7 input AND its 49 taps in registers; masked cells are skipped by warp-uniform
//     branches (the mask is per sample).
// LayerNorm over C runs from a shared [pixels][C] tile (warp per pixel), then the tile -- contiguous in the
// Z-ordered row layout -- leaves with coalesced float4 stores.
//
// Replaces MinkowskiEngine/src/depthwise_convolution_kernel.cu:27-52 (fwd) and :69-122 (bwd).
#pragma once
#include "dwconv.cuh"

namespace mpmae {

__device__ __forceinline__ int zorder3(int py, int px) {
  auto spread = [](int b) { return (b & 1) | ((b & 2) << 1) | ((b & 4) << 2); };
  return spread(py) | (spread(px) << 1);
}

// LayerNorm (no affine) of `rows` rows staged in shared memory ([rows][C]) straight to global memory with the
// (row, part) mapping of common.cuh.  blockDim.x must be a multiple of 32.  Returns false when C is not of the form
// 4 * NP * F4 with F4 <= 6 (the caller then runs its generic warp-per-row path).
template <int F4>
__device__ __forceinline__ void ln_tile_to_global_f(const float *ubuf, int rows, int C, int np, float eps, float *out,
                                                    float *rstd_out) {
  const int total = rows * np;
  for (int base = 0; base < total; base += (int)blockDim.x) {
    const int item = base + (int)threadIdx.x;
    const bool ok = item < total;
    const int r = ok ? item / np : 0, part = ok ? item - r * np : 0;
    const float4 *ur = reinterpret_cast<const float4 *>(ubuf + (size_t)r * C) + part;
    float4 v[F4];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < F4; ++j) {
      v[j] = ok ? ur[j * np] : make_float4(0.f, 0.f, 0.f, 0.f);
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
    if (ok) {
      float4 *o = reinterpret_cast<float4 *>(out + (size_t)r * C) + part;
#pragma unroll
      for (int j = 0; j < F4; ++j) o[j * np] = make_float4(v[j].x * rstd, v[j].y * rstd, v[j].z * rstd, v[j].w * rstd);
      if (part == 0) rstd_out[r] = rstd;
    }
  }
}
__device__ __forceinline__ bool ln_tile_to_global(const float *ubuf, int rows, int C, float eps, float *out, float *rstd_out) {
  const int np = ln_parts(C), f4 = (C >> 2) / np;
  switch (f4) {
    case 1: ln_tile_to_global_f<1>(ubuf, rows, C, np, eps, out, rstd_out); return true;
    case 2: ln_tile_to_global_f<2>(ubuf, rows, C, np, eps, out, rstd_out); return true;
    case 3: ln_tile_to_global_f<3>(ubuf, rows, C, np, eps, out, rstd_out); return true;
    case 4: ln_tile_to_global_f<4>(ubuf, rows, C, np, eps, out, rstd_out); return true;
    case 5: ln_tile_to_global_f<5>(ubuf, rows, C, np, eps, out, rstd_out); return true;
    case 6: ln_tile_to_global_f<6>(ubuf, rows, C, np, eps, out, rstd_out); return true;
    default: return false;
  }
}

struct DwTiledArgs {
  DwArgs a;
  const int *vis_patch;  // [B*V] patch index of every slot (null in dense mode)
};

// stage the halo window of visible patch `pu` = n*V + slot:  win[(wy*W + wx)*C + c]
template <int P>
__device__ __forceinline__ void load_patch_window(const float *__restrict__ x, const int *nb, int n, int V, int C,
                                                  float *win, int nthreads) {
  constexpr int W = P + 6, NBR = (3 + P - 1) / P, NBW = 2 * NBR + 1;
  const int C4 = C >> 2;
  for (int i = threadIdx.x; i < W * W * C4; i += nthreads) {
    const int wp = i / C4, c4 = i - wp * C4;
    const int wy = wp / W, wx = wp - wy * W;
    const int ry = wy - 3 + NBR * P, rx = wx - 3 + NBR * P;  // >= 0
    const int by = ry / P, bx = rx / P;
    const int slot = nb[by * NBW + bx];
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (slot >= 0) {
      const int64_t row = ((int64_t)n * V + slot) * (P * P) + zorder3(ry - by * P, rx - bx * P);
      v = __ldg(reinterpret_cast<const float4 *>(x + row * C) + c4);
    }
    *reinterpret_cast<float4 *>(win + (size_t)wp * C + c4 * 4) = v;
  }
}

template <int P>
__device__ __forceinline__ void load_neighbour_slots(const int *slot_of, const Geo &g, int n, int l, int *nb) {
  constexpr int NBR = (3 + P - 1) / P, NBW = 2 * NBR + 1;
  if (threadIdx.x < NBW * NBW) {
    const int qy = l / g.G + (int)threadIdx.x / NBW - NBR, qx = l % g.G + (int)threadIdx.x % NBW - NBR;
    nb[threadIdx.x] = (qy >= 0 && qx >= 0 && qy < g.G && qx < g.G) ? slot_of[n * g.L + qy * g.G + qx] : -1;
  }
}

// ------------------------------------------------------------------------------------------------ patch flavour
// CT = channel count known at compile time (0 = run time): shared-memory addresses become immediates
template <int P, int TR, int CT>
__global__ void __launch_bounds__(512) dwconv_patch_kernel(DwTiledArgs t) { pdl_prologue();
  constexpr int W = P + 6, TILES = P / TR;
  const DwArgs &p = t.a;
  extern __shared__ __align__(16) float smem[];
  __shared__ int nb[25];
  const int C = CT ? CT : p.C, V = p.geo.V;
  float *win = smem;                         // [W*W][C]
  float *ubuf = win + (size_t)W * W * C;     // [P*P][C]
  const int pu = blockIdx.x, n = pu / V;
  const int l = t.vis_patch[pu];
  const int nthreads = blockDim.x;
  load_neighbour_slots<P>(p.slot_of, p.geo, n, l, nb);
  __syncthreads();
  load_patch_window<P>(p.x, nb, n, V, C, win, nthreads);

  const int item = threadIdx.x;
  const bool active = item < C * TILES;
  const int c = item % C, y0 = (item / C) * TR;
  float wreg[49];
  if (active) {
#pragma unroll
    for (int k = 0; k < 49; ++k) {
      const int kh = k / 7, kw = k % 7;
      const int a = p.flip ? 6 - kh : kh, b = p.flip ? 6 - kw : kw;
      wreg[k] = __ldg(p.w + a * p.w_skh + b * p.w_skw + c * p.w_sc);
    }
  }
  __syncthreads();
  if (active) {
    float acc[TR][P];
    const float b0 = p.bias ? __ldg(p.bias + c) : 0.f;
#pragma unroll
    for (int r = 0; r < TR; ++r)
#pragma unroll
      for (int ox = 0; ox < P; ++ox) acc[r][ox] = b0;
#pragma unroll
    for (int iy = 0; iy < TR + 6; ++iy) {
      float in[W];
#pragma unroll
      for (int j = 0; j < W; ++j) in[j] = win[(size_t)((y0 + iy) * W + j) * C + c];
#pragma unroll
      for (int r = 0; r < TR; ++r) {
        const int kh = iy - r;
        if (kh >= 0 && kh < 7) {
#pragma unroll
          for (int kw = 0; kw < 7; ++kw)
#pragma unroll
            for (int ox = 0; ox < P; ++ox) acc[r][ox] = fmaf(in[ox + kw], wreg[kh * 7 + kw], acc[r][ox]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < TR; ++r)
#pragma unroll
      for (int ox = 0; ox < P; ++ox) ubuf[(size_t)zorder3(y0 + r, ox) * C + c] = acc[r][ox];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = nthreads >> 5;
  const int64_t row0 = (int64_t)pu * (P * P);
  if (p.do_ln && !p.resid && ln_tile_to_global(ubuf, P * P, C, p.eps, p.out + row0 * C, p.rstd + row0)) return;
  if (p.do_ln) {
    for (int o = warp; o < P * P; o += nw) {
      float *ur = ubuf + (size_t)o * C;
      float s = 0.f;
      for (int cc = lane; cc < C; cc += 32) s += ur[cc];
      const float mean = warp_sum(s) / (float)C;
      float v = 0.f;
      for (int cc = lane; cc < C; cc += 32) { const float d = ur[cc] - mean; v += d * d; }
      const float rstd = rsqrtf(warp_sum(v) / (float)C + p.eps);
      for (int cc = lane; cc < C; cc += 32) ur[cc] = (ur[cc] - mean) * rstd;
      if (lane == 0) p.rstd[row0 + o] = rstd;
    }
    __syncthreads();
  }
  // the patch's rows are contiguous: coalesced float4 copy-out (+ residual)
  const int n4 = P * P * C / 4;
  float4 *dst = reinterpret_cast<float4 *>(p.out + row0 * C);
  const float4 *res = p.resid ? reinterpret_cast<const float4 *>(p.resid + row0 * C) : nullptr;
  for (int i = threadIdx.x; i < n4; i += nthreads) {
    float4 v = reinterpret_cast<const float4 *>(ubuf)[i];
    if (res) { const float4 r = __ldg(res + i); v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w; }
    dst[i] = v;
  }
}

// dW[tap, c] += sum du[o, c] * x[o + off(tap), c] ; db[c] += sum du[o, c].  Persistent CTAs: every thread keeps its
// channel's 49 partial sums in registers across all the patches it visits.
template <int P, int TR, int CT>
__global__ void __launch_bounds__(512) dwconv_patch_wgrad_kernel(DwWgradArgs p, const int *__restrict__ vis_patch) { pdl_prologue();
  constexpr int W = P + 6, TILES = P / TR;
  extern __shared__ __align__(16) float smem[];
  __shared__ int nb[25];
  const int C = CT ? CT : p.C, V = p.geo.V;
  float *win = smem;                         // [W*W][C]   x halo window
  float *dus = win + (size_t)W * W * C;      // [P*P][C]   du of the patch (Z-order) ; reused for the final reduction
  const int nthreads = blockDim.x;
  const int item = threadIdx.x;
  const bool active = item < C * TILES;
  const int c = item % C, y0 = (item / C) * TR;
  float dw[49];
  float db = 0.f;
#pragma unroll
  for (int k = 0; k < 49; ++k) dw[k] = 0.f;
  const int units = p.geo.B * V;
  for (int pu = blockIdx.x; pu < units; pu += gridDim.x) {
    const int n = pu / V, l = vis_patch[pu];
    __syncthreads();
    load_neighbour_slots<P>(p.slot_of, p.geo, n, l, nb);
    __syncthreads();
    load_patch_window<P>(p.x, nb, n, V, C, win, nthreads);
    {
      const float4 *src = reinterpret_cast<const float4 *>(p.du + (int64_t)pu * (P * P) * C);
      for (int i = threadIdx.x; i < P * P * C / 4; i += nthreads) reinterpret_cast<float4 *>(dus)[i] = __ldg(src + i);
    }
    __syncthreads();
    if (active) {
      float d[TR][P];
#pragma unroll
      for (int r = 0; r < TR; ++r)
#pragma unroll
        for (int ox = 0; ox < P; ++ox) { d[r][ox] = dus[(size_t)zorder3(y0 + r, ox) * C + c]; db += d[r][ox]; }
#pragma unroll
      for (int iy = 0; iy < TR + 6; ++iy) {
        float in[W];
#pragma unroll
        for (int j = 0; j < W; ++j) in[j] = win[(size_t)((y0 + iy) * W + j) * C + c];
#pragma unroll
        for (int r = 0; r < TR; ++r) {
          const int kh = iy - r;
          if (kh >= 0 && kh < 7) {
#pragma unroll
            for (int kw = 0; kw < 7; ++kw)
#pragma unroll
              for (int ox = 0; ox < P; ++ox) dw[kh * 7 + kw] = fmaf(d[r][ox], in[ox + kw], dw[kh * 7 + kw]);
          }
        }
      }
    }
  }
  // reduce the TILES strips of each channel in shared memory, then one atomic per (tap, channel) per CTA
  __syncthreads();
  float *red = smem;  // [50][C]
  for (int i = threadIdx.x; i < 50 * C; i += nthreads) red[i] = 0.f;
  __syncthreads();
  if (active) {
#pragma unroll
    for (int k = 0; k < 49; ++k) atomicAdd(&red[k * C + c], dw[k]);
    atomicAdd(&red[49 * C + c], db);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 49 * C; i += nthreads) {
    const int k = i / C, cc = i - k * C;
    const int kh = k / 7, kw = k - kh * 7;
    atomicAdd(&p.dw[kh * p.w_skh + kw * p.w_skw + cc * p.w_sc], red[i]);
  }
  if (p.dbias)
    for (int cc = threadIdx.x; cc < C; cc += nthreads) atomicAdd(&p.dbias[cc], red[49 * C + cc]);
}

// ------------------------------------------------------------------------------------------------ grid flavour (P = 1, G = 7)
__global__ void __launch_bounds__(512) dwconv_grid7_kernel(DwArgs p) { pdl_prologue();
  constexpr int G = 7, L = 49;
  extern __shared__ __align__(16) float smem[];   // ubuf [V][C]
  __shared__ int slot_s[L];
  const int C = p.C, V = p.geo.V, n = blockIdx.x;
  const int nthreads = blockDim.x;
  if (threadIdx.x < L) slot_s[threadIdx.x] = p.slot_of ? p.slot_of[n * L + threadIdx.x] : (int)threadIdx.x;
  __syncthreads();
  const int64_t row0 = (int64_t)n * V;
  for (int c = threadIdx.x; c < C; c += nthreads) {
    float x[L], w[L];
#pragma unroll
    for (int i = 0; i < L; ++i) {
      const int s = slot_s[i];
      x[i] = s >= 0 ? __ldg(p.x + (row0 + s) * C + c) : 0.f;
      const int kh = i / 7, kw = i % 7;
      const int a = p.flip ? 6 - kh : kh, b = p.flip ? 6 - kw : kw;
      w[i] = __ldg(p.w + a * p.w_skh + b * p.w_skw + c * p.w_sc);
    }
    const float b0 = p.bias ? __ldg(p.bias + c) : 0.f;
#pragma unroll
    for (int oy = 0; oy < G; ++oy)
#pragma unroll
      for (int ox = 0; ox < G; ++ox) {
        const int s = slot_s[oy * G + ox];
        if (s >= 0) {  // warp-uniform: the mask is per sample
          float acc = b0;
#pragma unroll
          for (int kh = 0; kh < 7; ++kh)
#pragma unroll
            for (int kw = 0; kw < 7; ++kw) {
              const int iy = oy + kh - 3, ix = ox + kw - 3;
              if (iy >= 0 && iy < G && ix >= 0 && ix < G) acc = fmaf(x[iy * G + ix], w[kh * 7 + kw], acc);
            }
          smem[(size_t)s * C + c] = acc;
        }
      }
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = nthreads >> 5;
  if (p.do_ln && !p.resid && ln_tile_to_global(smem, V, C, p.eps, p.out + row0 * C, p.rstd + row0)) return;
  if (p.do_ln) {
    for (int o = warp; o < V; o += nw) {
      float *ur = smem + (size_t)o * C;
      float s = 0.f;
      for (int cc = lane; cc < C; cc += 32) s += ur[cc];
      const float mean = warp_sum(s) / (float)C;
      float v = 0.f;
      for (int cc = lane; cc < C; cc += 32) { const float d = ur[cc] - mean; v += d * d; }
      const float rstd = rsqrtf(warp_sum(v) / (float)C + p.eps);
      for (int cc = lane; cc < C; cc += 32) ur[cc] = (ur[cc] - mean) * rstd;
      if (lane == 0) p.rstd[row0 + o] = rstd;
    }
    __syncthreads();
  }
  const int n4 = V * C / 4;
  float4 *dst = reinterpret_cast<float4 *>(p.out + row0 * C);
  const float4 *res = p.resid ? reinterpret_cast<const float4 *>(p.resid + row0 * C) : nullptr;
  for (int i = threadIdx.x; i < n4; i += nthreads) {
    float4 v = reinterpret_cast<const float4 *>(smem)[i];
    if (res) { const float4 r = __ldg(res + i); v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w; }
    dst[i] = v;
  }
}

// weight gradient on the 7x7 maps: thread per channel, x[49] and dW[49] in registers, persistent over samples
__global__ void __launch_bounds__(128, 2) dwconv_grid7_wgrad_kernel(DwWgradArgs p) { pdl_prologue();
  constexpr int G = 7, L = 49;
  __shared__ int slot_s[L];
  const int C = p.C, V = p.geo.V;
  const int c = blockIdx.y * blockDim.x + threadIdx.x;
  const bool active = c < C;
  float dw[L];
  float db = 0.f;
#pragma unroll
  for (int i = 0; i < L; ++i) dw[i] = 0.f;
  for (int n = blockIdx.x; n < p.geo.B; n += gridDim.x) {
    __syncthreads();
    if (threadIdx.x < L) slot_s[threadIdx.x] = p.slot_of ? p.slot_of[n * L + threadIdx.x] : (int)threadIdx.x;
    __syncthreads();
    if (!active) continue;
    const int64_t row0 = (int64_t)n * V;
    float x[L], dreg[L];   // all 98 loads are issued before the first FMA needs one
#pragma unroll
    for (int i = 0; i < L; ++i) {
      const int s = slot_s[i];
      x[i] = s >= 0 ? __ldg(p.x + (row0 + s) * C + c) : 0.f;
      dreg[i] = s >= 0 ? __ldg(p.du + (row0 + s) * C + c) : 0.f;
    }
#pragma unroll
    for (int oy = 0; oy < G; ++oy)
#pragma unroll
      for (int ox = 0; ox < G; ++ox) {
        const int s = slot_s[oy * G + ox];
        if (s >= 0) {
          const float d = dreg[oy * G + ox];
          db += d;
#pragma unroll
          for (int kh = 0; kh < 7; ++kh)
#pragma unroll
            for (int kw = 0; kw < 7; ++kw) {
              const int iy = oy + kh - 3, ix = ox + kw - 3;
              if (iy >= 0 && iy < G && ix >= 0 && ix < G) dw[kh * 7 + kw] = fmaf(d, x[iy * G + ix], dw[kh * 7 + kw]);
            }
        }
      }
  }
  if (active) {
#pragma unroll
    for (int k = 0; k < L; ++k) {
      const int kh = k / 7, kw = k % 7;
      atomicAdd(&p.dw[kh * p.w_skh + kw * p.w_skw + c * p.w_sc], dw[k]);
    }
    if (p.dbias) atomicAdd(&p.dbias[c], db);
  }
}

// ------------------------------------------------------------------------------------------------ sample flavour (P = 2, G = 7)
// With 2x2-pixel patches a per-patch halo window is 16x larger than the patch, so stage 2 stages the WHOLE sample
// instead: a zero-padded 20x20 pixel grid per channel chunk of CC channels lives in shared memory, every visible pixel
// is loaded exactly once, and one thread computes a 2x2 patch of one channel from an 8x8 register window
// (64 shared loads for 196 FMAs).  The [76, C] result tile collects all chunks, then LayerNorm + coalesced copy-out.
template <int CC>
__global__ void __launch_bounds__(512) dwconv_s2_kernel(DwArgs p, const int *__restrict__ vis_patch) { pdl_prologue();
  constexpr int GW = 20, NCELL = GW * GW, CC4 = CC / 4;
  constexpr int NT = (512 / CC) * CC, NPG = NT / CC;
  extern __shared__ __align__(16) float smem[];
  __shared__ int vis[64];
  const int C = p.C, V = p.geo.V, n = blockIdx.x, tid = threadIdx.x;
  float *grid = smem;                       // [NCELL][CC]
  float *ubuf = smem + NCELL * CC;          // [V*4][C]
  const int64_t row0 = (int64_t)n * V * 4;
  if (tid < V) vis[tid] = vis_patch[n * V + tid];
  const int c = tid % CC, pg = tid / CC;
  for (int c0 = 0; c0 < C; c0 += CC) {
    __syncthreads();
    for (int i = tid; i < NCELL * CC4; i += NT) reinterpret_cast<float4 *>(grid)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    for (int i = tid; i < V * 4 * CC4; i += NT) {
      const int r = i / CC4, c4 = i - r * CC4;
      const int l = vis[r >> 2], m = r & 3;
      const int gy = (l / 7) * 2 + (m & 1) + 3, gx = (l % 7) * 2 + (m >> 1) + 3;
      reinterpret_cast<float4 *>(grid)[(gy * GW + gx) * CC4 + c4] =
          __ldg(reinterpret_cast<const float4 *>(p.x + (row0 + r) * C + c0) + c4);
    }
    float w[49];
#pragma unroll
    for (int k = 0; k < 49; ++k) {
      const int kh = k / 7, kw = k % 7;
      const int a = p.flip ? 6 - kh : kh, b = p.flip ? 6 - kw : kw;
      w[k] = __ldg(p.w + a * p.w_skh + b * p.w_skw + (c0 + c) * p.w_sc);
    }
    const float b0 = p.bias ? __ldg(p.bias + c0 + c) : 0.f;
    __syncthreads();
    if (tid < NT) {
      for (int slot = pg; slot < V; slot += NPG) {
        const int l = vis[slot];
        const int y0 = (l / 7) * 2, x0 = (l % 7) * 2;   // window origin in the padded grid
        float acc[2][2] = {{b0, b0}, {b0, b0}};
#pragma unroll
        for (int iy = 0; iy < 8; ++iy) {
          float in[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) in[j] = grid[((y0 + iy) * GW + x0 + j) * CC + c];
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            const int kh = iy - r;
            if (kh >= 0 && kh < 7) {
#pragma unroll
              for (int kw = 0; kw < 7; ++kw) {
                acc[r][0] = fmaf(in[kw], w[kh * 7 + kw], acc[r][0]);
                acc[r][1] = fmaf(in[kw + 1], w[kh * 7 + kw], acc[r][1]);
              }
            }
          }
        }
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
          for (int ox = 0; ox < 2; ++ox) ubuf[(size_t)(slot * 4 + (r | (ox << 1))) * C + c0 + c] = acc[r][ox];
      }
    }
  }
  __syncthreads();
  const int lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  const int rows = V * 4;
  if (p.do_ln && !p.resid && ln_tile_to_global(ubuf, rows, C, p.eps, p.out + row0 * C, p.rstd + row0)) return;
  if (p.do_ln) {
    for (int o = warp; o < rows; o += nw) {
      float *ur = ubuf + (size_t)o * C;
      float s = 0.f;
      for (int cc = lane; cc < C; cc += 32) s += ur[cc];
      const float mean = warp_sum(s) / (float)C;
      float v = 0.f;
      for (int cc = lane; cc < C; cc += 32) { const float d = ur[cc] - mean; v += d * d; }
      const float rstd = rsqrtf(warp_sum(v) / (float)C + p.eps);
      for (int cc = lane; cc < C; cc += 32) ur[cc] = (ur[cc] - mean) * rstd;
      if (lane == 0) p.rstd[row0 + o] = rstd;
    }
    __syncthreads();
  }
  const int n4 = rows * C / 4;
  float4 *dst = reinterpret_cast<float4 *>(p.out + row0 * C);
  const float4 *res = p.resid ? reinterpret_cast<const float4 *>(p.resid + row0 * C) : nullptr;
  for (int i = tid; i < n4; i += blockDim.x) {
    float4 v = reinterpret_cast<const float4 *>(ubuf)[i];
    if (res) { const float4 r = __ldg(res + i); v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w; }
    dst[i] = v;
  }
}

// weight gradient, same staging; blockIdx.y = channel chunk, blockIdx.x strides samples; every thread keeps its channel's
// 49 partial sums in registers across all the patches and samples it visits
template <int CC>
__global__ void __launch_bounds__(512) dwconv_s2_wgrad_kernel(DwWgradArgs p, const int *__restrict__ vis_patch) { pdl_prologue();
  constexpr int GW = 20, NCELL = GW * GW, CC4 = CC / 4;
  constexpr int NT = (512 / CC) * CC, NPG = NT / CC;
  extern __shared__ __align__(16) float smem[];
  __shared__ int vis[64];
  const int C = p.C, V = p.geo.V, tid = threadIdx.x;
  const int c0 = blockIdx.y * CC;
  float *grid = smem;
  const int c = tid % CC, pg = tid / CC;
  float dw[49];
  float db = 0.f;
#pragma unroll
  for (int k = 0; k < 49; ++k) dw[k] = 0.f;
  for (int n = blockIdx.x; n < p.geo.B; n += gridDim.x) {
    const int64_t row0 = (int64_t)n * V * 4;
    __syncthreads();
    if (tid < V) vis[tid] = vis_patch[n * V + tid];
    for (int i = tid; i < NCELL * CC4; i += NT) reinterpret_cast<float4 *>(grid)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    for (int i = tid; i < V * 4 * CC4; i += NT) {
      const int r = i / CC4, c4 = i - r * CC4;
      const int l = vis[r >> 2], m = r & 3;
      const int gy = (l / 7) * 2 + (m & 1) + 3, gx = (l % 7) * 2 + (m >> 1) + 3;
      reinterpret_cast<float4 *>(grid)[(gy * GW + gx) * CC4 + c4] =
          __ldg(reinterpret_cast<const float4 *>(p.x + (row0 + r) * C + c0) + c4);
    }
    __syncthreads();
    if (tid < NT) {
      for (int slot = pg; slot < V; slot += NPG) {
        const int l = vis[slot];
        const int y0 = (l / 7) * 2, x0 = (l % 7) * 2;
        float d[2][2];
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
          for (int ox = 0; ox < 2; ++ox) {
            d[r][ox] = __ldg(p.du + (row0 + slot * 4 + (r | (ox << 1))) * C + c0 + c);
            db += d[r][ox];
          }
#pragma unroll
        for (int iy = 0; iy < 8; ++iy) {
          float in[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) in[j] = grid[((y0 + iy) * GW + x0 + j) * CC + c];
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            const int kh = iy - r;
            if (kh >= 0 && kh < 7) {
#pragma unroll
              for (int kw = 0; kw < 7; ++kw)
                dw[kh * 7 + kw] = fmaf(d[r][0], in[kw], fmaf(d[r][1], in[kw + 1], dw[kh * 7 + kw]));
            }
          }
        }
      }
    }
  }
  __syncthreads();
  float *red = smem;  // [50][CC]
  for (int i = tid; i < 50 * CC; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
  if (tid < NT) {
#pragma unroll
    for (int k = 0; k < 49; ++k) atomicAdd(&red[k * CC + c], dw[k]);
    atomicAdd(&red[49 * CC + c], db);
  }
  __syncthreads();
  for (int i = tid; i < 49 * CC; i += blockDim.x) {
    const int k = i / CC, cc = i - k * CC;
    const int kh = k / 7, kw = k - kh * 7;
    atomicAdd(&p.dw[kh * p.w_skh + kw * p.w_skw + (c0 + cc) * p.w_sc], red[i]);
  }
  if (p.dbias)
    for (int cc = tid; cc < CC; cc += blockDim.x) atomicAdd(&p.dbias[c0 + cc], red[49 * CC + cc]);
}

template <int CC>
inline cudaError_t launch_s2(const DwArgs &a, const int *vis_patch, cudaStream_t st) {
  const size_t sm = ((size_t)400 * CC + (size_t)a.geo.V * 4 * a.C) * sizeof(float);
  if (sm > 226 * 1024 || a.geo.V > 49) return cudaErrorInvalidConfiguration;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(dwconv_s2_kernel<CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  pdl(dwconv_s2_kernel<CC>, a.geo.B, (512 / CC) * CC, sm, st)(a, vis_patch);
  return cudaGetLastError();
}
template <int CC>
inline cudaError_t launch_s2_wgrad(const DwWgradArgs &p, const int *vis_patch, cudaStream_t st) {
  const size_t sm = (size_t)400 * CC * sizeof(float);
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(dwconv_s2_wgrad_kernel<CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const int chunks = p.C / CC;
  int per_sm = (int)((220 * 1024) / sm);
  if (per_sm > 2) per_sm = 2;
  if (per_sm < 1) per_sm = 1;
  int gx = (148 * per_sm) / chunks;
  if (gx < 1) gx = 1;
  if (gx > p.geo.B) gx = p.geo.B;
  pdl(dwconv_s2_wgrad_kernel<CC>, dim3(gx, chunks), (512 / CC) * CC, sm, st)(p, vis_patch);
  return cudaGetLastError();
}
inline int s2_chunk(int C) {
  if (C % 80 == 0) return 80;
  if (C % 64 == 0) return 64;
  if (C % 48 == 0) return 48;
  return 0;
}

// ------------------------------------------------------------------------------------------------ launchers
template <int P, int TR, int CT>
inline cudaError_t launch_patch(const DwTiledArgs &t, cudaStream_t st) {
  const int C = t.a.C;
  const int threads = ((C * (P / TR) + 31) / 32) * 32;
  const size_t sm = ((size_t)(P + 6) * (P + 6) + P * P) * C * sizeof(float);
  if (threads > 512 || sm > 226 * 1024) return cudaErrorInvalidConfiguration;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(dwconv_patch_kernel<P, TR, CT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
    if (e != cudaSuccess) return e;
    (void)cudaFuncSetAttribute(dwconv_patch_kernel<P, TR, CT>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    configured = true;
  }
  pdl(dwconv_patch_kernel<P, TR, CT>, t.a.geo.B * t.a.geo.V, threads, sm, st)(t);
  return cudaGetLastError();
}

template <int P, int TR, int CT>
inline cudaError_t launch_patch_wgrad(const DwWgradArgs &p, const int *vis_patch, cudaStream_t st) {
  const int C = p.C;
  const int threads = ((C * (P / TR) + 31) / 32) * 32;
  size_t sm = ((size_t)(P + 6) * (P + 6) + P * P) * C * sizeof(float);
  if (sm < (size_t)50 * C * sizeof(float)) sm = (size_t)50 * C * sizeof(float);
  if (threads > 512 || sm > 226 * 1024) return cudaErrorInvalidConfiguration;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(dwconv_patch_wgrad_kernel<P, TR, CT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
    if (e != cudaSuccess) return e;
    (void)cudaFuncSetAttribute(dwconv_patch_wgrad_kernel<P, TR, CT>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    configured = true;
  }
  int per_sm = (int)((220 * 1024) / sm);
  if (per_sm > 2048 / threads) per_sm = 2048 / threads;
  if (per_sm < 1) per_sm = 1;
  int grid = 148 * per_sm;
  const int units = p.geo.B * p.geo.V;
  if (grid > units) grid = units;
  pdl(dwconv_patch_wgrad_kernel<P, TR, CT>, grid, threads, sm, st)(p, vis_patch);
  return cudaGetLastError();
}

// Dispatch: returns cudaErrorInvalidConfiguration when the tiled kernels do not take the shape (caller falls back).
inline cudaError_t launch_dwconv_tiled(const DwArgs &a, const int *vis_patch, cudaStream_t st) {
  if (a.C % 4 != 0) return cudaErrorInvalidConfiguration;
  if (a.P == 1) {
    if (a.geo.G != 7) return cudaErrorInvalidConfiguration;
    const size_t sm = (size_t)a.geo.V * a.C * sizeof(float);
    if (sm > 226 * 1024) return cudaErrorInvalidConfiguration;
    static bool configured = false;
    if (!configured) {
      cudaError_t e = cudaFuncSetAttribute(dwconv_grid7_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
      if (e != cudaSuccess) return e;
      configured = true;
    }
    int threads = ((a.C + 31) / 32) * 32;
    if (threads > 512) threads = 512;
    pdl(dwconv_grid7_kernel, a.geo.B, threads, sm, st)(a);
    return cudaGetLastError();
  }
  if (!vis_patch || !a.slot_of) return cudaErrorInvalidConfiguration;
  DwTiledArgs t{a, vis_patch};
  if (a.P == 8) {
    if (a.C == 40) return launch_patch<8, 2, 40>(t, st);
    if (a.C == 96) return launch_patch<8, 2, 96>(t, st);
    return launch_patch<8, 2, 0>(t, st);
  }
  if (a.P == 4) {
    if (a.C == 80) return launch_patch<4, 2, 80>(t, st);
    if (a.C == 192) return launch_patch<4, 2, 192>(t, st);
    return launch_patch<4, 2, 0>(t, st);
  }
  if (a.P == 2) {
    if (a.geo.G == 7) {
      cudaError_t e = cudaErrorInvalidConfiguration;
      const int cc = s2_chunk(a.C);
      if (cc == 80) e = launch_s2<80>(a, vis_patch, st);
      else if (cc == 64) e = launch_s2<64>(a, vis_patch, st);
      else if (cc == 48) e = launch_s2<48>(a, vis_patch, st);
      if (e != cudaErrorInvalidConfiguration) return e;
      (void)cudaGetLastError();
    }
    return launch_patch<2, 2, 0>(t, st);
  }
  return cudaErrorInvalidConfiguration;
}

inline cudaError_t launch_dwconv_wgrad_tiled(const DwWgradArgs &p, const int *vis_patch, cudaStream_t st) {
  if (p.C % 4 != 0) return cudaErrorInvalidConfiguration;
  if (p.P == 1) {
    if (p.geo.G != 7) return cudaErrorInvalidConfiguration;
    const int threads = p.C >= 128 ? 128 : ((p.C + 31) / 32) * 32;
    const int cy = (p.C + threads - 1) / threads;
    int gx = (148 * 4) / cy;
    if (gx < 1) gx = 1;
    if (gx > p.geo.B) gx = p.geo.B;
    pdl(dwconv_grid7_wgrad_kernel, dim3(gx, cy), threads, 0, st)(p);
    return cudaGetLastError();
  }
  if (!vis_patch || !p.slot_of) return cudaErrorInvalidConfiguration;
  if (p.P == 8) {
    if (p.C == 40) return launch_patch_wgrad<8, 2, 40>(p, vis_patch, st);
    if (p.C == 96) return launch_patch_wgrad<8, 2, 96>(p, vis_patch, st);
    return launch_patch_wgrad<8, 2, 0>(p, vis_patch, st);
  }
  if (p.P == 4) {
    if (p.C == 80) return launch_patch_wgrad<4, 2, 80>(p, vis_patch, st);
    if (p.C == 192) return launch_patch_wgrad<4, 2, 192>(p, vis_patch, st);
    return launch_patch_wgrad<4, 2, 0>(p, vis_patch, st);
  }
  if (p.P == 2) {
    if (p.geo.G == 7 && p.geo.V <= 49) {
      const int cc = s2_chunk(p.C);
      if (cc == 80) return launch_s2_wgrad<80>(p, vis_patch, st);
      if (cc == 64) return launch_s2_wgrad<64>(p, vis_patch, st);
      if (cc == 48) return launch_s2_wgrad<48>(p, vis_patch, st);
    }
    return launch_patch_wgrad<2, 2, 0>(p, vis_patch, st);
  }
  return cudaErrorInvalidConfiguration;
}

}  // namespace mpmae
