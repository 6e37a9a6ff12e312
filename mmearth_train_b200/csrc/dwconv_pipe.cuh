// Persistent, TMA-pipelined depthwise 7x7 kernels for the patch stages (P = 8, 4).
//
// Same maths and register tiling as dwconv_tiled.cuh (one thread = one channel x a strip of TR output rows, 49 taps in
// registers), but the CTA is persistent and the (P+6)^2 halo window of the NEXT patch is in flight while the current one
// is computed: every window pixel is one contiguous row of C floats in the Z-ordered layout, so it is fetched with one
// bulk async copy (cp.async.bulk.shared.global, completion counted on an mbarrier) -- no register staging, no per-float4
// index arithmetic, and masked / out-of-image pixels are zero-filled with plain shared stores.  Two window buffers
// alternate; the taps are loaded once per CTA instead of once per patch.
//
// Replaces MinkowskiEngine/src/depthwise_convolution_kernel.cu:27-52 (forward) and :69-122 (backward).
#pragma once
#include "dwconv_tiled.cuh"
#include "gemm_tc.cuh"

namespace mpmae {
namespace pipe {

__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(tc::smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(tc::smem_u32(bar))
               : "memory");
}

// Source row (or -1) of window pixel wp of the patch (n, l)
template <int P>
__device__ __forceinline__ int64_t window_row(const int *__restrict__ slot_of, const Geo &g, int n, int l, int wp) {
  constexpr int W = P + 6, NBR = (3 + P - 1) / P;
  const int wy = wp / W, wx = wp - wy * W;
  const int ry = wy - 3 + NBR * P, rx = wx - 3 + NBR * P;   // >= 0
  const int by = ry / P, bx = rx / P;
  const int qy = l / g.G + by - NBR, qx = l % g.G + bx - NBR;
  if (qy < 0 || qx < 0 || qy >= g.G || qx >= g.G) return -1;
  const int slot = __ldg(slot_of + n * g.L + qy * g.G + qx);
  if (slot < 0) return -1;
  return ((int64_t)n * g.V + slot) * (P * P) + zorder3(ry - by * P, rx - bx * P);
}

// Every thread fetches its share of the window pixels of patch `pu` into `win` and arrives on `bar` (count = blockDim.x)
template <int P, int C>
__device__ __forceinline__ void issue_window(const float *__restrict__ x, const int *__restrict__ slot_of,
                                             const int *__restrict__ vis_patch, const Geo &g, int pu, float *win, uint64_t *bar) {
  constexpr int W = P + 6, NPIX = W * W;
  constexpr uint32_t kRowBytes = C * 4;
  const int n = pu / g.V, l = __ldg(vis_patch + pu);
  constexpr int kMaxPer = 4;   // NPIX <= 196, blockDim.x >= 64
  int64_t rows[kMaxPer];
  uint32_t bytes = 0;
#pragma unroll
  for (int k = 0; k < kMaxPer; ++k) {
    const int wp = (int)threadIdx.x + k * (int)blockDim.x;
    rows[k] = -2;
    if (wp < NPIX) {
      rows[k] = window_row<P>(slot_of, g, n, l, wp);
      if (rows[k] >= 0) {
        bytes += kRowBytes;
      } else {   // masked / out-of-image pixel: zeros, written before this thread arrives
        float4 *d = reinterpret_cast<float4 *>(win + (size_t)wp * C);
#pragma unroll
        for (int j = 0; j < C / 4; ++j) d[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  }
  if (bytes) tc::mbar_expect_tx(bar, bytes);   // arrive (release) + expect_tx
  else tc::mbar_arrive(bar);
#pragma unroll
  for (int k = 0; k < kMaxPer; ++k) {
    const int wp = (int)threadIdx.x + k * (int)blockDim.x;
    if (rows[k] >= 0) bulk_g2s(win + (size_t)wp * C, x + rows[k] * C, kRowBytes, bar);
  }
}

// ------------------------------------------------------------------------------------------------ forward / dX
template <int P, int TR, int C>
__global__ void __launch_bounds__(C *(P / TR)) dwconv_patch_pipe_kernel(DwTiledArgs t, int units) { pdl_prologue();
  constexpr int W = P + 6, TILES = P / TR, NPIX = W * W;
  const DwArgs &p = t.a;
  extern __shared__ __align__(128) float smem[];
  __shared__ __align__(8) uint64_t bar[2];
  float *win0 = smem, *win1 = smem + (size_t)NPIX * C;
  float *ubuf = win1 + (size_t)NPIX * C;     // [P*P][C]
  if (threadIdx.x == 0) {
    tc::mbar_init(&bar[0], blockDim.x);
    tc::mbar_init(&bar[1], blockDim.x);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  int pu = blockIdx.x;
  if (pu < units) issue_window<P, C>(p.x, p.slot_of, t.vis_patch, p.geo, pu, win0, &bar[0]);

  const int c = threadIdx.x % C, y0 = (threadIdx.x / C) * TR;
  float wreg[49];
#pragma unroll
  for (int k = 0; k < 49; ++k) {
    const int kh = k / 7, kw = k % 7;
    const int a = p.flip ? 6 - kh : kh, b = p.flip ? 6 - kw : kw;
    wreg[k] = __ldg(p.w + a * p.w_skh + b * p.w_skw + c * p.w_sc);
  }
  const float b0 = p.bias ? __ldg(p.bias + c) : 0.f;
  float4 csum = make_float4(0.f, 0.f, 0.f, 0.f);
  static_assert((C * (P / TR)) % (C / 4) == 0, "copy-out column ownership");

  for (int i = 0; pu < units; pu += gridDim.x, ++i) {
    const int b = i & 1;
    float *win = b ? win1 : win0;
    const int pn = pu + gridDim.x;
    if (pn < units) issue_window<P, C>(p.x, p.slot_of, t.vis_patch, p.geo, pn, b ? win0 : win1, &bar[b ^ 1]);
    tc::mbar_wait(&bar[b], (uint32_t)(i >> 1) & 1u);
    {
      float acc[TR][P];
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
    __syncthreads();   // window buffer b is free again, ubuf is complete
    const int64_t row0 = (int64_t)pu * (P * P);
    if (p.do_ln) {
      ln_tile_to_global(ubuf, P * P, C, p.eps, p.out + row0 * C, p.rstd + row0);
    } else {
      constexpr int n4 = P * P * C / 4;
      float4 *dst = reinterpret_cast<float4 *>(p.out + row0 * C);
      const float4 *res = p.resid ? reinterpret_cast<const float4 *>(p.resid + row0 * C) : nullptr;
      for (int k = threadIdx.x; k < n4; k += blockDim.x) {   // blockDim.x % (C/4) == 0: a thread keeps its 4 columns
        float4 v = reinterpret_cast<const float4 *>(ubuf)[k];
        if (res) { const float4 r = __ldg(res + k); v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w; }
        dst[k] = v;
        csum.x += v.x; csum.y += v.y; csum.z += v.z; csum.w += v.w;
      }
    }
    __syncthreads();   // ubuf is free again
  }
  if (p.colsum_out) {   // column sums of everything this CTA wrote: shared-memory reduce, one atomic per channel
    for (int k = threadIdx.x; k < C; k += blockDim.x) ubuf[k] = 0.f;
    __syncthreads();
    const int c4 = (threadIdx.x % (C / 4)) * 4;
    atomicAdd(&ubuf[c4], csum.x); atomicAdd(&ubuf[c4 + 1], csum.y); atomicAdd(&ubuf[c4 + 2], csum.z); atomicAdd(&ubuf[c4 + 3], csum.w);
    __syncthreads();
    for (int k = threadIdx.x; k < C; k += blockDim.x) atomicAdd(&p.colsum_out[k], ubuf[k]);
  }
}

// ------------------------------------------------------------------------------------------------ weight gradient
// dW[tap, c] += sum du[o, c] * x[o + off(tap), c] ; db[c] += sum du[o, c]: 49 partial sums per thread for the whole kernel
template <int P, int TR, int C>
__global__ void __launch_bounds__(C *(P / TR)) dwconv_patch_wgrad_pipe_kernel(DwWgradArgs p, const int *__restrict__ vis_patch,
                                                                             int units) { pdl_prologue();
  constexpr int W = P + 6, TILES = P / TR, NPIX = W * W, NOUT = P * P;
  extern __shared__ __align__(128) float smem[];
  __shared__ __align__(8) uint64_t bar[2];
  constexpr size_t kBuf = (size_t)(NPIX + NOUT) * C;   // [x window | du tile]
  if (threadIdx.x == 0) {
    tc::mbar_init(&bar[0], blockDim.x + 1);
    tc::mbar_init(&bar[1], blockDim.x + 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int pu, int b) {
    float *buf = smem + (size_t)b * kBuf;
    issue_window<P, C>(p.x, p.slot_of, vis_patch, p.geo, pu, buf, &bar[b]);
    if (threadIdx.x == 0) {   // the patch's du rows are contiguous: one bulk copy
      tc::mbar_expect_tx(&bar[b], NOUT * C * 4);
      bulk_g2s(buf + (size_t)NPIX * C, p.du + (int64_t)pu * NOUT * C, NOUT * C * 4, &bar[b]);
    }
  };
  int pu = blockIdx.x;
  if (pu < units) issue(pu, 0);
  const int c = threadIdx.x % C, y0 = (threadIdx.x / C) * TR;
  float dw[49];
  float db = 0.f;
#pragma unroll
  for (int k = 0; k < 49; ++k) dw[k] = 0.f;
  for (int i = 0; pu < units; pu += gridDim.x, ++i) {
    const int b = i & 1;
    const float *win = smem + (size_t)b * kBuf;
    const float *dus = win + (size_t)NPIX * C;
    const int pn = pu + gridDim.x;
    if (pn < units) issue(pn, b ^ 1);
    tc::mbar_wait(&bar[b], (uint32_t)(i >> 1) & 1u);
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
    __syncthreads();   // buffer b is free again
  }
  // reduce the TILES strips of each channel in shared memory, then one atomic per (tap, channel) per CTA
  float *red = smem;  // [50][C]
  for (int k = threadIdx.x; k < 50 * C; k += blockDim.x) red[k] = 0.f;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 49; ++k) atomicAdd(&red[k * C + c], dw[k]);
  atomicAdd(&red[49 * C + c], db);
  __syncthreads();
  for (int k = threadIdx.x; k < 49 * C; k += blockDim.x) {
    const int tap = k / C, cc = k - tap * C;
    const int kh = tap / 7, kw = tap - kh * 7;
    atomicAdd(&p.dw[kh * p.w_skh + kw * p.w_skw + cc * p.w_sc], red[k]);
  }
  if (p.dbias)
    for (int cc = threadIdx.x; cc < C; cc += blockDim.x) atomicAdd(&p.dbias[cc], red[49 * C + cc]);
}

template <int P, int TR, int C>
inline cudaError_t launch_patch_pipe(const DwTiledArgs &t, cudaStream_t st) {
  constexpr int threads = C * (P / TR);
  constexpr size_t sm = ((size_t)2 * (P + 6) * (P + 6) + P * P) * C * sizeof(float);
  static_assert(threads % 32 == 0 && threads <= 1024 && threads >= 64, "thread mapping");
  if (sm > 226 * 1024) return cudaErrorInvalidConfiguration;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(dwconv_patch_pipe_kernel<P, TR, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
    if (e != cudaSuccess) return e;
    (void)cudaFuncSetAttribute(dwconv_patch_pipe_kernel<P, TR, C>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    configured = true;
  }
  const int units = t.a.geo.B * t.a.geo.V;
  int per_sm = (int)((227 * 1024) / (sm + 1024));
  if (per_sm > 2048 / threads) per_sm = 2048 / threads;
  if (per_sm < 1) per_sm = 1;
  int grid = 148 * per_sm;
  if (grid > units) grid = units;
  pdl(dwconv_patch_pipe_kernel<P, TR, C>, grid, threads, sm, st)(t, units);
  return cudaGetLastError();
}

template <int P, int TR, int C>
inline cudaError_t launch_patch_wgrad_pipe(const DwWgradArgs &p, const int *vis_patch, cudaStream_t st) {
  constexpr int threads = C * (P / TR);
  constexpr size_t sm = (size_t)2 * ((P + 6) * (P + 6) + P * P) * C * sizeof(float);
  static_assert(sm >= (size_t)50 * C * sizeof(float), "reduction buffer");
  if (sm > 226 * 1024) return cudaErrorInvalidConfiguration;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(dwconv_patch_wgrad_pipe_kernel<P, TR, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
    if (e != cudaSuccess) return e;
    (void)cudaFuncSetAttribute(dwconv_patch_wgrad_pipe_kernel<P, TR, C>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    configured = true;
  }
  const int units = p.geo.B * p.geo.V;
  int per_sm = (int)((227 * 1024) / (sm + 1024));
  if (per_sm > 2048 / threads) per_sm = 2048 / threads;
  if (per_sm < 1) per_sm = 1;
  int grid = 148 * per_sm;
  if (grid > units) grid = units;
  pdl(dwconv_patch_wgrad_pipe_kernel<P, TR, C>, grid, threads, sm, st)(p, vis_patch, units);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ stage with 2x2 patches
// Per-sample processing on a zero-padded 20x20 cell grid (G = 7), 32-channel chunks, one warp per visible patch (lane =
// channel), persistent CTAs.  The 4V visible rows of the next (sample, chunk) land in the other grid buffer through
// 128-byte bulk async copies while this one is computed; only the cells a sample touched are re-zeroed afterwards.
constexpr int kS2GW = 20, kS2Cells = kS2GW * kS2GW, kS2CC = 32;

// Forward / dX: one CTA per sample.  The sample's 4V visible rows are contiguous in HBM, so they arrive with ONE bulk async
// copy ([4V][C] floats); a 20x20 table maps every padded grid cell to the byte offset of its row (masked / outside cells
// point at a zero row), so no dense grid is filled or cleared.  Warp = (32-channel chunk, patch subset): the 49 taps of a
// lane's channel are loaded once; a window row of 8 cells costs four uniform 8-byte table loads + 8 data loads.
__global__ void __launch_bounds__(768) dwconv_s2_pipe_kernel(DwArgs p, const int *__restrict__ vis_patch) { pdl_prologue();
  extern __shared__ __align__(128) float smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ __align__(16) uint32_t tab[kS2Cells];
  __shared__ int vis[32];
  const int C = p.C, V = p.geo.V, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  const int nchunks = C / kS2CC, rows = V * 4, n = blockIdx.x;
  float *xs = smem;                          // [rows + 1][C], last row = zeros
  float *ubuf = smem + (size_t)(rows + 1) * C;   // [rows][C]
  const int64_t row0 = (int64_t)n * rows;
  if (tid == 0) {
    tc::mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    tc::mbar_expect_tx(&bar, (uint32_t)(rows * C * 4));
    bulk_g2s(xs, p.x + row0 * C, (uint32_t)(rows * C * 4), &bar);
  }
  for (int i = tid; i < C; i += blockDim.x) xs[(size_t)rows * C + i] = 0.f;
  for (int i = tid; i < kS2Cells; i += blockDim.x) tab[i] = (uint32_t)(rows * C * 4);
  if (tid < V) vis[tid] = __ldg(vis_patch + n * V + tid);
  __syncthreads();
  if (tid < rows) {
    const int l = vis[tid >> 2], m = tid & 3;
    tab[((l / 7) * 2 + (m & 1) + 3) * kS2GW + (l % 7) * 2 + (m >> 1) + 3] = (uint32_t)(tid * C * 4);
  }
  const int chunk = warp % nchunks, pg = warp / nchunks, npg = nw / nchunks;
  const int c = chunk * kS2CC + lane;
  float w[49];
#pragma unroll
  for (int t = 0; t < 49; ++t) {
    const int kh = t / 7, kw = t % 7;
    const int a = p.flip ? 6 - kh : kh, bb = p.flip ? 6 - kw : kw;
    w[t] = __ldg(p.w + a * p.w_skh + bb * p.w_skw + c * p.w_sc);
  }
  const float b0 = p.bias ? __ldg(p.bias + c) : 0.f;
  __syncthreads();
  tc::mbar_wait(&bar, 0);
  if (pg < npg) {
    const char *xc = reinterpret_cast<const char *>(xs + c);
    for (int slot = pg; slot < V; slot += npg) {
      const int l = vis[slot];
      const uint32_t *trow = tab + (l / 7) * 2 * kS2GW + (l % 7) * 2;   // window origin (8-byte aligned: even column)
      float acc[2][2] = {{b0, b0}, {b0, b0}};
#pragma unroll
      for (int iy = 0; iy < 8; ++iy) {
        float in[8];
#pragma unroll
        for (int q = 0; q < 8; q += 2) {
          const uint2 o = *reinterpret_cast<const uint2 *>(trow + iy * kS2GW + q);
          in[q] = *reinterpret_cast<const float *>(xc + o.x);
          in[q + 1] = *reinterpret_cast<const float *>(xc + o.y);
        }
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
        for (int ox = 0; ox < 2; ++ox) ubuf[(size_t)(slot * 4 + (r | (ox << 1))) * C + c] = acc[r][ox];
    }
  }
  __syncthreads();
  if (p.do_ln) {
    ln_tile_to_global(ubuf, rows, C, p.eps, p.out + row0 * C, p.rstd + row0);
  } else {
    const int n4 = rows * C / 4;
    float4 *dst = reinterpret_cast<float4 *>(p.out + row0 * C);
    const float4 *res = p.resid ? reinterpret_cast<const float4 *>(p.resid + row0 * C) : nullptr;
    float4 csum = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = tid; i < n4; i += blockDim.x) {   // blockDim.x % (C/4) == 0 when colsum_out is set (launcher)
      float4 v = reinterpret_cast<const float4 *>(ubuf)[i];
      if (res) { const float4 r = __ldg(res + i); v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w; }
      dst[i] = v;
      csum.x += v.x; csum.y += v.y; csum.z += v.z; csum.w += v.w;
    }
    if (p.colsum_out) {
      __syncthreads();
      for (int k = tid; k < C; k += blockDim.x) xs[k] = 0.f;
      __syncthreads();
      const int c4 = (tid % (C / 4)) * 4;
      atomicAdd(&xs[c4], csum.x); atomicAdd(&xs[c4 + 1], csum.y); atomicAdd(&xs[c4 + 2], csum.z); atomicAdd(&xs[c4 + 3], csum.w);
      __syncthreads();
      for (int k = tid; k < C; k += blockDim.x) atomicAdd(&p.colsum_out[k], xs[k]);
    }
  }
}

// Weight gradient: persistent CTAs over samples, warp = (32-channel chunk, patch subset) so a lane keeps its channel's 49
// partial sums in registers for the whole kernel.  Per sample two bulk async copies (x rows, du rows), double buffered,
// and a cell -> row-offset table computed straight from the slot table.
__global__ void __launch_bounds__(768) dwconv_s2_wgrad_pipe_kernel(DwWgradArgs p, const int *__restrict__ vis_patch) { pdl_prologue();
  extern __shared__ __align__(128) float smem[];
  __shared__ __align__(8) uint64_t bar[2];
  __shared__ __align__(16) uint32_t tab[2][kS2Cells];
  __shared__ int vis[2][32];
  const int C = p.C, V = p.geo.V, B = p.geo.B, L = p.geo.L, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5,
            nw = blockDim.x >> 5;
  const int nchunks = C / kS2CC, rows = V * 4;
  const size_t kTile = (size_t)rows * C;             // floats per [rows][C] tile
  float *zrow = smem + 4 * kTile;                    // [x0 | du0 | x1 | du1 | zero row]
  for (int i = tid; i < C; i += blockDim.x) zrow[i] = 0.f;
  if (tid == 0) {
    tc::mbar_init(&bar[0], 1);
    tc::mbar_init(&bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int total = (B - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  auto issue = [&](int j) {
    const int n = (int)blockIdx.x + j * (int)gridDim.x, b = j & 1;
    float *xs = smem + (size_t)b * 2 * kTile;
    if (tid == 0) {
      tc::mbar_expect_tx(&bar[b], (uint32_t)(2 * kTile * 4));
      bulk_g2s(xs, p.x + (int64_t)n * kTile, (uint32_t)(kTile * 4), &bar[b]);
      bulk_g2s(xs + kTile, p.du + (int64_t)n * kTile, (uint32_t)(kTile * 4), &bar[b]);
    }
    for (int i = tid; i < kS2Cells; i += blockDim.x) {   // byte offset of every padded grid cell's row relative to xs
      const int gy = i / kS2GW - 3, gx = i % kS2GW - 3;
      uint32_t off = (uint32_t)((zrow - xs) * 4);
      if (gy >= 0 && gx >= 0 && gy < 14 && gx < 14) {
        const int slot = __ldg(p.slot_of + n * L + (gy >> 1) * 7 + (gx >> 1));
        if (slot >= 0) off = (uint32_t)((slot * 4 + ((gy & 1) | ((gx & 1) << 1))) * C * 4);
      }
      tab[b][i] = off;
    }
    if (tid < V) vis[b][tid] = __ldg(vis_patch + n * V + tid);
  };
  const int chunk = warp % nchunks, pg = warp / nchunks, npg = nw / nchunks;
  const int c = chunk * kS2CC + lane;
  float dw[49];
  float db = 0.f;
#pragma unroll
  for (int t = 0; t < 49; ++t) dw[t] = 0.f;
  if (total > 0) issue(0);
  for (int j = 0; j < total; ++j) {
    const int b = j & 1;
    if (j + 1 < total) issue(j + 1);
    __syncthreads();                                    // tab / vis of buffer b (written one iteration ago) are visible
    tc::mbar_wait(&bar[b], (uint32_t)(j >> 1) & 1u);
    if (pg < npg) {
      const float *xs = smem + (size_t)b * 2 * kTile;
      const char *xc = reinterpret_cast<const char *>(xs + c);
      const float *dc = xs + kTile + c;
      for (int slot = pg; slot < V; slot += npg) {
        const int l = vis[b][slot];
        const uint32_t *trow = &tab[b][(l / 7) * 2 * kS2GW + (l % 7) * 2];
        float dd[2][2];
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
          for (int ox = 0; ox < 2; ++ox) { dd[r][ox] = dc[(size_t)(slot * 4 + (r | (ox << 1))) * C]; db += dd[r][ox]; }
#pragma unroll
        for (int iy = 0; iy < 8; ++iy) {
          float in[8];
#pragma unroll
          for (int q = 0; q < 8; q += 2) {
            const uint2 o = *reinterpret_cast<const uint2 *>(trow + iy * kS2GW + q);
            in[q] = *reinterpret_cast<const float *>(xc + o.x);
            in[q + 1] = *reinterpret_cast<const float *>(xc + o.y);
          }
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            const int kh = iy - r;
            if (kh >= 0 && kh < 7) {
#pragma unroll
              for (int kw = 0; kw < 7; ++kw)
                dw[kh * 7 + kw] = fmaf(dd[r][0], in[kw], fmaf(dd[r][1], in[kw + 1], dw[kh * 7 + kw]));
            }
          }
        }
      }
    }
    __syncthreads();                                    // buffer b (tiles, tab, vis) is free again
  }
  // warps with the same chunk hold partial sums of the same 32 channels: reduce in shared memory, then global atomics
  float *red = smem;   // [50][C]
  for (int i = tid; i < 50 * C; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
  if (pg < npg) {
#pragma unroll
    for (int t = 0; t < 49; ++t) atomicAdd(&red[t * C + c], dw[t]);
    atomicAdd(&red[49 * C + c], db);
  }
  __syncthreads();
  for (int i = tid; i < 49 * C; i += blockDim.x) {
    const int t = i / C, cc = i - t * C;
    const int kh = t / 7, kw = t - kh * 7;
    atomicAdd(&p.dw[kh * p.w_skh + kw * p.w_skw + cc * p.w_sc], red[i]);
  }
  if (p.dbias)
    for (int cc = tid; cc < C; cc += blockDim.x) atomicAdd(&p.dbias[cc], red[49 * C + cc]);
}

inline cudaError_t launch_s2_pipe(const DwArgs &a, const int *vis_patch, cudaStream_t st) {
  if (a.geo.G != 7 || a.P != 2 || a.C % kS2CC != 0 || a.geo.V > 32 || (a.do_ln && a.resid)) return cudaErrorInvalidConfiguration;
  const int np = ln_parts(a.C), f4 = (a.C >> 2) / np;
  if (a.do_ln && (f4 < 1 || f4 > 6)) return cudaErrorInvalidConfiguration;
  const int rows = a.geo.V * 4, nchunks = a.C / kS2CC;
  const size_t sm = (size_t)(2 * rows + 1) * a.C * sizeof(float);
  if (sm > 222 * 1024 || nchunks > 24 || ((size_t)rows * a.C * 4) % 16 != 0) return cudaErrorInvalidConfiguration;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(dwconv_s2_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 222 * 1024);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  // two CTAs per SM (smem and registers permitting) hide each other's load phase: ~10 warps per CTA
  int npg = (2 * sm + 4096 <= 227 * 1024) ? 10 / nchunks : 20 / nchunks;
  if (npg < 1) npg = 1;
  int warps = nchunks * npg;
  if (warps * 32 < rows) warps = (rows + 31) / 32;
  if (a.colsum_out && (a.do_ln || (warps * 32) % (a.C / 4) != 0)) return cudaErrorInvalidConfiguration;
  pdl(dwconv_s2_pipe_kernel, a.geo.B, warps * 32, sm, st)(a, vis_patch);
  return cudaGetLastError();
}
inline cudaError_t launch_s2_wgrad_pipe(const DwWgradArgs &p, const int *vis_patch, cudaStream_t st) {
  if (p.geo.G != 7 || p.P != 2 || p.C % kS2CC != 0 || p.geo.V > 32 || !p.slot_of) return cudaErrorInvalidConfiguration;
  const int rows = p.geo.V * 4, nchunks = p.C / kS2CC;
  const size_t sm = ((size_t)4 * rows + 1) * p.C * sizeof(float);
  if (sm > 218 * 1024 || nchunks > 24 || ((size_t)rows * p.C * 4) % 16 != 0 || 50 > 4 * rows) return cudaErrorInvalidConfiguration;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(dwconv_s2_wgrad_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 218 * 1024);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  int npg = 20 / nchunks;
  if (npg < 1) npg = 1;
  const int warps = nchunks * npg;
  const int grid = p.geo.B < 148 ? p.geo.B : 148;
  pdl(dwconv_s2_wgrad_pipe_kernel, grid, warps * 32, sm, st)(p, vis_patch);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ merged backward
// dX and dW of the depthwise convolution from ONE read of the du halo window (round 2).
//   dx[i]       = dy[i] + sum_k du[i + k - 3] * w[6 - k]                       (transposed stencil + residual)
//   dW[6 - k]  += sum_{i in this patch} x[i] * du[i + k - 3]                     (every (input, output) pair belongs to the
//   db         += sum_{i in this patch} du[i]                                    patch of its INPUT pixel: counted once)
// Both sums walk the same window values: a thread loads a window row once and issues two FMAs per (value, output) pair.
// Needs the du halo window (as the dX kernel), the patch's own x rows (one bulk copy: they are contiguous) and nothing else;
// the separate weight-gradient kernel fetched an x halo window AND the du tile.
struct DwBwdArgs {
  DwArgs a;               // a.x = du (conv input), a.w = forward taps, a.flip = 1, a.resid = dy, a.out = dx, a.colsum_out
  const float *xfwd;      // [R, C] forward input of the depthwise convolution
  float *dw;              // parameter-layout gradient (same strides as a.w)
  float *dbias;           // [C]
  const int *vis_patch;
};

template <int P, int TR, int C>
__global__ void __launch_bounds__(C *(P / TR)) dwconv_patch_bwd_pipe_kernel(DwBwdArgs t, int units) { pdl_prologue();
  constexpr int W = P + 6, NPIX = W * W, NOUT = P * P;
  const DwArgs &p = t.a;
  extern __shared__ __align__(128) float smem[];
  __shared__ __align__(8) uint64_t bar[2];
  constexpr size_t kBuf = (size_t)(NPIX + NOUT) * C;   // [du window | x tile]
  float *ubuf = smem + 2 * kBuf;                       // [P*P][C]
  if (threadIdx.x == 0) {
    tc::mbar_init(&bar[0], blockDim.x + 1);
    tc::mbar_init(&bar[1], blockDim.x + 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int pu, int b) {
    float *buf = smem + (size_t)b * kBuf;
    issue_window<P, C>(p.x, p.slot_of, t.vis_patch, p.geo, pu, buf, &bar[b]);
    if (threadIdx.x == 0) {   // the patch's x rows are contiguous: one bulk copy
      tc::mbar_expect_tx(&bar[b], NOUT * C * 4);
      bulk_g2s(buf + (size_t)NPIX * C, t.xfwd + (int64_t)pu * NOUT * C, NOUT * C * 4, &bar[b]);
    }
  };
  int pu = blockIdx.x;
  if (pu < units) issue(pu, 0);
  const int c = threadIdx.x % C, y0 = (threadIdx.x / C) * TR;
  float wreg[49], dw[49];
  float db = 0.f;
#pragma unroll
  for (int k = 0; k < 49; ++k) {
    const int kh = k / 7, kw = k % 7;
    wreg[k] = __ldg(p.w + (6 - kh) * p.w_skh + (6 - kw) * p.w_skw + c * p.w_sc);   // transposed stencil
    dw[k] = 0.f;
  }
  float4 csum = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i = 0; pu < units; pu += gridDim.x, ++i) {
    const int b = i & 1;
    const float *win = smem + (size_t)b * kBuf;
    const float *xt = win + (size_t)NPIX * C;
    const int pn = pu + gridDim.x;
    if (pn < units) issue(pn, b ^ 1);
    tc::mbar_wait(&bar[b], (uint32_t)(i >> 1) & 1u);
    {
      float acc[TR][P], xv[TR][P];
#pragma unroll
      for (int r = 0; r < TR; ++r)
#pragma unroll
        for (int ox = 0; ox < P; ++ox) { acc[r][ox] = 0.f; xv[r][ox] = xt[(size_t)zorder3(y0 + r, ox) * C + c]; }
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
              for (int ox = 0; ox < P; ++ox) {
                acc[r][ox] = fmaf(in[ox + kw], wreg[kh * 7 + kw], acc[r][ox]);
                dw[kh * 7 + kw] = fmaf(xv[r][ox], in[ox + kw], dw[kh * 7 + kw]);   // slot (kh, kw) holds dW[6 - kh][6 - kw]
              }
            if (kh == 3) {
#pragma unroll
              for (int ox = 0; ox < P; ++ox) db += in[ox + 3];                      // du at the patch's own pixels
            }
          }
        }
      }
#pragma unroll
      for (int r = 0; r < TR; ++r)
#pragma unroll
        for (int ox = 0; ox < P; ++ox) ubuf[(size_t)zorder3(y0 + r, ox) * C + c] = acc[r][ox];
    }
    __syncthreads();   // buffer b is free again, ubuf is complete
    const int64_t row0 = (int64_t)pu * (P * P);
    constexpr int n4 = P * P * C / 4;
    float4 *dst = reinterpret_cast<float4 *>(p.out + row0 * C);
    const float4 *res = p.resid ? reinterpret_cast<const float4 *>(p.resid + row0 * C) : nullptr;
    for (int k = threadIdx.x; k < n4; k += blockDim.x) {   // blockDim.x % (C/4) == 0: a thread keeps its 4 columns
      float4 v = reinterpret_cast<const float4 *>(ubuf)[k];
      if (res) { const float4 r = __ldg(res + k); v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w; }
      dst[k] = v;
      csum.x += v.x; csum.y += v.y; csum.z += v.z; csum.w += v.w;
    }
    __syncthreads();   // ubuf is free again
  }
  // reductions: column sums of dx (one atomic per channel), then the strips of each channel's dW / db
  float *red = smem;  // [50][C] (+ [C] for the column sums behind it)
  for (int k = threadIdx.x; k < 51 * C; k += blockDim.x) red[k] = 0.f;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 49; ++k) atomicAdd(&red[k * C + c], dw[k]);
  atomicAdd(&red[49 * C + c], db);
  if (p.colsum_out) {
    const int c4 = (threadIdx.x % (C / 4)) * 4;
    atomicAdd(&red[50 * C + c4], csum.x); atomicAdd(&red[50 * C + c4 + 1], csum.y);
    atomicAdd(&red[50 * C + c4 + 2], csum.z); atomicAdd(&red[50 * C + c4 + 3], csum.w);
  }
  __syncthreads();
  for (int k = threadIdx.x; k < 49 * C; k += blockDim.x) {
    const int tap = k / C, cc = k - tap * C;
    const int kh = 6 - tap / 7, kw = 6 - tap % 7;          // slot (a, b) accumulated dW[6 - a][6 - b]
    atomicAdd(&t.dw[kh * p.w_skh + kw * p.w_skw + cc * p.w_sc], red[k]);
  }
  for (int cc = threadIdx.x; cc < C; cc += blockDim.x) {
    if (t.dbias) atomicAdd(&t.dbias[cc], red[49 * C + cc]);
    if (p.colsum_out) atomicAdd(&p.colsum_out[cc], red[50 * C + cc]);
  }
}

template <int P, int TR, int C>
inline cudaError_t launch_patch_bwd_pipe(const DwBwdArgs &t, cudaStream_t st) {
  constexpr int threads = C * (P / TR);
  constexpr size_t sm = ((size_t)2 * ((P + 6) * (P + 6) + P * P) + P * P) * C * sizeof(float);
  static_assert(threads % 32 == 0 && threads <= 1024 && threads >= 64 && threads % (C / 4) == 0, "thread mapping");
  static_assert(sm >= (size_t)51 * C * sizeof(float), "reduction buffer");
  if (sm > 226 * 1024) return cudaErrorInvalidConfiguration;
  static bool configured = false;
  static int per_sm = 1;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(dwconv_patch_bwd_pipe_kernel<P, TR, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
    if (e != cudaSuccess) return e;
    (void)cudaFuncSetAttribute(dwconv_patch_bwd_pipe_kernel<P, TR, C>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    int v = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, dwconv_patch_bwd_pipe_kernel<P, TR, C>, threads, sm) != cudaSuccess || v < 1) {
      (void)cudaGetLastError();
      v = 1;
    }
    per_sm = v;
    configured = true;
  }
  const int units = t.a.geo.B * t.a.geo.V;
  int grid = 148 * per_sm;
  if (grid > units) grid = units;
  pdl(dwconv_patch_bwd_pipe_kernel<P, TR, C>, grid, threads, sm, st)(t, units);
  return cudaGetLastError();
}

}  // namespace pipe

// Dispatch for the instantiated (P, C) pairs; cudaErrorInvalidConfiguration = not taken (caller falls back)
inline cudaError_t launch_dwconv_pipe(const DwArgs &a, const int *vis_patch, cudaStream_t st) {
  if (!vis_patch || !a.slot_of) return cudaErrorInvalidConfiguration;
  if (a.do_ln && (a.resid || a.colsum_out)) return cudaErrorInvalidConfiguration;
  DwTiledArgs t{a, vis_patch};
  if (a.P == 8 && a.C == 40) return pipe::launch_patch_pipe<8, 2, 40>(t, st);
  if (a.P == 8 && a.C == 96) return pipe::launch_patch_pipe<8, 2, 96>(t, st);
  if (a.P == 4 && a.C == 80) return pipe::launch_patch_pipe<4, 2, 80>(t, st);
  if (a.P == 4 && a.C == 192) return pipe::launch_patch_pipe<4, 2, 192>(t, st);
  if (a.P == 2) return pipe::launch_s2_pipe(a, vis_patch, st);
  return cudaErrorInvalidConfiguration;
}
// merged dX + dW (+ db, + column sums of dX) for the patch stages; cudaErrorInvalidConfiguration = not taken
inline cudaError_t launch_dwconv_bwd_pipe(const pipe::DwBwdArgs &t, cudaStream_t st) {
  const DwArgs &a = t.a;
  if (!t.vis_patch || !a.slot_of || a.do_ln || !a.flip || !t.xfwd || !t.dw) return cudaErrorInvalidConfiguration;
  if (a.P == 8 && a.C == 40) return pipe::launch_patch_bwd_pipe<8, 2, 40>(t, st);
  if (a.P == 8 && a.C == 96) return pipe::launch_patch_bwd_pipe<8, 2, 96>(t, st);
  if (a.P == 4 && a.C == 80) return pipe::launch_patch_bwd_pipe<4, 2, 80>(t, st);
  if (a.P == 4 && a.C == 192) return pipe::launch_patch_bwd_pipe<4, 2, 192>(t, st);
  return cudaErrorInvalidConfiguration;
}
inline cudaError_t launch_dwconv_wgrad_pipe(const DwWgradArgs &p, const int *vis_patch, cudaStream_t st) {
  if (!vis_patch || !p.slot_of) return cudaErrorInvalidConfiguration;
  if (p.P == 8 && p.C == 40) return pipe::launch_patch_wgrad_pipe<8, 2, 40>(p, vis_patch, st);
  if (p.P == 8 && p.C == 96) return pipe::launch_patch_wgrad_pipe<8, 2, 96>(p, vis_patch, st);
  if (p.P == 4 && p.C == 80) return pipe::launch_patch_wgrad_pipe<4, 2, 80>(p, vis_patch, st);
  if (p.P == 4 && p.C == 192) return pipe::launch_patch_wgrad_pipe<4, 2, 192>(p, vis_patch, st);
  if (p.P == 2) return pipe::launch_s2_wgrad_pipe(p, vis_patch, st);
  return cudaErrorInvalidConfiguration;
}

}  // namespace mpmae
