// Depthwise 7x7 kernels for the patch stages (P = 8, 4), second generation: packed FP32 math and warp-parallel async loads.
//
// What the ncu source page of the first-generation kernels (dwconv_pipe.cuh) showed (profiles/r2_dwconv_before.md): half of
// the issued instructions were not FMAs -- every window pixel was fetched with its own bulk async copy, which the hardware
// serialises lane by lane (ELECT / R2UR / UBLKCP sequences, ~10 instructions per 160-byte copy) -- and a quarter of the
// stall samples sat on the mbarrier wait.  The FMA pipe was 36 % busy.
//
// Here:
//   * a thread owns a PAIR of adjacent channels and a strip of two output rows; the 49 taps of both channels live in 49
//     64-bit registers and every multiply-add is one FFMA2 (fma.rn.f32x2, sm_100): the FMA pipe is saturated with half
//     the issue slots, the other half is left for the shared-memory loads (one LDS.64 per window pixel per thread);
//   * the halo window is fetched with 16-byte cp.async (LDGSTS) instructions issued by all threads in parallel: a thread
//     resolves the source row of a window pixel once and copies its C floats with immediate offsets; masked / outside
//     pixels are zero-filled with plain stores;
//   * CTAs are small (one patch in flight, <= 42 KB of shared memory) and persistent, 4-5 per SM, so the load, compute
//     and LayerNorm / copy-out phases of different CTAs overlap; the next window is requested before the LayerNorm /
//     copy-out phase of the current patch.
//
// Same maths as MinkowskiEngine/src/depthwise_convolution_kernel.cu:27-52 (forward) and :69-122 (backward).
#pragma once
#include "dwconv_pipe.cuh"

namespace mpmae {
namespace dw2 {

__device__ __forceinline__ unsigned long long pack2(float a, float b) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float &a, float &b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ void cp_async16(void *dst, const void *src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(tc::smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template <int P, int C>
struct Shape {
  static constexpr int W = P + 6, NPIX = W * W, NOUT = P * P, CP = C / 2, STRIPS = P / 2;
  static constexpr int ACTIVE = CP * STRIPS;                    // threads that compute
  // window row stride in floats: padded so that two strips sharing a warp hit disjoint banks (a thread reads 8 bytes at
  // 2 cp + 2 RS strip floats: conflict-free within a half-warp when RS = C / 2 mod 16)
  static constexpr int PAD = ((C / 2 - W * C) % 16 + 16) % 16;
  static constexpr int RS = W * C + PAD;
  static constexpr int WIN = W * RS;                            // floats of a window buffer
  static_assert(PAD % 4 == 0, "16-byte aligned pixels");
  static constexpr int NT = (ACTIVE + 31) / 32 * 32;            // block size
  static constexpr int MINB = NT <= 96 ? 4 : 2;                 // CTAs per SM the register budget is sized for (<= 168 regs)
  static constexpr int NBUF = ((2 * WIN + NOUT * C) * 4 + 1024) * 3 <= 227 * 1024 ? 2 : 1;   // forward: window buffers
  static_assert(C % 4 == 0 && P % 2 == 0, "channel pairs, 16-byte chunks, two-row strips");
};

// All threads: fetch the (P+6)^2 halo window of visible patch `pu` into win[wp][C] (zeros where nothing is active)
template <int P, int C>
__device__ __forceinline__ void load_window(const float *__restrict__ x, const int *__restrict__ slot_of,
                                            const int *__restrict__ vis_patch, const Geo &g, int pu, float *win) {
  using S = Shape<P, C>;
  const int n = pu / g.V, l = __ldg(vis_patch + pu);
  for (int wp = threadIdx.x; wp < S::NPIX; wp += S::NT) {
    const int64_t row = pipe::window_row<P>(slot_of, g, n, l, wp);
    float *d = win + (size_t)(wp / S::W) * S::RS + (size_t)(wp % S::W) * C;
    if (row >= 0) {
      const float *s = x + row * C;
#pragma unroll
      for (int q = 0; q < C / 4; ++q) cp_async16(d + 4 * q, s + 4 * q);
    } else {
#pragma unroll
      for (int q = 0; q < C / 4; ++q) reinterpret_cast<float4 *>(d)[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}

// ------------------------------------------------------------------------------------------------ forward / dX
template <int P, int C>
__global__ void __launch_bounds__(Shape<P, C>::NT, (Shape<P, C>::NBUF == 2 && Shape<P, C>::NT <= 96) ? 3 : Shape<P, C>::MINB) dwconv_patch_v2_kernel(DwTiledArgs t, int units) { pdl_prologue();
  using S = Shape<P, C>;
  constexpr int W = S::W;
  const DwArgs &p = t.a;
  extern __shared__ __align__(128) float smem[];
  // window buffers ([W][RS]: W pixels of C floats per row, padded): two when they fit next to the output tile with >= 3
  // CTAs per SM (the window of the NEXT patch then loads during the whole compute phase), else one
  constexpr int NBUF = S::NBUF;
  float *win0 = smem;
  float *ubuf = smem + (size_t)NBUF * S::WIN;     // [P*P][C]
  int pu = blockIdx.x;
  if (pu < units) load_window<P, C>(p.x, p.slot_of, t.vis_patch, p.geo, pu, win0);
  cp_async_commit();

  const bool active = threadIdx.x < S::ACTIVE;
  const int cp = active ? threadIdx.x % S::CP : 0, y0 = active ? (threadIdx.x / S::CP) * 2 : 0;
  unsigned long long w2[49];
#pragma unroll
  for (int k = 0; k < 49; ++k) {
    const int kh = k / 7, kw = k % 7;
    const int a = p.flip ? 6 - kh : kh, b = p.flip ? 6 - kw : kw;
    const float *wp = p.w + a * p.w_skh + b * p.w_skw + 2 * cp * p.w_sc;
    w2[k] = pack2(__ldg(wp), __ldg(wp + p.w_sc));
  }
  const unsigned long long b2 = p.bias ? pack2(__ldg(p.bias + 2 * cp), __ldg(p.bias + 2 * cp + 1)) : 0ull;
  constexpr int kCols = C / 4;                                   // float4 columns of a row
  constexpr int kStride = (S::NT / kCols) * kCols;               // copy-out: a thread keeps its 4 columns
  float4 csum = make_float4(0.f, 0.f, 0.f, 0.f);

  for (int it = 0; pu < units; pu += gridDim.x, ++it) {
    float *win = win0 + (size_t)(NBUF == 2 ? (it & 1) : 0) * S::WIN;
    const int pn = pu + gridDim.x;
    cp_async_wait_all();
    __syncthreads();                               // the window is complete; the previous patch's ubuf has been read
    if (NBUF == 2) {                               // the other buffer was last read before the barrier above
      if (pn < units) load_window<P, C>(p.x, p.slot_of, t.vis_patch, p.geo, pn, win0 + (size_t)((it + 1) & 1) * S::WIN);
      cp_async_commit();
    }
    if (active) {
      unsigned long long acc[2][P];
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int ox = 0; ox < P; ++ox) acc[r][ox] = b2;
      const char *wbase = reinterpret_cast<const char *>(win) + ((size_t)y0 * S::RS + 2 * cp) * 4;
#pragma unroll
      for (int iy = 0; iy < 8; ++iy) {
#pragma unroll
        for (int j = 0; j < W; ++j) {
          const unsigned long long v = *reinterpret_cast<const unsigned long long *>(wbase + (size_t)(iy * S::RS + j * C) * 4);
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            const int kh = iy - r;
            if (kh >= 0 && kh < 7) {
#pragma unroll
              for (int ox = 0; ox < P; ++ox) {
                const int kw = j - ox;
                if (kw >= 0 && kw < 7) acc[r][ox] = fma2(v, w2[kh * 7 + kw], acc[r][ox]);
              }
            }
          }
        }
      }
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int ox = 0; ox < P; ++ox)
          *reinterpret_cast<unsigned long long *>(ubuf + (size_t)zorder3(y0 + r, ox) * C + 2 * cp) = acc[r][ox];
    }
    __syncthreads();                               // the window is free again, ubuf is complete
    if (NBUF == 1) {
      if (pn < units) load_window<P, C>(p.x, p.slot_of, t.vis_patch, p.geo, pn, win);   // in flight during the phase below
      cp_async_commit();
    }
    const int64_t row0 = (int64_t)pu * (P * P);
    if (p.do_ln) {
      ln_tile_to_global(ubuf, P * P, C, p.eps, p.out + row0 * C, p.rstd + row0);
    } else {
      constexpr int n4 = P * P * C / 4;
      float4 *dst = reinterpret_cast<float4 *>(p.out + row0 * C);
      const float4 *res = p.resid ? reinterpret_cast<const float4 *>(p.resid + row0 * C) : nullptr;
      if (threadIdx.x < kStride) {
        for (int k = threadIdx.x; k < n4; k += kStride) {
          float4 v = reinterpret_cast<const float4 *>(ubuf)[k];
          if (res) { const float4 r = __ldg(res + k); v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w; }
          dst[k] = v;
          csum.x += v.x; csum.y += v.y; csum.z += v.z; csum.w += v.w;
        }
      }
    }
  }
  cp_async_wait_all();
  if (p.colsum_out) {   // column sums of everything this CTA wrote: shared-memory reduce, one atomic per channel
    __syncthreads();
    for (int k = threadIdx.x; k < C; k += S::NT) ubuf[k] = 0.f;
    __syncthreads();
    if (threadIdx.x < kStride) {
      const int c4 = (threadIdx.x % kCols) * 4;
      atomicAdd(&ubuf[c4], csum.x); atomicAdd(&ubuf[c4 + 1], csum.y); atomicAdd(&ubuf[c4 + 2], csum.z); atomicAdd(&ubuf[c4 + 3], csum.w);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < C; k += S::NT) atomicAdd(&p.colsum_out[k], ubuf[k]);
  }
}

// ------------------------------------------------------------------------------------------------ weight gradient
// dW[tap, c] += sum du[o, c] * x[o + off(tap), c] ; db[c] += sum du[o, c]: 49 packed partial sums per thread for the whole kernel
template <int P, int C>
__global__ void __launch_bounds__(Shape<P, C>::NT, Shape<P, C>::MINB) dwconv_patch_wgrad_v2_kernel(DwWgradArgs p, const int *__restrict__ vis_patch,
                                                                                  int units) { pdl_prologue();
  using S = Shape<P, C>;
  constexpr int W = S::W;
  extern __shared__ __align__(128) float smem[];
  float *win = smem;                              // [W][RS]    x halo window (padded rows)
  float *dus = smem + (size_t)S::WIN;             // [P*P][C]   du tile
  auto load = [&](int pu) {
    load_window<P, C>(p.x, p.slot_of, vis_patch, p.geo, pu, win);
    const float *src = p.du + (int64_t)pu * S::NOUT * C;   // the patch's du rows are contiguous
    for (int k = threadIdx.x; k < S::NOUT * C / 4; k += S::NT) cp_async16(dus + 4 * k, src + 4 * k);
  };
  int pu = blockIdx.x;
  if (pu < units) load(pu);
  cp_async_commit();
  const bool active = threadIdx.x < S::ACTIVE;
  const int cp = active ? threadIdx.x % S::CP : 0, y0 = active ? (threadIdx.x / S::CP) * 2 : 0;
  unsigned long long dw2[49];
  unsigned long long db2 = 0ull;
#pragma unroll
  for (int k = 0; k < 49; ++k) dw2[k] = 0ull;
  for (; pu < units; pu += gridDim.x) {
    cp_async_wait_all();
    __syncthreads();
    if (active) {
      unsigned long long d[2][P];
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int ox = 0; ox < P; ++ox) {
          d[r][ox] = *reinterpret_cast<const unsigned long long *>(dus + (size_t)zorder3(y0 + r, ox) * C + 2 * cp);
          db2 = add2(db2, d[r][ox]);
        }
      const char *wbase = reinterpret_cast<const char *>(win) + ((size_t)y0 * S::RS + 2 * cp) * 4;
#pragma unroll
      for (int iy = 0; iy < 8; ++iy) {
#pragma unroll
        for (int j = 0; j < W; ++j) {
          const unsigned long long v = *reinterpret_cast<const unsigned long long *>(wbase + (size_t)(iy * S::RS + j * C) * 4);
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            const int kh = iy - r;
            if (kh >= 0 && kh < 7) {
#pragma unroll
              for (int ox = 0; ox < P; ++ox) {
                const int kw = j - ox;
                if (kw >= 0 && kw < 7) dw2[kh * 7 + kw] = fma2(d[r][ox], v, dw2[kh * 7 + kw]);
              }
            }
          }
        }
      }
    }
    __syncthreads();                               // both buffers are free again
    const int pn = pu + gridDim.x;
    if (pn < units) load(pn);
    cp_async_commit();
  }
  cp_async_wait_all();
  __syncthreads();
  // reduce the strips of each channel in shared memory, then one atomic per (tap, channel) per CTA
  float *red = smem;  // [50][C]
  for (int k = threadIdx.x; k < 50 * C; k += S::NT) red[k] = 0.f;
  __syncthreads();
  if (active) {
#pragma unroll
    for (int k = 0; k < 49; ++k) {
      float a, b;
      unpack2(dw2[k], a, b);
      atomicAdd(&red[k * C + 2 * cp], a);
      atomicAdd(&red[k * C + 2 * cp + 1], b);
    }
    float a, b;
    unpack2(db2, a, b);
    atomicAdd(&red[49 * C + 2 * cp], a);
    atomicAdd(&red[49 * C + 2 * cp + 1], b);
  }
  __syncthreads();
  for (int k = threadIdx.x; k < 49 * C; k += S::NT) {
    const int tap = k / C, cc = k - tap * C;
    const int kh = tap / 7, kw = tap - kh * 7;
    atomicAdd(&p.dw[kh * p.w_skh + kw * p.w_skw + cc * p.w_sc], red[k]);
  }
  if (p.dbias)
    for (int cc = threadIdx.x; cc < C; cc += S::NT) atomicAdd(&p.dbias[cc], red[49 * C + cc]);
}

inline int ctas_per_sm(size_t smem, int threads) {
  int n = (int)((227 * 1024) / (smem + 1024));
  if (n > 2048 / threads) n = 2048 / threads;
  return n < 1 ? 1 : n;
}

template <int P, int C>
inline cudaError_t launch_patch_v2(const DwTiledArgs &t, cudaStream_t st) {
  using S = Shape<P, C>;
  constexpr size_t sm = ((size_t)S::NBUF * S::WIN + (size_t)S::NOUT * C) * sizeof(float);
  static_assert(sm <= 226 * 1024, "window + output tile must fit");
  const int np = ln_parts(C), f4 = (C >> 2) / np;
  if (t.a.do_ln && (f4 < 1 || f4 > 6)) return cudaErrorInvalidConfiguration;
  static bool configured = false;
  static int per_sm = 1;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(dwconv_patch_v2_kernel<P, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    if (e != cudaSuccess) return e;
    (void)cudaFuncSetAttribute(dwconv_patch_v2_kernel<P, C>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    int v = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, dwconv_patch_v2_kernel<P, C>, S::NT, sm) != cudaSuccess || v < 1) {
      (void)cudaGetLastError();
      v = ctas_per_sm(sm, S::NT);
    }
    per_sm = v;
    configured = true;
  }
  const int units = t.a.geo.B * t.a.geo.V;
  int grid = 148 * per_sm;
  if (grid > units) grid = units;
  pdl(dwconv_patch_v2_kernel<P, C>, grid, S::NT, sm, st)(t, units);
  return cudaGetLastError();
}

template <int P, int C>
inline cudaError_t launch_patch_wgrad_v2(const DwWgradArgs &p, const int *vis_patch, cudaStream_t st) {
  using S = Shape<P, C>;
  constexpr size_t sm = ((size_t)S::WIN + (size_t)S::NOUT * C) * sizeof(float);
  static_assert(sm >= (size_t)50 * C * sizeof(float), "reduction buffer");
  static bool configured = false;
  static int per_sm = 1;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(dwconv_patch_wgrad_v2_kernel<P, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    if (e != cudaSuccess) return e;
    (void)cudaFuncSetAttribute(dwconv_patch_wgrad_v2_kernel<P, C>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    int v = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, dwconv_patch_wgrad_v2_kernel<P, C>, S::NT, sm) != cudaSuccess || v < 1) {
      (void)cudaGetLastError();
      v = ctas_per_sm(sm, S::NT);
    }
    per_sm = v;
    configured = true;
  }
  const int units = p.geo.B * p.geo.V;
  int grid = 148 * per_sm;
  if (grid > units) grid = units;
  pdl(dwconv_patch_wgrad_v2_kernel<P, C>, grid, S::NT, sm, st)(p, vis_patch, units);
  return cudaGetLastError();
}

}  // namespace dw2

// Dispatch for the instantiated (P, C) pairs; cudaErrorInvalidConfiguration = not taken (caller falls back)
inline cudaError_t launch_dwconv_v2(const DwArgs &a, const int *vis_patch, cudaStream_t st) {
  if (!vis_patch || !a.slot_of) return cudaErrorInvalidConfiguration;
  if (a.do_ln && (a.resid || a.colsum_out)) return cudaErrorInvalidConfiguration;
  if (a.w_sc != 1) return cudaErrorInvalidConfiguration;   // channel pairs are read as adjacent floats
  DwTiledArgs t{a, vis_patch};
  if (a.P == 8 && a.C == 40) return dw2::launch_patch_v2<8, 40>(t, st);
  if (a.P == 8 && a.C == 96) return dw2::launch_patch_v2<8, 96>(t, st);
  if (a.P == 4 && a.C == 80) return dw2::launch_patch_v2<4, 80>(t, st);
  if (a.P == 4 && a.C == 192) return dw2::launch_patch_v2<4, 192>(t, st);
  return cudaErrorInvalidConfiguration;
}
inline cudaError_t launch_dwconv_wgrad_v2(const DwWgradArgs &p, const int *vis_patch, cudaStream_t st) {
  if (!vis_patch || !p.slot_of || p.w_sc != 1) return cudaErrorInvalidConfiguration;
  if (p.P == 8 && p.C == 40) return dw2::launch_patch_wgrad_v2<8, 40>(p, vis_patch, st);
  if (p.P == 8 && p.C == 96) return dw2::launch_patch_wgrad_v2<8, 96>(p, vis_patch, st);
  if (p.P == 4 && p.C == 80) return dw2::launch_patch_wgrad_v2<4, 80>(p, vis_patch, st);
  if (p.P == 4 && p.C == 192) return dw2::launch_patch_wgrad_v2<4, 192>(p, vis_patch, st);
  return cudaErrorInvalidConfiguration;
}

}  // namespace mpmae
