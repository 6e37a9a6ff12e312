// Patch embedding of the sparse encoder (models/convnextv2_sparse.py:113-130):
//   initial_conv = MinkowskiConvolution 3x3 (in_chans -> C0) + LayerNorm + GELU
//   stem         = MinkowskiDepthwiseConvolution k = s = patch/8 + LayerNorm
// evaluated on visible pixels only, reading the NCHW input directly (to_sparse is folded in:
// MinkowskiOps.py:308-317).  The reference runs 9 gather-GEMM-scatter launches with atomics
// (MinkowskiEngine/src/convolution_kernel.cu:114-180) plus a coordinate-map build.
#pragma once
#include "common.cuh"

namespace mpmae {

struct InitConvArgs {
  const float *img;      // [B, Cin, S, S]
  const float *kernel;   // [9, Cin, C0], k = kh + 3*kw
  const float *bias;     // [C0]
  float *chat;           // out [Rpre, C0] normalised conv output
  float *rstd;           // out [Rpre]
  const int *slot_of;    // [B*L]
  const int *vis_patch;  // [B*V]
  int32_t *flags;
  Geo geo;
  int S, Cin, C0, Ppre;  // Ppre = patch_size
  float eps;
};

// stage the (8+2)^2 x Cin input window of an 8x8 pixel tile; masked patches and the image border read 0
__device__ __forceinline__ void load_input_window(const InitConvArgs &p, int n, int gy0, int gx0, float *xin) {
  for (int i = threadIdx.x; i < p.Cin * 100; i += blockDim.x) {
    const int ci = i / 100, w = i - ci * 100;
    const int wy = w / 10, wx = w - wy * 10;
    const int gy = gy0 + wy - 1, gx = gx0 + wx - 1;
    float v = 0.f;
    if (gy >= 0 && gx >= 0 && gy < p.S && gx < p.S) {
      const int l = (gy / p.Ppre) * p.geo.G + gx / p.Ppre;
      if (p.slot_of[n * p.geo.L + l] >= 0) {
        v = p.img[(((int64_t)n * p.Cin + ci) * p.S + gy) * p.S + gx];
        if (!isfinite(v)) v = 0.f;   // torch.nan_to_num(nan=0, posinf=0, neginf=0) on the input, models/fcmae.py:445-449
      }
    }
    xin[i] = v;
  }
}

struct StemArgs;
// LayerNorm of the 64 conv outputs of a tile ((row, part) mapping of common.cuh) and, when `fuse` is set (stem k = s = 1:
// patch 8), the whole stem on top of it: LN0 affine -> GELU -> per-channel scale + bias -> LN1 -> shat, x0.
// vec = [ln0_w | ln0_b | kernel | bias | ln1_w | ln1_b] (6 x C0 floats in shared memory)
template <int F4>
__device__ __forceinline__ void embed_ln_phase(const float *cbuf, int pitch, int C0, int np, float eps, int64_t row0,
                                               float *chat, float *rstd_c, bool fuse, const float *vec, float *shat,
                                               float *rstd_s, float *x0) {
  const int total = 64 * np;
  for (int base = 0; base < total; base += (int)blockDim.x) {
    const int item = base + (int)threadIdx.x;
    const bool ok = item < total;
    const int px = ok ? item / np : 0, part = ok ? item - px * np : 0;
    const float4 *cr = reinterpret_cast<const float4 *>(cbuf + (size_t)px * pitch) + part;
    float4 v[F4];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < F4; ++j) {
      v[j] = ok ? cr[j * np] : make_float4(0.f, 0.f, 0.f, 0.f);
      s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
    }
    const float mean = group_sum(s, np) / (float)C0;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < F4; ++j) {
      v[j].x -= mean; v[j].y -= mean; v[j].z -= mean; v[j].w -= mean;
      q += v[j].x * v[j].x + v[j].y * v[j].y + v[j].z * v[j].z + v[j].w * v[j].w;
    }
    const float rstd = rsqrtf(group_sum(q, np) / (float)C0 + eps);
    const int64_t row = row0 + px;
    if (ok) {
      float4 *o = reinterpret_cast<float4 *>(chat + row * C0) + part;
#pragma unroll
      for (int j = 0; j < F4; ++j) {
        v[j].x *= rstd; v[j].y *= rstd; v[j].z *= rstd; v[j].w *= rstd;
        o[j * np] = v[j];
      }
      if (part == 0) rstd_c[row] = rstd;
    }
    if (!fuse) continue;
    float s1 = 0.f;
#pragma unroll
    for (int j = 0; j < F4; ++j) {
      const int c = (part + j * np) * 4;
      const float4 w0 = *reinterpret_cast<const float4 *>(vec + c), b0 = *reinterpret_cast<const float4 *>(vec + C0 + c);
      const float4 kk = *reinterpret_cast<const float4 *>(vec + 2 * C0 + c), bb = *reinterpret_cast<const float4 *>(vec + 3 * C0 + c);
      v[j].x = fmaf(gelu_f(fmaf(v[j].x, w0.x, b0.x)), kk.x, bb.x);
      v[j].y = fmaf(gelu_f(fmaf(v[j].y, w0.y, b0.y)), kk.y, bb.y);
      v[j].z = fmaf(gelu_f(fmaf(v[j].z, w0.z, b0.z)), kk.z, bb.z);
      v[j].w = fmaf(gelu_f(fmaf(v[j].w, w0.w, b0.w)), kk.w, bb.w);
      s1 += (v[j].x + v[j].y) + (v[j].z + v[j].w);
    }
    const float mean1 = group_sum(s1, np) / (float)C0;
    float q1 = 0.f;
#pragma unroll
    for (int j = 0; j < F4; ++j) {
      v[j].x -= mean1; v[j].y -= mean1; v[j].z -= mean1; v[j].w -= mean1;
      q1 += v[j].x * v[j].x + v[j].y * v[j].y + v[j].z * v[j].z + v[j].w * v[j].w;
    }
    const float rstd1 = rsqrtf(group_sum(q1, np) / (float)C0 + eps);
    if (ok) {
      float4 *os = reinterpret_cast<float4 *>(shat + row * C0) + part;
      float4 *ox = reinterpret_cast<float4 *>(x0 + row * C0) + part;
#pragma unroll
      for (int j = 0; j < F4; ++j) {
        const int c = (part + j * np) * 4;
        const float4 w1 = *reinterpret_cast<const float4 *>(vec + 4 * C0 + c), b1 = *reinterpret_cast<const float4 *>(vec + 5 * C0 + c);
        const float4 nh = make_float4(v[j].x * rstd1, v[j].y * rstd1, v[j].z * rstd1, v[j].w * rstd1);
        os[j * np] = nh;
        ox[j * np] = make_float4(fmaf(nh.x, w1.x, b1.x), fmaf(nh.y, w1.y, b1.y), fmaf(nh.z, w1.z, b1.z), fmaf(nh.w, w1.w, b1.w));
      }
      if (part == 0) rstd_s[row] = rstd1;
    }
  }
}

// Persistent CTAs over the 8x8 pixel tiles: the 3x3 kernel is staged in shared memory once per CTA.
// fuse_vec != null: [ln0_w | ln0_b | stem kernel | stem bias | ln1_w | ln1_b] and the stem outputs (k = s = 1 only).
__global__ void __launch_bounds__(512) initial_conv_fwd_kernel(InitConvArgs p, int64_t units, const float *ln0_w,
                                                               const float *ln0_b, const float *st_k, const float *st_b,
                                                               const float *ln1_w, const float *ln1_b, float *shat,
                                                               float *rstd_s, float *x0, int fuse) { pdl_prologue();
  extern __shared__ __align__(16) float smem[];
  const int C0 = p.C0, Cin = p.Cin, pitch = C0 + 4;
  float *xin = smem;                       // [Cin][10][10]
  float *wsm = xin + Cin * 100;            // [9*Cin][C0]
  float *cbuf = wsm + 9 * Cin * C0;        // [64][C0+4]
  float *vec = cbuf + 64 * pitch;          // [6][C0]
  const int tiles = (p.Ppre / 8) * (p.Ppre / 8);
  for (int i = threadIdx.x; i < 9 * Cin * C0; i += blockDim.x) wsm[i] = p.kernel[i];
  if (fuse)
    for (int i = threadIdx.x; i < C0; i += blockDim.x) {
      vec[i] = ln0_w[i]; vec[C0 + i] = ln0_b[i]; vec[2 * C0 + i] = st_k[i]; vec[3 * C0 + i] = st_b[i];
      vec[4 * C0 + i] = ln1_w[i]; vec[5 * C0 + i] = ln1_b[i];
    }
  const int np = ln_parts(C0), f4 = (C0 >> 2) / np;
  // thread = (vertical pixel pair, 8 output channels); blockDim.x = 32 * (C0 / 8).  Per (ci, kw) four window loads
  // serve 3 taps x 2 pixels and six 16-byte weight loads (warp-broadcast) feed 48 FMAs.
  const int pp = threadIdx.x & 31, q = threadIdx.x >> 5;
  int iy, ix;
  morton_decode(2 * pp, iy, ix);          // pixels 2pp and 2pp+1 are (iy, ix) and (iy+1, ix)
  for (int64_t u = blockIdx.x; u < units; u += gridDim.x) {
    const int pu = (int)(u / tiles), tile = (int)(u - (int64_t)pu * tiles);  // pu = n*V + slot
    const int n = pu / p.geo.V;
    const int l = p.vis_patch[pu];
    int ty, tx;
    morton_decode(tile, ty, tx);
    const int gy0 = (l / p.geo.G) * p.Ppre + ty * 8, gx0 = (l % p.geo.G) * p.Ppre + tx * 8;
    __syncthreads();                       // previous tile's cbuf / xin readers are done (and wsm / vec are staged)
    load_input_window(p, n, gy0, gx0, xin);
    __syncthreads();
    if (q == 0) {
      float s0 = 0.f, s1 = 0.f;
      for (int ci = 0; ci < Cin; ++ci) {
        s0 += fabsf(xin[ci * 100 + (iy + 1) * 10 + ix + 1]);
        s1 += fabsf(xin[ci * 100 + (iy + 2) * 10 + ix + 1]);
      }
      if (s0 == 0.f) atomicAdd(&p.flags[0], 1);
      if (s1 == 0.f) atomicAdd(&p.flags[0], 1);
    }
    {
      float acc[2][8];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[0][j] = acc[1][j] = p.bias[q * 8 + j];
      for (int kw = 0; kw < 3; ++kw)
        for (int ci = 0; ci < Cin; ++ci) {
          const float *xc = xin + ci * 100 + iy * 10 + ix + kw;
          const float xr[4] = {xc[0], xc[10], xc[20], xc[30]};
#pragma unroll
          for (int kh = 0; kh < 3; ++kh) {
            const float *wk = wsm + (size_t)((kh + 3 * kw) * Cin + ci) * C0 + q * 8;
            const float4 w0 = *reinterpret_cast<const float4 *>(wk);
            const float4 w1 = *reinterpret_cast<const float4 *>(wk + 4);
            const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              acc[0][j] = fmaf(xr[kh], wv[j], acc[0][j]);
              acc[1][j] = fmaf(xr[kh + 1], wv[j], acc[1][j]);
            }
          }
        }
      float4 *c0p = reinterpret_cast<float4 *>(cbuf + (size_t)(2 * pp) * pitch + q * 8);
      float4 *c1p = reinterpret_cast<float4 *>(cbuf + (size_t)(2 * pp + 1) * pitch + q * 8);
      c0p[0] = make_float4(acc[0][0], acc[0][1], acc[0][2], acc[0][3]); c0p[1] = make_float4(acc[0][4], acc[0][5], acc[0][6], acc[0][7]);
      c1p[0] = make_float4(acc[1][0], acc[1][1], acc[1][2], acc[1][3]); c1p[1] = make_float4(acc[1][4], acc[1][5], acc[1][6], acc[1][7]);
    }
    __syncthreads();
    const int64_t row0 = (int64_t)pu * p.Ppre * p.Ppre + tile * 64;
    switch (f4) {
      case 1: embed_ln_phase<1>(cbuf, pitch, C0, np, p.eps, row0, p.chat, p.rstd, fuse != 0, vec, shat, rstd_s, x0); break;
      case 2: embed_ln_phase<2>(cbuf, pitch, C0, np, p.eps, row0, p.chat, p.rstd, fuse != 0, vec, shat, rstd_s, x0); break;
      case 3: embed_ln_phase<3>(cbuf, pitch, C0, np, p.eps, row0, p.chat, p.rstd, fuse != 0, vec, shat, rstd_s, x0); break;
      case 4: embed_ln_phase<4>(cbuf, pitch, C0, np, p.eps, row0, p.chat, p.rstd, fuse != 0, vec, shat, rstd_s, x0); break;
      case 5: embed_ln_phase<5>(cbuf, pitch, C0, np, p.eps, row0, p.chat, p.rstd, fuse != 0, vec, shat, rstd_s, x0); break;
      default: embed_ln_phase<6>(cbuf, pitch, C0, np, p.eps, row0, p.chat, p.rstd, fuse != 0, vec, shat, rstd_s, x0); break;
    }
  }
}

// dW[k, ci, co] += sum_pix x[pix + off(k), ci] * dc[pix, co] ; db[co] += sum_pix dc[pix, co]
struct InitConvWgradArgs {
  InitConvArgs f;      // geometry + img + tables (kernel/bias/chat unused)
  const float *dc;     // [Rpre, C0]
  float *dkernel;      // [9, Cin, C0]
  float *dbias;        // [C0]
};
// Persistent CTAs; a thread owns dW[all 9 taps][one ci][4 consecutive co] (36 accumulators that live in registers for
// the whole kernel) for one half of each tile's pixels: per pixel 9 window loads + one 16-byte dc load feed 36 FMAs.
// Threads: Cin * (C0 / 4) owners x `parts` pixel groups (blockDim.x = owners * parts, parts in {1, 2, 4}).  One flush
// (shared-memory atomics, then global) at the end.
__global__ void __launch_bounds__(512) initial_conv_wgrad_kernel(InitConvWgradArgs a) { pdl_prologue();
  extern __shared__ __align__(16) float smem[];
  const InitConvArgs &p = a.f;
  const int C0 = p.C0, Cin = p.Cin;
  float *xin = smem;                     // [Cin][100]
  float *dcs = xin + Cin * 100;          // [64][C0]   (pixel index = Z-order)
  float *dws = dcs + 64 * C0;            // [9*Cin + 1][C0]  (last row = bias)
  const int nacc = (9 * Cin + 1) * C0;
  for (int i = threadIdx.x; i < nacc; i += blockDim.x) dws[i] = 0.f;
  const int cog = C0 >> 2, owners = Cin * cog;
  const int parts = blockDim.x / owners, ppp = 64 / parts;   // pixels per part
  const int half = threadIdx.x / owners, own = threadIdx.x - half * owners;
  const bool active = half < parts;
  const int ci = own / cog, co = (own - ci * cog) * 4;
  float acc[9][4];
  float accb[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int k = 0; k < 9; ++k)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[k][j] = 0.f;
  const int tiles = (p.Ppre / 8) * (p.Ppre / 8);
  const int64_t units = (int64_t)p.geo.B * p.geo.V * tiles;
  for (int64_t u = blockIdx.x; u < units; u += gridDim.x) {
    const int pu = (int)(u / tiles), tile = (int)(u - (int64_t)pu * tiles);
    const int n = pu / p.geo.V;
    const int l = p.vis_patch[pu];
    int ty, tx;
    morton_decode(tile, ty, tx);
    const int gy0 = (l / p.geo.G) * p.Ppre + ty * 8, gx0 = (l % p.geo.G) * p.Ppre + tx * 8;
    __syncthreads();
    load_input_window(p, n, gy0, gx0, xin);
    const int64_t row0 = (int64_t)pu * p.Ppre * p.Ppre + tile * 64;
    {
      const float4 *src = reinterpret_cast<const float4 *>(a.dc + row0 * C0);
      for (int i = threadIdx.x; i < 16 * C0; i += blockDim.x) reinterpret_cast<float4 *>(dcs)[i] = __ldg(src + i);
    }
    __syncthreads();
    if (active) {
      const float *xc = xin + ci * 100;
#pragma unroll 4
      for (int pp = 0; pp < ppp; ++pp) {
        const int px = half * ppp + pp;
        // Z-order decode of the pixel (row bit is the LSB of every pair)
        const int iy = (px & 1) | ((px >> 1) & 2) | ((px >> 2) & 4), ix = ((px >> 1) & 1) | ((px >> 2) & 2) | ((px >> 3) & 4);
        const float4 d = *reinterpret_cast<const float4 *>(dcs + px * C0 + co);
        const float *xw = xc + iy * 10 + ix;
#pragma unroll
        for (int k = 0; k < 9; ++k) {
          const float xv = xw[(k % 3) * 10 + k / 3];   // k = kh + 3*kw
          acc[k][0] = fmaf(xv, d.x, acc[k][0]); acc[k][1] = fmaf(xv, d.y, acc[k][1]);
          acc[k][2] = fmaf(xv, d.z, acc[k][2]); acc[k][3] = fmaf(xv, d.w, acc[k][3]);
        }
        if (ci == 0) { accb[0] += d.x; accb[1] += d.y; accb[2] += d.z; accb[3] += d.w; }
      }
    }
  }
  __syncthreads();
  if (active) {
#pragma unroll
    for (int k = 0; k < 9; ++k)
#pragma unroll
      for (int j = 0; j < 4; ++j) atomicAdd(&dws[(k * Cin + ci) * C0 + co + j], acc[k][j]);
    if (ci == 0)
#pragma unroll
      for (int j = 0; j < 4; ++j) atomicAdd(&dws[9 * Cin * C0 + co + j], accb[j]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 9 * Cin * C0; i += blockDim.x) atomicAdd(&a.dkernel[i], dws[i]);
  for (int c = threadIdx.x; c < C0; c += blockDim.x) atomicAdd(&a.dbias[c], dws[9 * Cin * C0 + c]);
}

// ------------------------------------------------------------------------------------------------
struct StemArgs {
  const float *chat;   // [Rpre, C0]
  const float *rstd_c; // [Rpre]
  const float *ln0_w, *ln0_b;  // initial_conv.1.ln
  const float *kernel; // [s*s, C0]
  const float *bias;   // [C0]
  const float *ln1_w, *ln1_b;  // stem.1.ln
  float *shat;         // [R0, C0]
  float *rstd_s;       // [R0]
  float *x0;           // [R0, C0]
  int64_t R0;
  int C0, s2;          // s2 = s*s children per output pixel
  float eps;
};
constexpr int kStemMaxPerLane = 4;  // C0 <= 128

// NQ = ceil(C0 / 32) channel rounds per lane (compile time: no dead rounds)
template <int NQ>
__global__ void stem_fwd_kernel(StemArgs p) { pdl_prologue();
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= p.R0) return;
  const int lane = threadIdx.x & 31, C0 = p.C0;
  float sv[NQ];
  float sum = 0.f;
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    const int c = lane + 32 * q;
    sv[q] = 0.f;
    if (c < C0) {
      float acc = p.bias[c];
      for (int j = 0; j < p.s2; ++j) {
        const float t = p.chat[(r * p.s2 + j) * C0 + c] * p.ln0_w[c] + p.ln0_b[c];
        acc = fmaf(gelu_f(t), p.kernel[j * C0 + c], acc);
      }
      sv[q] = acc; sum += acc;
    }
  }
  const float mean = warp_sum(sum) / (float)C0;
  float var = 0.f;
#pragma unroll
  for (int q = 0; q < NQ; ++q)
    if (lane + 32 * q < C0) { const float d = sv[q] - mean; var += d * d; }
  const float rstd = rsqrtf(warp_sum(var) / (float)C0 + p.eps);
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    const int c = lane + 32 * q;
    if (c < C0) {
      const float nh = (sv[q] - mean) * rstd;
      p.shat[r * C0 + c] = nh;
      p.x0[r * C0 + c] = nh * p.ln1_w[c] + p.ln1_b[c];
    }
  }
  if (lane == 0) p.rstd_s[r] = rstd;
}

struct StemBwdArgs {
  StemArgs f;
  const float *dx0;   // [R0, C0]
  float *dc;          // out [Rpre, C0] gradient at the initial conv output (pre-LN)
  float *d_ln0_w, *d_ln0_b, *d_kernel, *d_bias, *d_ln1_w, *d_ln1_b;
};
// per-lane partial column sums: [0]=d_ln1_w [1]=d_ln1_b [2]=d_bias [3]=d_ln0_w [4]=d_ln0_b [5..5+s2)=d_kernel[j]
template <int NQ>
__global__ void __launch_bounds__(256) stem_bwd_kernel(StemBwdArgs a) { pdl_prologue();
  extern __shared__ float red[];  // [(5 + s2)][C0]
  const StemArgs &p = a.f;
  const int C0 = p.C0, s2 = p.s2, nvec = 5 + s2;
  for (int i = threadIdx.x; i < nvec * C0; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  float part[9][NQ];
#pragma unroll
  for (int v = 0; v < 9; ++v)
#pragma unroll
    for (int q = 0; q < NQ; ++q) part[v][q] = 0.f;

  for (int64_t r = (int64_t)blockIdx.x * nw + warp; r < p.R0; r += (int64_t)gridDim.x * nw) {
    // LN1 backward
    float dsh[NQ], nh[NQ];
    float s1 = 0.f, s2s = 0.f;
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const int c = lane + 32 * q;
      dsh[q] = 0.f; nh[q] = 0.f;
      if (c < C0) {
        const float g = a.dx0[r * C0 + c];
        nh[q] = p.shat[r * C0 + c];
        part[0][q] += g * nh[q];
        part[1][q] += g;
        dsh[q] = g * p.ln1_w[c];
        s1 += dsh[q]; s2s += dsh[q] * nh[q];
      }
    }
    s1 = warp_sum(s1) / (float)C0; s2s = warp_sum(s2s) / (float)C0;
    const float rs = p.rstd_s[r];
    float ds[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      ds[q] = (lane + 32 * q < C0) ? rs * (dsh[q] - s1 - nh[q] * s2s) : 0.f;
      part[2][q] += ds[q];
    }
    // children
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (j >= s2) break;
      const int64_t rc = r * s2 + j;
      float dch[NQ], ch[NQ];
      float t1 = 0.f, t2 = 0.f;
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const int c = lane + 32 * q;
        dch[q] = 0.f; ch[q] = 0.f;
        if (c < C0) {
          ch[q] = p.chat[rc * C0 + c];
          const float t = ch[q] * p.ln0_w[c] + p.ln0_b[c];
          float gj, dgj;
          gelu_both_f(t, gj, dgj);
          part[5 + j][q] += ds[q] * gj;  // d_kernel[j]; s2 <= 4 (patch 8 or 16)
          const float dt = ds[q] * p.kernel[j * C0 + c] * dgj;
          part[3][q] += dt * ch[q];
          part[4][q] += dt;
          dch[q] = dt * p.ln0_w[c];
          t1 += dch[q]; t2 += dch[q] * ch[q];
        }
      }
      t1 = warp_sum(t1) / (float)C0; t2 = warp_sum(t2) / (float)C0;
      const float rc_std = p.rstd_c[rc];
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const int c = lane + 32 * q;
        if (c < C0) a.dc[rc * C0 + c] = rc_std * (dch[q] - t1 - ch[q] * t2);
      }
    }
  }
#pragma unroll
  for (int v = 0; v < 9; ++v)
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const int c = lane + 32 * q;
      if (v < nvec && c < C0) atomicAdd(&red[v * C0 + c], part[v][q]);
    }
  __syncthreads();
  float *dst[5] = {a.d_ln1_w, a.d_ln1_b, a.d_bias, a.d_ln0_w, a.d_ln0_b};
  for (int i = threadIdx.x; i < nvec * C0; i += blockDim.x) {
    const int v = i / C0, c = i - v * C0;
    if (v < 5) atomicAdd(&dst[v][c], red[i]);
    else atomicAdd(&a.d_kernel[(v - 5) * C0 + c], red[i]);
  }
}

// The same backward with the (row, part) mapping of the LayerNorm kernels: LPR adjacent lanes (a power of two >= C0 / 4)
// share one row and every lane owns one float4 column, so a warp handles 32 / LPR rows per iteration with 16-byte loads and
// log2(LPR)-step row reductions.  The warp-per-row kernel above leaves 24 of 64 lane slots idle at C0 = 40 and issues four
// times as many instructions (226 us of the 7.8 ms cfg2 step in round 1).  C0 % 4 == 0, C0 <= 128, s2 <= 4.
template <int LPR>
__global__ void __launch_bounds__(256) stem_bwd_vec_kernel(StemBwdArgs a) { pdl_prologue();
  extern __shared__ float red[];  // [(5 + s2)][C0]
  const StemArgs &p = a.f;
  const int C0 = p.C0, s2 = p.s2, nvec = 5 + s2;
  for (int i = threadIdx.x; i < nvec * C0; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
  constexpr int RPW = 32 / LPR;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int sub = lane % LPR, rsub = lane / LPR, c = sub * 4;
  const bool colok = c < C0;
  const float inv_c = 1.f / (float)C0;
  auto ld4 = [](const float *ptr) { return __ldg(reinterpret_cast<const float4 *>(ptr)); };
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 w1 = colok ? ld4(p.ln1_w + c) : z4, w0 = colok ? ld4(p.ln0_w + c) : z4, b0 = colok ? ld4(p.ln0_b + c) : z4;
  float4 kj[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) kj[j] = (colok && j < s2) ? ld4(p.kernel + j * C0 + c) : z4;
  float4 part[9];
#pragma unroll
  for (int v = 0; v < 9; ++v) part[v] = z4;
  auto fma4 = [](float4 &acc, const float4 &x, const float4 &y) {
    acc.x = fmaf(x.x, y.x, acc.x); acc.y = fmaf(x.y, y.y, acc.y); acc.z = fmaf(x.z, y.z, acc.z); acc.w = fmaf(x.w, y.w, acc.w);
  };
  auto add4 = [](float4 &acc, const float4 &x) { acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w; };
  auto sum4 = [](const float4 &x) { return (x.x + x.y) + (x.z + x.w); };
  auto dot4 = [](const float4 &x, const float4 &y) { return (x.x * y.x + x.y * y.y) + (x.z * y.z + x.w * y.w); };
  // two row groups per iteration, every global load issued before the first reduction: the loop is bound by the latency of
  // its dependent load -> shuffle -> store chain, so the loads of both groups (and of all children) are put in flight at once
  const int64_t stride = (int64_t)gridDim.x * nw;
  for (int64_t rg0 = (int64_t)blockIdx.x * nw + warp; rg0 * RPW < p.R0; rg0 += 2 * stride) {
    int64_t r[2];
    bool ok[2];
    float4 g[2], nh[2], ch[2][4];
    float rs[2], rcs[2][4];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      r[u] = (rg0 + u * stride) * RPW + rsub;
      ok[u] = colok && r[u] < p.R0;
      g[u] = ok[u] ? ld4(a.dx0 + r[u] * C0 + c) : z4;
      nh[u] = ok[u] ? ld4(p.shat + r[u] * C0 + c) : z4;
      rs[u] = ok[u] ? __ldg(p.rstd_s + r[u]) : 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const bool okj = ok[u] && j < s2;
        ch[u][j] = okj ? ld4(p.chat + (r[u] * s2 + j) * C0 + c) : z4;
        rcs[u][j] = okj ? __ldg(p.rstd_c + r[u] * s2 + j) : 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      // LN1 backward
      fma4(part[0], g[u], nh[u]);
      add4(part[1], g[u]);
      const float4 dsh = make_float4(g[u].x * w1.x, g[u].y * w1.y, g[u].z * w1.z, g[u].w * w1.w);
      const float s1 = group_sum(sum4(dsh), LPR) * inv_c, s2s = group_sum(dot4(dsh, nh[u]), LPR) * inv_c;
      const float4 ds = make_float4(rs[u] * (dsh.x - s1 - nh[u].x * s2s), rs[u] * (dsh.y - s1 - nh[u].y * s2s),
                                    rs[u] * (dsh.z - s1 - nh[u].z * s2s), rs[u] * (dsh.w - s1 - nh[u].w * s2s));
      add4(part[2], ds);
      // children
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (j >= s2) break;
        const float4 cj = ch[u][j];
        float4 gj, dgj;
        gelu_both4_f(make_float4(fmaf(cj.x, w0.x, b0.x), fmaf(cj.y, w0.y, b0.y), fmaf(cj.z, w0.z, b0.z), fmaf(cj.w, w0.w, b0.w)), gj, dgj);
        fma4(part[5 + j], ds, gj);                       // d_kernel[j]
        const float4 dt = make_float4(ds.x * kj[j].x * dgj.x, ds.y * kj[j].y * dgj.y, ds.z * kj[j].z * dgj.z, ds.w * kj[j].w * dgj.w);
        fma4(part[3], dt, cj);
        add4(part[4], dt);
        const float4 dch = make_float4(dt.x * w0.x, dt.y * w0.y, dt.z * w0.z, dt.w * w0.w);
        const float t1 = group_sum(sum4(dch), LPR) * inv_c, t2 = group_sum(dot4(dch, cj), LPR) * inv_c;
        if (ok[u]) {
          const float rc_std = rcs[u][j];
          *reinterpret_cast<float4 *>(a.dc + (r[u] * s2 + j) * C0 + c) =
              make_float4(rc_std * (dch.x - t1 - cj.x * t2), rc_std * (dch.y - t1 - cj.y * t2), rc_std * (dch.z - t1 - cj.z * t2),
                          rc_std * (dch.w - t1 - cj.w * t2));
        }
      }
    }
  }
  if (colok) {
#pragma unroll
    for (int v = 0; v < 9; ++v)
      if (v < nvec) {
        atomicAdd(&red[v * C0 + c], part[v].x); atomicAdd(&red[v * C0 + c + 1], part[v].y);
        atomicAdd(&red[v * C0 + c + 2], part[v].z); atomicAdd(&red[v * C0 + c + 3], part[v].w);
      }
  }
  __syncthreads();
  float *dst[5] = {a.d_ln1_w, a.d_ln1_b, a.d_bias, a.d_ln0_w, a.d_ln0_b};
  for (int i = threadIdx.x; i < nvec * C0; i += blockDim.x) {
    const int v = i / C0, cc = i - v * C0;
    if (v < 5) atomicAdd(&dst[v][cc], red[i]);
    else atomicAdd(&a.d_kernel[(v - 5) * C0 + cc], red[i]);
  }
}
inline bool launch_stem_bwd_vec(const StemBwdArgs &sb, cudaStream_t st) {
  const int C0 = sb.f.C0, c4 = C0 / 4;
  if (C0 % 4 != 0 || c4 > 32 || sb.f.s2 > 4) return false;
  if ((((uintptr_t)sb.dx0 | (uintptr_t)sb.dc | (uintptr_t)sb.f.shat | (uintptr_t)sb.f.chat) & 15) != 0) return false;
  const size_t ssm = (size_t)(5 + sb.f.s2) * C0 * 4;
  const int grid = 148 * 8;
  if (c4 <= 8) pdl(stem_bwd_vec_kernel<8>, grid, 256, ssm, st)(sb);
  else if (c4 <= 16) pdl(stem_bwd_vec_kernel<16>, grid, 256, ssm, st)(sb);
  else pdl(stem_bwd_vec_kernel<32>, grid, 256, ssm, st)(sb);
  return true;
}

}  // namespace mpmae
