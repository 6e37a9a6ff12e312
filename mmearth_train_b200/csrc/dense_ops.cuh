// Geometry-agnostic dense operators for the reference's DENSE ConvNeXt-V2 (models/convnextv2.py:59-207: the network that
// finetuning / linear probing runs on a pretrained encoder, hubconf.py:77-93).  Its geometry differs from the masked encoder's
// (un-padded 3x3 stem convolution: 56 -> 54 -> 27 -> 13 -> 6), so the patch-aligned kernels of the pretraining step do not
// apply; activations are plain channels-last rows [B*H*W, C] here and the heavy lifting stays in the tcgen05 GEMMs
// (mpmae_gemm_epi).  Inference only.
#pragma once
#include "common.cuh"

namespace mpmae {

// out[r, ci*k*k + kh*k + kw] = x[n, ci, oy*s + kh, ox*s + kw]   (r = (n*Ho + oy)*Wo + ox; columns >= C*k*k are zero: the GEMM
// wants K % 8 == 0).  The column order is torch's weight.reshape(Cout, -1).  nchw: x is [B, C, H, W], else rows [B, H, W, C].
__global__ void __launch_bounds__(256) dense_im2col_kernel(const float *__restrict__ x, float *__restrict__ out, int B, int C, int H,
                                                          int W, int k, int s, int Ho, int Wo, int Kpad, int nchw) { pdl_prologue();
  const int64_t total = (int64_t)B * Ho * Wo * Kpad;
  const int kk = k * k;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int col = (int)(i % Kpad);
    const int64_t r = i / Kpad;
    float v = 0.f;
    if (col < C * kk) {
      const int ci = col / kk, t = col - ci * kk, kh = t / k, kw = t - kh * k;
      const int ox = (int)(r % Wo), oy = (int)((r / Wo) % Ho), n = (int)(r / ((int64_t)Wo * Ho));
      const int iy = oy * s + kh, ix = ox * s + kw;
      v = nchw ? __ldg(x + (((int64_t)n * C + ci) * H + iy) * W + ix) : __ldg(x + (((int64_t)n * H + iy) * W + ix) * C + ci);
    }
    out[i] = v;
  }
}

// LayerNorm over the channels of every row with affine (models/norm_layers.py:23-31) and optional GELU; one warp per row
__global__ void __launch_bounds__(256) ln_affine_rows_kernel(const float *__restrict__ x, const float *__restrict__ w,
                                                            const float *__restrict__ b, float *__restrict__ out, int64_t R, int C,
                                                            float eps, int gelu) { pdl_prologue();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int64_t r = (int64_t)blockIdx.x * nw + warp; r < R; r += (int64_t)gridDim.x * nw) {
    const float *xr = x + r * C;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += xr[c];
    const float mean = warp_sum(s) / (float)C;
    float v = 0.f;
    for (int c = lane; c < C; c += 32) { const float d = xr[c] - mean; v += d * d; }
    const float rstd = rsqrtf(warp_sum(v) / (float)C + eps);
    for (int c = lane; c < C; c += 32) {
      float y = (xr[c] - mean) * rstd;
      if (w) y = y * __ldg(w + c) + __ldg(b + c);
      out[r * C + c] = gelu ? gelu_f(y) : y;
    }
  }
}

// Depthwise k x k convolution (stride s, zero padding p) on channels-last rows [B, H, W, C] with the torch weight layout
// [C, 1, k, k]; optional LayerNorm without affine over the channels of every output pixel (the block's norm: its affine is
// folded into pwconv1 by the caller).  One warp per output pixel, lanes over channels (C <= 1024).
__global__ void __launch_bounds__(256) dense_dwconv_kernel(const float *__restrict__ x, const float *__restrict__ w,
                                                          const float *__restrict__ bias, float *__restrict__ out, int B, int H,
                                                          int W, int C, int k, int s, int p, int Ho, int Wo, int ln, float eps) { pdl_prologue();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int64_t total = (int64_t)B * Ho * Wo;
  for (int64_t o = (int64_t)blockIdx.x * nw + warp; o < total; o += (int64_t)gridDim.x * nw) {
    const int ox = (int)(o % Wo), oy = (int)((o / Wo) % Ho), n = (int)(o / ((int64_t)Wo * Ho));
    float acc[32];
    float sum = 0.f;
#pragma unroll
    for (int q = 0; q < 32; ++q) {
      const int c = lane + 32 * q;
      acc[q] = 0.f;
      if (c < C) {
        float a = bias ? __ldg(bias + c) : 0.f;
        for (int kh = 0; kh < k; ++kh) {
          const int iy = oy * s - p + kh;
          if (iy < 0 || iy >= H) continue;
          for (int kw = 0; kw < k; ++kw) {
            const int ix = ox * s - p + kw;
            if (ix < 0 || ix >= W) continue;
            a = fmaf(__ldg(x + (((int64_t)n * H + iy) * W + ix) * C + c), __ldg(w + (c * k + kh) * k + kw), a);
          }
        }
        acc[q] = a;
        sum += a;
      }
      if (32 * (q + 1) >= C) break;
    }
    float mean = 0.f, rstd = 1.f;
    if (ln) {
      mean = warp_sum(sum) / (float)C;
      float v = 0.f;
#pragma unroll
      for (int q = 0; q < 32; ++q) {
        if (lane + 32 * q < C) { const float d = acc[q] - mean; v += d * d; }
        if (32 * (q + 1) >= C) break;
      }
      rstd = rsqrtf(warp_sum(v) / (float)C + eps);
    }
#pragma unroll
    for (int q = 0; q < 32; ++q) {
      const int c = lane + 32 * q;
      if (c < C) out[o * C + c] = (acc[q] - mean) * rstd;
      if (32 * (q + 1) >= C) break;
    }
  }
}

}  // namespace mpmae
