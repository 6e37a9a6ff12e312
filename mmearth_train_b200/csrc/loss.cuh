// Fused per-modality reconstruction losses + uncertainty weighting (models/fcmae.py:267-412,
// custom_loss.py:19-30).  The reference spends ~150 small ATen kernels and several boolean-index
// host syncs here; this is one pass over the predictions that also emits the closed-form
// gradient d(loss_i)/d(pred) up to the per-modality scalar 1/denominator_i, which is only known
// after the global reduction and is applied later (folded into the head-gradient GEMMs).
//
//   acc[2*i]   = numerator  of modality i (sum of per-patch / per-pixel / per-element losses)
//   acc[2*i+1] = denominator (count of contributing patches / pixels / elements)
#pragma once
#include "common.cuh"
#include "../../include/mpmae.h"

namespace mpmae {

struct LossMod {
  int kind, chans, col_off, norm_pix;
  const void *target;
};
struct LossArgs {
  LossMod mod[MPMAE_MAX_MOD];
  int n_mod;
  const float *pred_pix; int npix;   // [B*L, npix]
  const float *pred_img; int nimg;   // [B, nimg]
  const float *mask;                 // [B*L]
  float *dpix;                       // [B*L, npix] raw gradient
  float *dimg;                       // [B, nimg]
  float *acc;                        // [2*n_mod]
  int B, L, G, p, S;
};

__device__ __forceinline__ float nan_to_zero(float t) { return isfinite(t) ? t : 0.f; }

// block-wide sum of two values (blockDim.x <= 1024); result valid in all threads
__device__ __forceinline__ void block_sum2(float &a, float &b, float *red) {
  a = warp_sum(a); b = warp_sum(b);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (lane == 0) { red[warp] = a; red[32 + warp] = b; }
  __syncthreads();
  float x = lane < nw ? red[lane] : 0.f, y = lane < nw ? red[32 + lane] : 0.f;
  a = warp_sum(x); b = warp_sum(y);
}

__global__ void __launch_bounds__(256) pixel_loss_kernel(LossArgs a) {
  __shared__ float red[64];
  const int cell = blockIdx.x;               // n*L + l
  const int n = cell / a.L, l = cell - n * a.L;
  const int ph = l / a.G, pw = l - ph * a.G;
  const int p = a.p, p2 = p * p;
  const bool masked = a.mask[cell] != 0.f;
  for (int mi = 0; mi < a.n_mod; ++mi) {
    const LossMod m = a.mod[mi];
    if (m.kind != MPMAE_PIXEL_CONTINUOUS && m.kind != MPMAE_PIXEL_CATEGORICAL) continue;
    const int len = p2 * m.chans;
    const float *pr = a.pred_pix + (int64_t)cell * a.npix + m.col_off;
    float *dp = a.dpix + (int64_t)cell * a.npix + m.col_off;
    if (!masked) {  // visible patch: no loss, zero gradient
      for (int j = threadIdx.x; j < len; j += blockDim.x) dp[j] = 0.f;
      continue;
    }
    if (m.kind == MPMAE_PIXEL_CONTINUOUS) {
      const float *tg = static_cast<const float *>(m.target);
      const int c = m.chans;
      float mean = 0.f, inv_std = 1.f;
      if (m.norm_pix) {  // per-patch (t - mean) / sqrt(var_unbiased + 1e-6), fcmae.py:377-382
        float s = 0.f, dummy = 0.f;
        for (int j = threadIdx.x; j < len; j += blockDim.x) {
          const int ch = j % c, q = j / c, pi = q / p, qi = q - pi * p;
          s += nan_to_zero(tg[(((int64_t)n * c + ch) * a.S + ph * p + pi) * a.S + pw * p + qi]);
        }
        block_sum2(s, dummy, red);
        mean = s / (float)len;
        float v = 0.f;
        dummy = 0.f;
        for (int j = threadIdx.x; j < len; j += blockDim.x) {
          const int ch = j % c, q = j / c, pi = q / p, qi = q - pi * p;
          const float d = nan_to_zero(tg[(((int64_t)n * c + ch) * a.S + ph * p + pi) * a.S + pw * p + qi]) - mean;
          v += d * d;
        }
        block_sum2(v, dummy, red);
        inv_std = rsqrtf(v / (float)(len - 1) + 1.0e-6f);
      }
      float se = 0.f, cnt = 0.f;
      for (int j = threadIdx.x; j < len; j += blockDim.x) {
        const int ch = j % c, q = j / c, pi = q / p, qi = q - pi * p;
        float t = nan_to_zero(tg[(((int64_t)n * c + ch) * a.S + ph * p + pi) * a.S + pw * p + qi]);
        t = (t - mean) * inv_std;
        const float d = pr[j] - t;
        const float e = d * d;
        if (e == e) { se += e; cnt += 1.f; }   // NaN squared errors are dropped (fcmae.py:385-388)
      }
      block_sum2(se, cnt, red);
      const float patch_loss = se / cnt;       // cnt == 0 -> NaN -> dropped below
      const bool counted = (patch_loss == patch_loss) && patch_loss != 0.f;
      if (threadIdx.x == 0 && counted) { atomicAdd(&a.acc[2 * mi], patch_loss); atomicAdd(&a.acc[2 * mi + 1], 1.f); }
      for (int j = threadIdx.x; j < len; j += blockDim.x) {
        const int ch = j % c, q = j / c, pi = q / p, qi = q - pi * p;
        float t = nan_to_zero(tg[(((int64_t)n * c + ch) * a.S + ph * p + pi) * a.S + pw * p + qi]);
        t = (t - mean) * inv_std;
        const float d = pr[j] - t;
        dp[j] = (counted && d == d) ? 2.f * d / cnt : 0.f;
      }
    } else {  // categorical pixels: logits [p2][K], int64 target, -1 ignored (fcmae.py:302-346)
      const long long *tg = static_cast<const long long *>(m.target);
      const int K = m.chans;
      float ls = 0.f, nsel = 0.f;
      for (int q = threadIdx.x; q < p2; q += blockDim.x) {
        const int pi = q / p, qi = q - pi * p;
        const long long t = tg[((int64_t)n * a.S + ph * p + pi) * a.S + pw * p + qi];
        const float *lg = pr + q * K;
        float *dl = dp + q * K;
        if (t < 0 || t >= K) {
          for (int k = 0; k < K; ++k) dl[k] = 0.f;
          continue;
        }
        float mx = lg[0];
        for (int k = 1; k < K; ++k) mx = fmaxf(mx, lg[k]);
        float se = 0.f;
        for (int k = 0; k < K; ++k) se += expf(lg[k] - mx);
        const float lse = mx + logf(se);
        ls += lse - lg[t];
        nsel += 1.f;
        for (int k = 0; k < K; ++k) dl[k] = expf(lg[k] - lse) - (k == (int)t ? 1.f : 0.f);
      }
      block_sum2(ls, nsel, red);
      if (threadIdx.x == 0 && nsel > 0.f) { atomicAdd(&a.acc[2 * mi], ls); atomicAdd(&a.acc[2 * mi + 1], nsel); }
    }
  }
}

// image-level heads: one CTA per sample (fcmae.py:281-301)
__global__ void __launch_bounds__(256) image_loss_kernel(LossArgs a) {
  __shared__ float red[64];
  __shared__ int cls_sh;
  const int n = blockIdx.x;
  for (int mi = 0; mi < a.n_mod; ++mi) {
    const LossMod m = a.mod[mi];
    const int c = m.chans;
    const float *pr = a.pred_img + (int64_t)n * a.nimg + m.col_off;
    float *dp = a.dimg + (int64_t)n * a.nimg + m.col_off;
    if (m.kind == MPMAE_IMAGE_CATEGORICAL) {
      const long long *tg = static_cast<const long long *>(m.target) + (int64_t)n * c;
      // argmax of the one-hot row (first maximum, as torch.argmax)
      if (threadIdx.x == 0) cls_sh = 0x7fffffff;
      long long best = tg[0];
      for (int k = 1; k < c; ++k) best = tg[k] > best ? tg[k] : best;  // small c; every thread scans
      __syncthreads();
      for (int k = threadIdx.x; k < c; k += blockDim.x)
        if (tg[k] == best) atomicMin(&cls_sh, k);
      __syncthreads();
      const int cls = cls_sh;
      float mx = -INFINITY, dummy = 0.f;
      for (int k = threadIdx.x; k < c; k += blockDim.x) mx = fmaxf(mx, pr[k]);
      mx = warp_max(mx);
      __syncthreads();
      if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
      __syncthreads();
      mx = red[0];
      for (int w = 1; w < (blockDim.x >> 5); ++w) mx = fmaxf(mx, red[w]);
      float se = 0.f;
      for (int k = threadIdx.x; k < c; k += blockDim.x) se += expf(pr[k] - mx);
      block_sum2(se, dummy, red);
      const float lse = mx + logf(se);
      if (threadIdx.x == 0) { atomicAdd(&a.acc[2 * mi], lse - pr[cls]); atomicAdd(&a.acc[2 * mi + 1], 1.f); }
      for (int k = threadIdx.x; k < c; k += blockDim.x) dp[k] = expf(pr[k] - lse) - (k == cls ? 1.f : 0.f);
      __syncthreads();
    } else if (m.kind == MPMAE_IMAGE_CONTINUOUS) {
      const float *tg = static_cast<const float *>(m.target) + (int64_t)n * c;
      float se = 0.f, cnt = 0.f;
      for (int k = threadIdx.x; k < c; k += blockDim.x) {
        const float t = tg[k];
        if (t == t) {
          const float d = pr[k] - t;
          se += d * d; cnt += 1.f; dp[k] = 2.f * d;
        } else {
          dp[k] = 0.f;
        }
      }
      block_sum2(se, cnt, red);
      if (threadIdx.x == 0 && cnt > 0.f) { atomicAdd(&a.acc[2 * mi], se); atomicAdd(&a.acc[2 * mi + 1], cnt); }
    }
  }
}

// losses[0..n) = L_i ; losses[n..2n) = weighted_i ; losses[2n] = total
__global__ void loss_finalize_kernel(const float *__restrict__ acc, const float *__restrict__ log_vars, int n_mod,
                                     int uncertainty, float *__restrict__ losses) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  float total = 0.f;
  for (int i = 0; i < n_mod; ++i) {
    const float L = acc[2 * i] / acc[2 * i + 1];
    losses[i] = L;
    float w = L;
    if (uncertainty) {
      const float s = log_vars[i];
      w = (L != 0.f) ? expf(-s) * L + s : 0.f;
    }
    losses[n_mod + i] = w;
    total += w;
  }
  losses[2 * n_mod] = total;
}

// backward seeds: per-column scale of the raw prediction gradients and d(total)/d(log_vars)
struct SeedArgs {
  const float *acc, *log_vars, *losses, *grad_out;
  float *d_log_vars;     // accumulates
  float *colscale_pix;   // [npix]
  float *colscale_img;   // [nimg]
  int n_mod, uncertainty;
  int col_off[MPMAE_MAX_MOD], col_len[MPMAE_MAX_MOD], is_img[MPMAE_MAX_MOD];
};
__global__ void loss_seed_kernel(SeedArgs a) {
  const int i = blockIdx.x;
  const float go = a.grad_out ? a.grad_out[0] : 1.f;
  const float L = a.losses[i];
  float dL = 1.f;
  if (a.uncertainty) {
    const float s = a.log_vars[i];
    dL = (L != 0.f) ? expf(-s) : 0.f;
    if (threadIdx.x == 0) atomicAdd(&a.d_log_vars[i], go * ((L != 0.f) ? 1.f - expf(-s) * L : 0.f));
  }
  const float den = a.acc[2 * i + 1];
  const float sc = den > 0.f ? go * dL / den : 0.f;
  float *dst = a.is_img[i] ? a.colscale_img : a.colscale_pix;
  for (int c = threadIdx.x; c < a.col_len[i]; c += blockDim.x) dst[a.col_off[i] + c] = sc;
}

}  // namespace mpmae
