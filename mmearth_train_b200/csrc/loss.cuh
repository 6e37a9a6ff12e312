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
  int smem_cache;                    // pixel_loss_kernel was launched with sum(p^2 * chans) floats of dynamic shared memory
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

// One CTA per patch cell, ONE WARP PER PIXEL MODALITY (blockDim.x = 32 * number of pixel modalities): the modalities of
// a cell are independent, so nothing is block-synchronised and every reduction is a warp shuffle.
__global__ void __launch_bounds__(32 * MPMAE_MAX_MOD) pixel_loss_kernel(LossArgs a) { pdl_prologue();
  const int cell = blockIdx.x;               // n*L + l
  const int n = cell / a.L, l = cell - n * a.L;
  const int ph = l / a.G, pw = l - ph * a.G;
  const int p = a.p, p2 = p * p;
  const bool masked = a.mask[cell] != 0.f;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int mi = -1;
  for (int i = 0, k = 0; i < a.n_mod; ++i)
    if (a.mod[i].kind == MPMAE_PIXEL_CONTINUOUS || a.mod[i].kind == MPMAE_PIXEL_CATEGORICAL) {
      if (k == warp) { mi = i; break; }
      ++k;
    }
  if (mi < 0) return;
  {
    const LossMod m = a.mod[mi];
    const int len = p2 * m.chans;
    const float *pr = a.pred_pix + (int64_t)cell * a.npix + m.col_off;
    float *dp = a.dpix + (int64_t)cell * a.npix + m.col_off;
    if (!masked) {  // visible patch: no loss, zero gradient
      for (int j = lane; j < len; j += 32) dp[j] = 0.f;
      return;
    }
    if (m.kind == MPMAE_PIXEL_CONTINUOUS) {
      const float *tg = static_cast<const float *>(m.target);
      const int c = m.chans;
      const float *tbase = tg + ((int64_t)n * c * a.S + ph * p) * a.S + pw * p;   // + (ch*S + pi)*S + qi
      auto target_at = [&](int j) {
        const int q = j / c, ch = j - q * c, pi = q / p, qi = q - pi * p;
        return nan_to_zero(tbase[((int64_t)ch * a.S + pi) * a.S + qi]);
      };
      constexpr int NE = 24;   // register-cached path: up to 768 elements per patch (patch 8: 64 pixels x 12 bands)
      if (len <= 32 * NE && !a.smem_cache) {
        float tv[NE], pv[NE];
#pragma unroll
        for (int k = 0; k < NE; ++k) {
          const int j = lane + 32 * k;
          tv[k] = 0.f; pv[k] = 0.f;
          if (j < len) { tv[k] = target_at(j); pv[k] = pr[j]; }
        }
        float mean = 0.f, inv_std = 1.f;
        if (m.norm_pix) {
          float s = 0.f;
#pragma unroll
          for (int k = 0; k < NE; ++k) s += tv[k];
          mean = warp_sum(s) / (float)len;
          float v = 0.f;
#pragma unroll
          for (int k = 0; k < NE; ++k)
            if (lane + 32 * k < len) { const float d = tv[k] - mean; v += d * d; }
          inv_std = rsqrtf(warp_sum(v) / (float)(len - 1) + 1.0e-6f);
        }
        float se = 0.f, cnt = 0.f;
#pragma unroll
        for (int k = 0; k < NE; ++k)
          if (lane + 32 * k < len) {
            pv[k] -= (tv[k] - mean) * inv_std;            // d
            const float e = pv[k] * pv[k];
            if (e == e) { se += e; cnt += 1.f; }
          }
        se = warp_sum(se); cnt = warp_sum(cnt);
        const float patch_loss = se / cnt;
        const bool counted = (patch_loss == patch_loss) && patch_loss != 0.f;
        if (lane == 0 && counted) { atomicAdd(&a.acc[2 * mi], patch_loss); atomicAdd(&a.acc[2 * mi + 1], 1.f); }
        const float sc = 2.f / cnt;
#pragma unroll
        for (int k = 0; k < NE; ++k) {
          const int j = lane + 32 * k;
          if (j < len) dp[j] = (counted && pv[k] == pv[k]) ? pv[k] * sc : 0.f;
        }
        return;
      }
      if (a.smem_cache) {
        // larger patches (patch 16: 256 pixels x c bands): the cleaned target is staged ONCE in shared memory, read from the
        // NCHW image in MEMORY order (a patch row of a band is contiguous: coalesced) and stored in the prediction's
        // (pixel, band) order; the three passes below then run out of shared memory.  The gather path under this block
        // evaluated target_at() three times with a band-strided pattern (0.52 ms of a 5.8 ms cfg3 step).
        extern __shared__ float tcache[];
        int off = 0;
        for (int i = 0; i < mi; ++i)
          if (a.mod[i].kind == MPMAE_PIXEL_CONTINUOUS) off += p2 * a.mod[i].chans;
        float *ts = tcache + off;
        float s = 0.f;
        for (int i = lane; i < len; i += 32) {
          const int ch = i / p2, r = i - ch * p2, pi = r / p, qi = r - pi * p;
          const float t = nan_to_zero(tbase[((int64_t)ch * a.S + pi) * a.S + qi]);
          ts[r * c + ch] = t;
          s += t;
        }
        __syncwarp();
        float mean = 0.f, inv_std = 1.f;
        if (m.norm_pix) {
          mean = warp_sum(s) / (float)len;
          float v = 0.f;
          for (int j = lane; j < len; j += 32) { const float d = ts[j] - mean; v += d * d; }
          inv_std = rsqrtf(warp_sum(v) / (float)(len - 1) + 1.0e-6f);
        }
        float se = 0.f, cnt = 0.f;
        for (int j = lane; j < len; j += 32) {
          const float d = pr[j] - (ts[j] - mean) * inv_std;
          const float e = d * d;
          if (e == e) { se += e; cnt += 1.f; }
          ts[j] = d;                                   // own element: no hazard
        }
        se = warp_sum(se); cnt = warp_sum(cnt);
        const float patch_loss = se / cnt;
        const bool counted = (patch_loss == patch_loss) && patch_loss != 0.f;
        if (lane == 0 && counted) { atomicAdd(&a.acc[2 * mi], patch_loss); atomicAdd(&a.acc[2 * mi + 1], 1.f); }
        const float sc = 2.f / cnt;
        for (int j = lane; j < len; j += 32) { const float d = ts[j]; dp[j] = (counted && d == d) ? d * sc : 0.f; }
        return;
      }
      float mean = 0.f, inv_std = 1.f;
      if (m.norm_pix) {  // per-patch (t - mean) / sqrt(var_unbiased + 1e-6), fcmae.py:377-382
        float s = 0.f;
        for (int j = lane; j < len; j += 32) s += target_at(j);
        mean = warp_sum(s) / (float)len;
        float v = 0.f;
        for (int j = lane; j < len; j += 32) { const float d = target_at(j) - mean; v += d * d; }
        inv_std = rsqrtf(warp_sum(v) / (float)(len - 1) + 1.0e-6f);
      }
      float se = 0.f, cnt = 0.f;
      for (int j = lane; j < len; j += 32) {
        const float d = pr[j] - (target_at(j) - mean) * inv_std;
        const float e = d * d;
        if (e == e) { se += e; cnt += 1.f; }   // NaN squared errors are dropped (fcmae.py:385-388)
      }
      se = warp_sum(se); cnt = warp_sum(cnt);
      const float patch_loss = se / cnt;       // cnt == 0 -> NaN -> dropped below
      const bool counted = (patch_loss == patch_loss) && patch_loss != 0.f;
      if (lane == 0 && counted) { atomicAdd(&a.acc[2 * mi], patch_loss); atomicAdd(&a.acc[2 * mi + 1], 1.f); }
      for (int j = lane; j < len; j += 32) {
        const float d = pr[j] - (target_at(j) - mean) * inv_std;
        dp[j] = (counted && d == d) ? 2.f * d / cnt : 0.f;
      }
    } else {  // categorical pixels: logits [p2][K], int64 target, -1 ignored (fcmae.py:302-346)
      const long long *tg = static_cast<const long long *>(m.target);
      const int K = m.chans;
      float ls = 0.f, nsel = 0.f;
      for (int q = lane; q < p2; q += 32) {
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
      ls = warp_sum(ls); nsel = warp_sum(nsel);
      if (lane == 0 && nsel > 0.f) { atomicAdd(&a.acc[2 * mi], ls); atomicAdd(&a.acc[2 * mi + 1], nsel); }
    }
  }
}

// image-level heads (fcmae.py:281-301): one CTA per sample, one warp per image modality (blockDim.x = 32 * their count)
__global__ void __launch_bounds__(32 * MPMAE_MAX_MOD) image_loss_kernel(LossArgs a) { pdl_prologue();
  const int n = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int mi = -1;
  for (int i = 0, k = 0; i < a.n_mod; ++i)
    if (a.mod[i].kind == MPMAE_IMAGE_CATEGORICAL || a.mod[i].kind == MPMAE_IMAGE_CONTINUOUS) {
      if (k == warp) { mi = i; break; }
      ++k;
    }
  if (mi < 0) return;
  const LossMod m = a.mod[mi];
  const int c = m.chans;
  const float *pr = a.pred_img + (int64_t)n * a.nimg + m.col_off;
  float *dp = a.dimg + (int64_t)n * a.nimg + m.col_off;
  if (m.kind == MPMAE_IMAGE_CATEGORICAL) {
    const long long *tg = static_cast<const long long *>(m.target) + (int64_t)n * c;
    // argmax of the one-hot row (first maximum, as torch.argmax): lexicographic (value desc, index asc) reduction
    long long best = tg[lane < c ? lane : 0];
    int cls = lane < c ? lane : 0;
    for (int k = lane + 32; k < c; k += 32) {
      const long long t = tg[k];
      if (t > best) { best = t; cls = k; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const long long ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oc = __shfl_xor_sync(0xffffffffu, cls, o);
      if (ob > best || (ob == best && oc < cls)) { best = ob; cls = oc; }
    }
    float mx = -INFINITY;
    for (int k = lane; k < c; k += 32) mx = fmaxf(mx, pr[k]);
    mx = warp_max(mx);
    float se = 0.f;
    for (int k = lane; k < c; k += 32) se += expf(pr[k] - mx);
    se = warp_sum(se);
    const float lse = mx + logf(se);
    if (lane == 0) { atomicAdd(&a.acc[2 * mi], lse - pr[cls]); atomicAdd(&a.acc[2 * mi + 1], 1.f); }
    for (int k = lane; k < c; k += 32) dp[k] = expf(pr[k] - lse) - (k == cls ? 1.f : 0.f);
  } else {
    const float *tg = static_cast<const float *>(m.target) + (int64_t)n * c;
    float se = 0.f, cnt = 0.f;
    for (int k = lane; k < c; k += 32) {
      const float t = tg[k];
      if (t == t) {
        const float d = pr[k] - t;
        se += d * d; cnt += 1.f; dp[k] = 2.f * d;
      } else {
        dp[k] = 0.f;
      }
    }
    se = warp_sum(se); cnt = warp_sum(cnt);
    if (lane == 0 && cnt > 0.f) { atomicAdd(&a.acc[2 * mi], se); atomicAdd(&a.acc[2 * mi + 1], cnt); }
  }
}

// losses[0..n) = L_i ; losses[n..2n) = weighted_i ; losses[2n] = total
__global__ void loss_finalize_kernel(const float *__restrict__ acc, const float *__restrict__ log_vars, int n_mod,
                                     int uncertainty, float *__restrict__ losses) { pdl_prologue();
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
__global__ void loss_seed_kernel(SeedArgs a) { pdl_prologue();
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
