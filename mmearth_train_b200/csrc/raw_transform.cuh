// The reference loader's per-sample transform (mmearth_dataset.py:58-153) for a whole batch of STORED arrays, one pass:
// band selection, no-data -> NaN, label remap through a 256-entry table, per-band z-scoring with the Sentinel-2 L1C / L2A
// statistics picked per sample, NaN -> -1 for class targets.  The arithmetic is the reference's -- float64 subtraction and
// division, then one rounding to float32 -- so the result is bit-identical to MMEarthDataset.__getitem__.
// What crosses PCIe is the stored uint16 / uint8 / float32 data (56.7 MB per 256 samples instead of 91.7 MB widened).
#pragma once
#include "common.cuh"
#include "../../include/mpmae.h"

namespace mpmae {

struct RawArgs {
  const void *src;          // [B, src_bands, inner]
  void *out;                // [B, n_bands, inner] float32 or int64
  const uint8_t *l2a;       // [B] or null: which statistics set applies to the sample
  const double *lut;        // [256] or null: stored byte -> value (NaN = ignore), applied before everything else
  int64_t inner;            // elements per (sample, band): H*W, or 1 for vectors
  int B, src_bands, n_bands;
  int src_type;             // 0 = uint8, 1 = uint16, 2 = float32
  int out_int64;            // 1: out = isnan(v) ? -1 : (int64) v ; 0: out = (float) v
  int has_nodata, normalize;
  double nodata;
  int band[MPMAE_RAW_MAX_BANDS];
  double mean[2][MPMAE_RAW_MAX_BANDS], stdv[2][MPMAE_RAW_MAX_BANDS];
};

__global__ void __launch_bounds__(256) raw_transform_kernel(RawArgs a) { pdl_prologue();
  const int64_t per_sample = (int64_t)a.n_bands * a.inner, total = (int64_t)a.B * per_sample;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(i / per_sample);
    const int64_t rem = i - (int64_t)n * per_sample;
    const int b = (int)(rem / a.inner);
    const int64_t j = rem - (int64_t)b * a.inner;
    const int64_t s = ((int64_t)n * a.src_bands + a.band[b]) * a.inner + j;
    double v;
    if (a.src_type == 0) v = (double)static_cast<const uint8_t *>(a.src)[s];
    else if (a.src_type == 1) v = (double)static_cast<const uint16_t *>(a.src)[s];
    else v = (double)static_cast<const float *>(a.src)[s];
    if (a.lut) {
      v = a.lut[(int)v & 255];
    } else if (a.has_nodata && v == a.nodata) {
      v = __longlong_as_double(0x7ff8000000000000ll);
    }
    if (a.normalize) {
      const int set = (a.l2a && a.l2a[n]) ? 1 : 0;
      v = (v - a.mean[set][b]) / a.stdv[set][b];
    }
    if (a.out_int64) static_cast<long long *>(a.out)[i] = (v != v) ? -1ll : (long long)v;
    else static_cast<float *>(a.out)[i] = (float)v;
  }
}

inline cudaError_t launch_raw_transform(const RawArgs &a, cudaStream_t st) {
  const int64_t total = (int64_t)a.B * a.n_bands * a.inner;
  if (total <= 0) return cudaSuccess;
  int64_t grid = cdiv64(total, 256 * 4);
  if (grid > 148 * 16) grid = 148 * 16;
  if (grid < 1) grid = 1;
  pdl(raw_transform_kernel, (unsigned)grid, 256, 0, st)(a);
  return cudaGetLastError();
}

}  // namespace mpmae
