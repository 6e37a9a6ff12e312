// fp32 SIMT GEMMs with fused epilogues.  These are the always-correct baseline the tcgen05 path
// (gemm_tc.cuh) is checked against, and they serve the shapes the tensor-core path does not take
// (tiny M, ragged N).
//
//   gemm_rows : out[M,N] = A[M,K] . Bw[N,K]^T  (+ epilogue)      -- pointwise convs, dX products
//   gemm_wgrad: dW[N,K] += rs[n] * sum_r X[r,n] * Y[r,k]          -- weight gradients (split over r)
#pragma once
#include "common.cuh"

namespace mpmae {

enum EpiMode : int {
  EPI_STORE = 0,    // out = acc + bias (+ resid)
  EPI_GELU_SQ = 1,  // a = acc + bias -> out ; h = gelu(a) -> out2 (when given) ; colsum[g, n] += h^2
  EPI_DG = 2,       // out = acc ; colsum[g, n] += acc * aux[m, n] ; colsum2[n] += acc
  EPI_DH_GELU = 3   // out = (acc * acc_scale[n] + kg[g, n] * gelu(aux2[m, n])) * gelu'(aux2[m, n]) ; colsum2[n] += out
};

constexpr int kMaxGroupsPerTile = 12;

struct GemmArgs {
  const float *A;      // [M, K] row-major
  const float *Bw;     // [N, K] row-major ("weight [out, in]")
  const float *Bw_lo;  // optional low part of a TF32 hi/lo split of the weight (Bw then holds the high part); null = none
  const float *ln_xhat;  // EPI_STORE on the tcgen05 path only: when set (with ln_rstd) the epilogue applies the LayerNorm
  const float *ln_rstd;  // backward to the product row: out = rstd * (acc - mean(acc) - xhat * mean(acc * xhat))
  int b16;             // 3xBF16: Bw is the full fp32 weight and Bw_lo points at the packed bf16 pair (hi [N,K] | lo [N,K])
  const float *bias;   // [N] or null
  const float *resid;  // [M, N] or null (EPI_STORE)
  float *out;          // [M, N]
  float *out2;         // [M, N] (EPI_GELU_SQ: h)
  const float *aux;    // [M, N] (EPI_DG / EPI_DH_GELU: h)
  const float *aux2;   // [M, N] (EPI_DH_GELU: a)
  const float *kg;     // [groups, N] (EPI_DH_GELU)
  float *colsum;       // [groups, N]
  float *colsum2;      // [N]
  int64_t M;
  int N, K;
  int group_rows;      // rows per statistics group (>= M means one group)
  // The A operand consumed as gelu(A[m, k]) * a_scale[k] (a_scale null = 1): pw2 of a sparse block reads the saved
  // pre-activation `a` and applies GELU and the GRN scale on the way into the tensor core, so `h` is never materialised
  int a_gelu;
  const float *a_scale;    // [K] or null
  const float *acc_scale;  // [N] or null (EPI_DH_GELU): per-column factor of the accumulator (the GRN scale of dg = dy . W2)
  // Batch-global GRN statistic (models/sparse_norm_layers.py:24-33) computed by the a_gelu kernel itself (tcgen05 path):
  // when grn_gsq = sum_rows h^2 [K] is given, every CTA derives nx = sqrt(gsq) / (mean sqrt(gsq) + eps) and the A-operand
  // scale s = 1 + gamma * nx in its prologue (it overlaps the pipeline fill) and CTA 0 writes nx / scale / denom for the
  // backward pass; a_scale is then ignored.
  const float *grn_gsq, *grn_gamma;   // [K]
  float *grn_nx, *grn_scale, *grn_denom;
  float grn_eps;
  // Backward of the batch-global GRN statistic in the prologue of the EPI_DH_GELU kernel (tcgen05 path; replaces kg): from
  // ds[n] = sum_rows dg * h (complete before the launch), nx, the denominator and gamma every CTA derives
  //   dNx = gamma * ds ; dGx = dNx / den - (sum_j dNx_j Gx_j) / (N den^2) ; kg[n] = dGx / Gx   
  // and CTA 0 accumulates dgamma[n] += nx * ds.
  const float *grnb_ds, *grnb_nx, *grnb_denom, *grnb_gamma;
  float *grnb_dgamma;
};

// Column accumulation helper shared by the SIMT and tcgen05 epilogues: a thread owns `nrows`
// consecutive rows of one column; rows of equal group are summed in registers first.
__device__ __forceinline__ void colacc_add(float *colacc, int bn, int col, int g, float v) {
  atomicAdd(&colacc[g * bn + col], v);
}

template <int NT, int TN, int MODE>
__global__ void __launch_bounds__(16 * NT) gemm_rows_kernel(GemmArgs p) { pdl_prologue();
  constexpr int BM = 128, TM = 8, BK = 8, BN = NT * TN, NTHR = 16 * NT;
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  __shared__ float colacc[(MODE == EPI_STORE) ? 1 : kMaxGroupsPerTile * BN];
  __shared__ float colacc2[(MODE == EPI_DG || MODE == EPI_DH_GELU) ? BN : 1];

  const int tid = threadIdx.x;
  const int tx = tid % NT, ty = tid / NT;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;

  if (MODE != EPI_STORE) {
    for (int i = tid; i < kMaxGroupsPerTile * BN; i += NTHR) colacc[i] = 0.f;
    if (MODE == EPI_DG || MODE == EPI_DH_GELU)
      for (int i = tid; i < BN; i += NTHR) colacc2[i] = 0.f;
  }

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < p.K; k0 += BK) {
    // A tile: BM x BK = 256 float4
    for (int t = tid; t < BM * 2; t += NTHR) {
      const int r = t >> 1, kq = (t & 1) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m0 + r < p.M) v = *reinterpret_cast<const float4 *>(p.A + (m0 + r) * p.K + k0 + kq);
      if (p.a_gelu) {
        const float4 sc = p.a_scale ? *reinterpret_cast<const float4 *>(p.a_scale + k0 + kq) : make_float4(1.f, 1.f, 1.f, 1.f);
        v.x = gelu_f(v.x) * sc.x; v.y = gelu_f(v.y) * sc.y; v.z = gelu_f(v.z) * sc.z; v.w = gelu_f(v.w) * sc.w;
      }
      As[kq + 0][r] = v.x; As[kq + 1][r] = v.y; As[kq + 2][r] = v.z; As[kq + 3][r] = v.w;
    }
    for (int t = tid; t < BN * 2; t += NTHR) {
      const int r = t >> 1, kq = (t & 1) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (n0 + r < p.N) {
        v = *reinterpret_cast<const float4 *>(p.Bw + (int64_t)(n0 + r) * p.K + k0 + kq);
        if (p.Bw_lo && !p.b16) {
          const float4 l = *reinterpret_cast<const float4 *>(p.Bw_lo + (int64_t)(n0 + r) * p.K + k0 + kq);
          v.x += l.x; v.y += l.y; v.z += l.z; v.w += l.w;
        }
      }
      Bs[kq + 0][r] = v.x; Bs[kq + 1][r] = v.y; Bs[kq + 2][r] = v.z; Bs[kq + 3][r] = v.w;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], b[TN];
      const float4 a0 = *reinterpret_cast<const float4 *>(&As[k][ty * TM]);
      const float4 a1 = *reinterpret_cast<const float4 *>(&As[k][ty * TM + 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Bs[k][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  // ---------------------------------------------------------------- epilogue
  const int64_t g_first = m0 / p.group_rows;
  float part[TN], part2[TN];
#pragma unroll
  for (int j = 0; j < TN; ++j) { part[j] = 0.f; part2[j] = 0.f; }
  int g_cur = -1;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int64_t m = m0 + ty * TM + i;
    if (m >= p.M) break;
    int g = 0;
    if (MODE != EPI_STORE) {
      g = (int)(m / p.group_rows - g_first);
      if (g != g_cur) {
        if (g_cur >= 0) {
#pragma unroll
          for (int j = 0; j < TN; ++j) {
            if (MODE != EPI_DH_GELU) colacc_add(colacc, BN, tx * TN + j, g_cur, part[j]);
            part[j] = 0.f;
          }
        }
        g_cur = g;
      }
    }
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx * TN + j;
      if (n >= p.N) continue;
      const int64_t o = m * p.N + n;
      float v = acc[i][j];
      if (MODE == EPI_STORE) {
        if (p.bias) v += p.bias[n];
        if (p.resid) v += p.resid[o];
        p.out[o] = v;
      } else if (MODE == EPI_GELU_SQ) {
        if (p.bias) v += p.bias[n];
        p.out[o] = v;
        const float h = gelu_f(v);
        if (p.out2) p.out2[o] = h;
        part[j] += h * h;
      } else if (MODE == EPI_DG) {
        p.out[o] = v;
        part[j] += v * p.aux[o];
        part2[j] += v;
      } else {  // EPI_DH_GELU
        const float kgv = p.kg ? p.kg[(g_first + g) * p.N + n] : 0.f;
        const float hv = p.aux ? p.aux[o] : gelu_f(p.aux2[o]);
        if (p.acc_scale) v *= p.acc_scale[n];
        const float da = (v + kgv * hv) * gelu_grad_f(p.aux2[o]);
        p.out[o] = da;
        part2[j] += da;
      }
    }
  }
  if (MODE != EPI_STORE) {
    if (g_cur >= 0) {
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        if (MODE != EPI_DH_GELU) colacc_add(colacc, BN, tx * TN + j, g_cur, part[j]);
        if (MODE == EPI_DG || MODE == EPI_DH_GELU) atomicAdd(&colacc2[tx * TN + j], part2[j]);
      }
    }
    __syncthreads();
    const int64_t m_last = (m0 + BM < p.M ? m0 + BM : p.M) - 1;
    const int ng = (int)(m_last / p.group_rows - g_first) + 1;
    if (MODE != EPI_DH_GELU && p.colsum) {
      for (int i = tid; i < ng * BN; i += NTHR) {
        const int g = i / BN, c = i % BN;
        if (n0 + c < p.N) atomicAdd(&p.colsum[(g_first + g) * p.N + n0 + c], colacc[g * BN + c]);
      }
    }
    if ((MODE == EPI_DG || MODE == EPI_DH_GELU) && p.colsum2) {
      for (int c = tid; c < BN; c += NTHR)
        if (n0 + c < p.N) atomicAdd(&p.colsum2[n0 + c], colacc2[c]);
    }
  }
}

template <int MODE>
inline cudaError_t launch_gemm_rows_simt(const GemmArgs &p, cudaStream_t st) {
  if (p.M <= 0) return cudaSuccess;
  const unsigned gm = (unsigned)cdiv64(p.M, 128);
  if (p.N % 80 == 0) {
    pdl(gemm_rows_kernel<16, 5, MODE>, dim3(gm, p.N / 80), 256, 0, st)(p);
  } else if (p.N % 64 == 0 || p.N > 64) {
    pdl(gemm_rows_kernel<16, 4, MODE>, dim3(gm, cdiv(p.N, 64)), 256, 0, st)(p);
  } else if (p.N % 40 == 0 || p.N > 32) {
    pdl(gemm_rows_kernel<8, 5, MODE>, dim3(gm, cdiv(p.N, 40)), 128, 0, st)(p);
  } else {
    pdl(gemm_rows_kernel<8, 4, MODE>, dim3(gm, cdiv(p.N, 32)), 128, 0, st)(p);
  }
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Weight gradient: dW[n, k] += rs[n] * sum_r X[r, n] * Y[r, k];  db[n] += rs[n] * sum_r X[r, n].
// X = upstream gradient [R, N], Y = forward operand [R, K].  The reduction runs over rows, so the
// work is split over blockIdx.z and merged with fp32 atomics (the reference merges with atomics
// too: MinkowskiEngine/src/convolution_kernel.cu:198-290).
struct WgradArgs {
  const float *X;   // [R, N]
  const float *Y;   // [R, K]
  const float *rs;  // [N] row scale of the result or null
  float *dW;        // [N, K]
  float *db;        // [N] or null
  int64_t R;
  int N, K;
  int rows_per_split;
  int exact;        // tensor-core path: 3xTF32 (the result feeds back into the data path)
  int y_gelu;       // Y is consumed as gelu(Y): dW2f = dy^T . gelu(a) without a materialised h
};

__device__ __forceinline__ float4 load4_guard(const float *base, int64_t row, int ld, int col, int ncols, bool vec_ok) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  const float *ptr = base + row * ld + col;
  if (vec_ok && col + 3 < ncols) {
    v = *reinterpret_cast<const float4 *>(ptr);
  } else {
    if (col + 0 < ncols) v.x = ptr[0];
    if (col + 1 < ncols) v.y = ptr[1];
    if (col + 2 < ncols) v.z = ptr[2];
    if (col + 3 < ncols) v.w = ptr[3];
  }
  return v;
}

__global__ void __launch_bounds__(256) gemm_wgrad_kernel(WgradArgs p) { pdl_prologue();
  constexpr int T = 64, RC = 16;
  __shared__ __align__(16) float Xs[RC][T];
  __shared__ __align__(16) float Ys[RC][T];
  const int tid = threadIdx.x;
  const int tk = tid % 16, tn = tid / 16;
  const int n0 = blockIdx.x * T, k0 = blockIdx.y * T;
  const int64_t r_begin = (int64_t)blockIdx.z * p.rows_per_split;
  const int64_t r_end = (r_begin + p.rows_per_split < p.R) ? r_begin + p.rows_per_split : p.R;
  const bool xvec = (p.N % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.X) & 15) == 0);
  const bool yvec = (p.K % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.Y) & 15) == 0);

  float acc[4][4];
  float bsum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const bool do_bias = (p.db != nullptr) && (blockIdx.y == 0) && (tk == 0);

  const int lr = tid / 16, lc = (tid % 16) * 4;  // loader: row within chunk, column within tile
  for (int64_t r0 = r_begin; r0 < r_end; r0 += RC) {
    float4 xv = make_float4(0.f, 0.f, 0.f, 0.f), yv = xv;
    if (r0 + lr < r_end) {
      xv = load4_guard(p.X, r0 + lr, p.N, n0 + lc, p.N, xvec);
      yv = load4_guard(p.Y, r0 + lr, p.K, k0 + lc, p.K, yvec);
      if (p.y_gelu) { yv.x = gelu_f(yv.x); yv.y = gelu_f(yv.y); yv.z = gelu_f(yv.z); yv.w = gelu_f(yv.w); }
    }
    *reinterpret_cast<float4 *>(&Xs[lr][lc]) = xv;
    *reinterpret_cast<float4 *>(&Ys[lr][lc]) = yv;
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RC; ++r) {
      const float4 x = *reinterpret_cast<const float4 *>(&Xs[r][tn * 4]);
      const float4 y = *reinterpret_cast<const float4 *>(&Ys[r][tk * 4]);
      const float xa[4] = {x.x, x.y, x.z, x.w};
      const float ya[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(xa[i], ya[j], acc[i][j]);
        bsum[i] += xa[i];
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + tn * 4 + i;
    if (n >= p.N) continue;
    const float s = p.rs ? p.rs[n] : 1.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + tk * 4 + j;
      if (k < p.K) atomicAdd(&p.dW[(int64_t)n * p.K + k], s * acc[i][j]);
    }
    if (do_bias) atomicAdd(&p.db[n], s * bsum[i]);
  }
}

inline cudaError_t launch_gemm_wgrad(WgradArgs p, cudaStream_t st, int target_ctas = 592) {
  if (p.R <= 0) return cudaSuccess;
  const int tiles = cdiv(p.N, 64) * cdiv(p.K, 64);
  int64_t splits = cdiv(target_ctas, tiles);
  const int64_t max_splits = cdiv64(p.R, 64);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  int64_t rps = cdiv64(p.R, splits);
  rps = cdiv64(rps, 16) * 16;
  splits = cdiv64(p.R, rps);
  p.rows_per_split = (int)rps;
  pdl(gemm_wgrad_kernel, dim3(cdiv(p.N, 64), cdiv(p.K, 64), (unsigned)splits), 256, 0, st)(p);
  return cudaGetLastError();
}

}  // namespace mpmae
