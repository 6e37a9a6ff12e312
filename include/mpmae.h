/*
 * mpmae.h -- C ABI of the B200-native MP-MAE (FCMAE) pretraining step.
 *
 * Drop-in boundary.  The reference reaches its native code through the
 * pybind module MinkowskiEngineBackend._C:
 *     MinkowskiEngine/pybind/extern.hpp:118-145   DepthwiseConvolution{Forward,Backward}GPU
 *     MinkowskiEngine/pybind/extern.hpp:610-625   py::class_/m.def registrations (GIL released)
 *     MinkowskiEngine/MinkowskiEngine/MinkowskiDepthwiseConvolution.py:52-64   the call site
 * and through ATen/cuBLAS for everything else in models/fcmae.py:FCMAE.forward (lines 414-456).
 * This library replaces that whole forward+backward with two calls.  Plain C types only; every
 * device buffer is allocated by the caller (torch caching allocator) and passed in; the library
 * allocates no device memory, keeps no global state besides immutable constants, never
 * synchronises the device and never throws across the boundary.
 *
 * All functions return 0 on success or a negative mpmae_status; mpmae_last_error() gives a
 * thread-local message.
 */
#ifndef MPMAE_H
#define MPMAE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPMAE_MAX_MOD 16

typedef enum mpmae_status {
  MPMAE_OK = 0,
  MPMAE_ERR_INVALID = -1,     /* bad configuration / null pointer                     */
  MPMAE_ERR_UNSUPPORTED = -2, /* shape outside what the kernels were instantiated for */
  MPMAE_ERR_CUDA = -3,        /* a CUDA launch or API call failed                     */
  MPMAE_ERR_WORKSPACE = -4    /* workspace too small                                  */
} mpmae_status;

/* modality kinds, models/fcmae.py:281-403 */
typedef enum mpmae_mod_kind {
  MPMAE_PIXEL_CONTINUOUS = 0,  /* sentinel2, sentinel1, aster, canopy_height_eth : float [B,c,S,S]      */
  MPMAE_PIXEL_CATEGORICAL = 1, /* dynamic_world, esa_worldcover                  : int64 [B,1,S,S], -1 = ignore */
  MPMAE_IMAGE_CATEGORICAL = 2, /* biome, eco_region                              : int64 one-hot [B,c]  */
  MPMAE_IMAGE_CONTINUOUS = 3   /* lat, lon, month, era5                          : float [B,c], NaN = ignore */
} mpmae_mod_kind;

/* Mirrors the FCMAE constructor (models/fcmae.py:30-43) + args.out_modalities / args.loss_aggr. */
typedef struct mpmae_cfg {
  int32_t batch;      /* per-GPU batch B                                   */
  int32_t img_size;   /* 56 or 112 (img_size / patch_size = patch grid G)   */
  int32_t patch_size; /* multiple of 8: 8 or 16                             */
  int32_t in_chans;   /* sentinel-2 bands fed to the encoder (12)           */
  int32_t depths[4];
  int32_t dims[4];
  int32_t dec_dim;       /* decoder_embed_dim                               */
  int32_t dec_depth;     /* decoder_depth                                   */
  float mask_ratio;      /* visible patches = (int)(L * (1 - mask_ratio))   */
  int32_t loss_aggr;     /* 0 = unweighted, 1 = uncertainty                 */
  int32_t n_mod;         /* output modalities, in args.out_modalities order */
  int32_t mod_kind[MPMAE_MAX_MOD];
  int32_t mod_chans[MPMAE_MAX_MOD];    /* out_chans (classes for categorical) */
  int32_t mod_norm_pix[MPMAE_MAX_MOD]; /* per-patch target normalisation (sentinel2 && norm_pix_loss) */
  int32_t gemm_backend;  /* 0 = fp32 SIMT tiles, 1 = tcgen05 3xTF32 (fp32-faithful), 2 = tcgen05 single-pass TF32,
                          * 3 = tcgen05 3xBF16 (operands split into bf16 hi + lo, |error| <= 2^-16 per product; the default) */
} mpmae_cfg;

typedef struct mpmae_plan mpmae_plan; /* opaque: parameter layout, workspace layout, launch plan */

/* Per-step device pointers.  Everything is fp32 unless stated. */
typedef struct mpmae_io {
  const float *params;     /* flat parameter buffer, layout given by mpmae_param_info     */
  float *grads;            /* flat gradient buffer, same layout; backward ACCUMULATES      */
  void *workspace;         /* >= mpmae_workspace_bytes(plan), 256-byte aligned             */
  size_t workspace_bytes;
  const float *noise;      /* [B, L]  N(0,1) noise of gen_random_mask (fcmae.py:220)       */
  const float *s2_input;   /* [B, in_chans, S, S] encoder input (before nan_to_num)        */
  const void *targets[MPMAE_MAX_MOD]; /* per output modality, dtype/shape by kind          */
  float *mask;             /* out [B, L]  0 keep / 1 remove                                 */
  float *pred_pixel;       /* out [B*L, n_pix_cols] channels-last pixel-head predictions    */
  float *pred_image;       /* out [B, n_img_cols] image-head predictions                    */
  float *losses;           /* out [2*n_mod + 1]: per-modality loss, weighted loss, total    */
  const float *grad_out;   /* backward only: d(total)/d(total) upstream scalar (device)     */
  int32_t *flags;          /* out [4] device flags: [0] all-zero visible input pixels seen  */
} mpmae_io;

const char *mpmae_last_error(void);
int mpmae_version(void);

int mpmae_plan_create(const mpmae_cfg *cfg, mpmae_plan **out);
void mpmae_plan_destroy(mpmae_plan *plan);

/* parameter layout: reference state-dict names, shapes and offsets (floats) in the flat buffer */
int64_t mpmae_param_total(const mpmae_plan *plan);
int32_t mpmae_param_count(const mpmae_plan *plan);
int mpmae_param_info(const mpmae_plan *plan, int32_t index, char *name, int32_t name_cap,
                     int64_t shape[4], int32_t *ndim, int64_t *offset);
/* 1 if AdamW weight decay applies to parameter `index` under the rule the reference uses
 * (timm param_groups_weight_decay, main_pretrain.py:312-319: no decay for ndim <= 1 or "*.bias") */
int32_t mpmae_param_decay(const mpmae_plan *plan, int32_t index);
/* visible patches per sample = int(L * (1 - mask_ratio)), models/fcmae.py:216-217 */
int32_t mpmae_visible_patches(const mpmae_plan *plan);

size_t mpmae_workspace_bytes(const mpmae_plan *plan);
int32_t mpmae_pred_pixel_cols(const mpmae_plan *plan);
int32_t mpmae_pred_image_cols(const mpmae_plan *plan);
/* column offset of modality m inside pred_pixel / pred_image (by kind) */
int32_t mpmae_pred_col_offset(const mpmae_plan *plan, int32_t mod);

/* named intermediate tensors inside the workspace (parity tests): offset in bytes, rows, cols */
int mpmae_tap_info(const mpmae_plan *plan, const char *name, int64_t *byte_offset, int64_t *rows,
                   int64_t *cols);

int32_t mpmae_tap_count(const mpmae_plan *plan);
int mpmae_tap_name(const mpmae_plan *plan, int32_t index, char *name, int32_t name_cap);

/* number of kernels launched by one forward / backward call (bench.py "gpu_launches") */
int32_t mpmae_launch_count(const mpmae_plan *plan, int32_t backward);

/* Per-launch device timing (bench.py roofline leg): after mpmae_profile_begin every launch of
 * mpmae_forward / mpmae_backward is bracketed by CUDA events on the caller's stream; mpmae_profile_report
 * synchronises on the last event, stops profiling and writes CSV "name,launches,ms,alg_bytes,alg_flops"
 * (algorithmic bytes/flops per SURVEY.md section 8d) aggregated by kernel call site. */
int mpmae_profile_begin(mpmae_plan *plan);
int mpmae_profile_report(mpmae_plan *plan, char *buf, int32_t cap);

/* FCMAE.forward: mask -> sparse encoder -> decoder -> heads -> losses.  Asynchronous on `stream`. */
int mpmae_forward(mpmae_plan *plan, const mpmae_io *io, void *cuda_stream);
/* FCMAE.forward_encoder (models/fcmae.py:242-247): mask + sparse encoder only; needs params, workspace,
 * noise, s2_input, mask, flags.  Read the features with mpmae_encoder_features. */
int mpmae_forward_encoder(mpmae_plan *plan, const mpmae_io *io, void *cuda_stream);
/* The forward in stages, for the reference's step-wise methods (models/fcmae.py:242-265 forward_encoder /
 * forward_decoder, :267-412 forward_loss).  `stages` is a mask of MPMAE_STAGE_*; stages that are left out read what an
 * earlier call (or the caller) left in the workspace / io buffers:
 *   MASK    noise -> mask, slot tables          (a {0,1} mask passed as noise reproduces itself: ranks are stable)
 *   ENCODER s2_input -> encoder rows            (tap "stage3.block<last>.y", [B*V, C3], visible cells in patch order)
 *   DECODER encoder rows -> pred_pixel / pred_image
 *   LOSS    pred_pixel / pred_image + targets + mask -> losses
 * Inference-style calls: mpmae_backward is only defined after a full mpmae_forward. */
enum { MPMAE_STAGE_MASK = 1, MPMAE_STAGE_ENCODER = 2, MPMAE_STAGE_DECODER = 4, MPMAE_STAGE_LOSS = 8 };
int mpmae_forward_stages(mpmae_plan *plan, const mpmae_io *io, int32_t stages, void *cuda_stream);
/* hand-derived backward of the same step; accumulates into io->grads.  Must follow mpmae_forward
 * on the same workspace. */
int mpmae_backward(mpmae_plan *plan, const mpmae_io *io, void *cuda_stream);

/* The same backward in three parts, in reverse layer order, so the caller can all-reduce the part of the flat gradient
 * buffer that is final while the rest is still being computed (replaces DDP's bucket hooks, main_pretrain.py:306-310):
 *   part 0 = loss seeds + heads + decoder + proj, part 1 = stages 3 and 2, part 2 = stages 1, 0 + patch embedding.
 * Must be called in order 0, 1, 2 after mpmae_forward.  mpmae_backward_part_range gives the [lo, hi) float range of
 * the gradient buffer completed by each part. */
int mpmae_backward_part(mpmae_plan *plan, const mpmae_io *io, int32_t part, void *cuda_stream);
int mpmae_backward_part_range(const mpmae_plan *plan, int32_t part, int64_t *lo, int64_t *hi);

/* dense encoder features [B, C3, G, G] (zeros at masked cells) from the last forward */
int mpmae_encoder_features(mpmae_plan *plan, const mpmae_io *io, float *out_nchw, void *cuda_stream);

/* stand-alone GEMM entry (unit tests / microbench): out[M,N] = a[M,K] . b[N,K]^T (+bias).
 * backend 0 = fp32 SIMT, 1 = tcgen05 3xTF32, 2 = tcgen05 single-pass TF32, 3 = tcgen05 3xBF16 (1 and 3 need scratch of 2*N*ceil32(K) floats, ceil32 = K rounded up to a multiple of 32) */
int mpmae_gemm_rows(int32_t backend, const float *a, const float *b, const float *bias, float *out,
                    int64_t M, int32_t N, int32_t K, float *scratch, void *cuda_stream);

/* stand-alone fused-epilogue GEMM (unit tests / microbench of the epilogue modes the step uses):
 *   mode 0  out = a.b^T + bias (+ resid)
 *   mode 1  out = a.b^T + bias ; out2 = gelu(out) ; colsum[g, n] += out2^2           (pw1 + GELU + GRN statistic)
 *   mode 2  out = a.b^T ; colsum[g, n] += out * aux ; colsum2[n] += out              (decoder dg)
 *   mode 3  out = (a.b^T + kg[n] * gelu(aux2)) * gelu'(aux2) ; colsum2[n] += out    (GELU/GRN backward)
 * group_rows = rows per statistics group (>= M: one group).  scratch: 2*N*ceil32(K) floats (backends 1 and 3). */
typedef struct mpmae_gemm_desc {
  const float *a, *b, *bias, *resid, *aux, *aux2, *kg;
  float *out, *out2, *colsum, *colsum2, *scratch;
  int64_t M;
  int32_t N, K, group_rows;
  /* optional (zero = off).  a_gelu: the A operand is consumed as gelu(a[m,k]) * a_scale[k] (a_scale null = 1), applied by
   * the operand-splitter warps (backends 1, 3) or on load (backend 0) -- pw2 of a sparse block reading the saved
   * pre-activation.  acc_scale [N] (mode 3): out = (a.b^T * acc_scale[n] + kg[n] * gelu(aux2)) * gelu'(aux2).
   * grn_*: mode 0 with a_gelu, backends 1 / 3: the A scale is the batch-global GRN scale, derived in the kernel from
   * grn_gsq[k] = sum_rows h^2: nx = sqrt(gsq) / (mean sqrt(gsq) + grn_eps), scale = 1 + grn_gamma * nx (both written out,
   * with the denominator, for the backward pass); a_scale is then ignored. */
  int32_t a_gelu;
  const float *a_scale, *acc_scale, *grn_gsq, *grn_gamma;
  float *grn_nx, *grn_scale, *grn_denom;
  float grn_eps;
} mpmae_gemm_desc;
int mpmae_gemm_epi(int32_t mode, int32_t backend, const mpmae_gemm_desc *d, void *cuda_stream);

/* stand-alone weight-gradient product (unit tests): dw[N,K] += x[R,N]^T . y[R,K].
 * backend 0 = fp32 SIMT, 1 or 3 = tcgen05 3xTF32, 2 = tcgen05 single-pass TF32 */
int mpmae_gemm_wgrad(int32_t backend, const float *x, const float *y, float *dw, int64_t R, int32_t N, int32_t K,
                     void *cuda_stream);
/* the same with y consumed as gelu(y) when y_gelu != 0 (dW2f = dy^T . gelu(a) without a materialised h): applied by the
 * splitter warps on backends 1 / 3, on load on backend 0; backend 2 has no splitter and refuses it */
int mpmae_gemm_wgrad_act(int32_t backend, const float *x, const float *y, float *dw, int64_t R, int32_t N, int32_t K,
                         int32_t y_gelu, void *cuda_stream);

/* Fused AdamW over flat buffers (torch.optim.AdamW semantics; the reference builds its optimizer at
 * main_pretrain.py:312-320).  decay_mask: one byte per element (null = decay everything);
 * grad_scale_inv multiplies the gradient first (GradScaler unscale, helpers.py:485-497); step >= 1. */
int mpmae_adamw_step(float *params, const float *grads, float *exp_avg, float *exp_avg_sq,
                     const uint8_t *decay_mask, int64_t n, float lr, float beta1, float beta2, float eps,
                     float weight_decay, int64_t step, float grad_scale_inv, void *cuda_stream);

/* The same step with its per-step scalars in DEVICE memory, for a loss-scaler loop without the host sync of
 * torch.cuda.amp.GradScaler.step (helpers.py:470-506): dev_state[0] = factor applied to the gradient (1 / loss scale,
 * times a clipping coefficient), dev_state[1] != 0 = a non-finite gradient was found: parameters and moments are left
 * untouched (the step is skipped), dev_state[2] = 1-based number of this step; lr < 0 = read the learning rate from
 * dev_state[3] (a step captured in a CUDA graph).  n % 4 == 0, 16-byte aligned buffers. */
int mpmae_adamw_step_dev(float *params, const float *grads, float *exp_avg, float *exp_avg_sq,
                         const uint8_t *decay_mask, int64_t n, float lr, float beta1, float beta2, float eps,
                         float weight_decay, const float *dev_state, void *cuda_stream);

/* Dense operators for the reference's dense ConvNeXt-V2 (models/convnextv2.py:59-207, the finetuning / linear-probe network;
 * inference forward).  Activations are channels-last rows [B*H*W, C]; the pointwise / strided convolutions are
 * mpmae_gemm_epi products on these rows.
 *   dense_im2col: out[r, ci*k*k + kh*k + kw] = x[n, ci, oy*s+kh, ox*s+kw], r = (n*Ho+oy)*Wo+ox, columns >= C*k*k zero
 *                 (kpad % 8 == 0; torch weight.reshape(Cout, -1) column order); nchw != 0: x is [B, C, H, W], else [B, H, W, C]
 *   ln_rows:      out = LayerNorm_C(x) (* w + b when w is given), then GELU when gelu != 0
 *   dense_dwconv: depthwise k x k, stride s, zero padding p, torch weight [C, 1, k, k], on [B, H, W, C]; ln != 0: LayerNorm
 *                 (no affine) over the channels of every output pixel
 *   grn_apply:    per-sample GRN (models/norm_layers.py:33-44, eps 1e-4) of h [R, D] given gsq[g, d] = sum_rows h^2 of every
 *                 group of group_rows rows: out = gamma * (h * Nx) + beta + h;  scratch: (2*groups*D + groups) floats */
int mpmae_dense_im2col(const float *x, float *out, int32_t B, int32_t C, int32_t H, int32_t W, int32_t k, int32_t s,
                       int32_t kpad, int32_t nchw, void *cuda_stream);
int mpmae_ln_rows(const float *x, const float *w, const float *b, float *out, int64_t R, int32_t C, float eps, int32_t gelu,
                  void *cuda_stream);
int mpmae_dense_dwconv(const float *x, const float *w, const float *bias, float *out, int32_t B, int32_t H, int32_t W, int32_t C,
                       int32_t k, int32_t s, int32_t p, int32_t ln, float eps, void *cuda_stream);
int mpmae_grn_apply(const float *h, const float *gsq, const float *gamma, const float *beta, float *out, int64_t R, int32_t D,
                    int32_t group_rows, float eps, float *scratch, void *cuda_stream);

/* Step-wise backward for the reference's step-wise surface (models/fcmae.py:242-412: forward_encoder -> forward_decoder ->
 * forward_loss, each autograd-connected there).  The activations of the matching forward stages must be in the workspace.
 *   MPMAE_BWD_LOSS     io->losses, io->grad_out (d total)  ->  dpred_pixel [B*L, npix], dpred_image [B, nimg] (OUT: gradient of
 *                      the total loss wrt the predictions); d log_vars accumulates into io->grads
 *   MPMAE_BWD_DECODER  dpred_pixel, dpred_image (IN)  ->  head / decoder / proj / mask-token gradients accumulate into io->grads,
 *                      d_x3 [B*V, dims[3]] (OUT: gradient wrt the encoder output rows, visible cells in ascending patch order)
 *   MPMAE_BWD_ENCODER  d_x3 (IN)  ->  encoder gradients accumulate into io->grads */
enum { MPMAE_BWD_LOSS = 0, MPMAE_BWD_DECODER = 1, MPMAE_BWD_ENCODER = 2 };
int mpmae_backward_step(mpmae_plan *plan, const mpmae_io *io, int32_t which, float *dpred_pixel, float *dpred_image,
                        float *d_x3, void *cuda_stream);

/* The loader's per-sample transform (mmearth_dataset.py:58-153) for a whole batch of arrays in their STORED dtypes, on the
 * device, bit-identical to MMEarthDataset.__getitem__ (float64 arithmetic, one rounding to float32):
 *   v = src[n, band[b], j]  ->  lut[v] (label maps; NaN = ignore)  |  NaN where v == nodata
 *   ->  (v - mean[set][b]) / std[set][b] with set = l2a[n] (normalize != 0)
 *   ->  out float32, or int64 with NaN -> -1 (out_int64 != 0).
 * src_type: 0 uint8, 1 uint16, 2 float32.  inner = elements per (sample, band).  All pointers are device pointers. */
#define MPMAE_RAW_MAX_BANDS 16
typedef struct mpmae_raw_desc {
  const void *src;
  void *out;
  const uint8_t *l2a;      /* [B] or null */
  const double *lut;       /* [256] or null */
  int64_t inner;
  int32_t B, src_bands, n_bands, src_type, out_int64, has_nodata, normalize;
  double nodata;
  int32_t band[MPMAE_RAW_MAX_BANDS];
  double mean[2][MPMAE_RAW_MAX_BANDS], std[2][MPMAE_RAW_MAX_BANDS];
} mpmae_raw_desc;
int mpmae_raw_transform(const mpmae_raw_desc *d, void *cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* MPMAE_H */
