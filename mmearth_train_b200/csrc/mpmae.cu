// C ABI of the B200-native MP-MAE (FCMAE) pretraining step: plan, forward, hand-derived backward.
//
// Replaces models/fcmae.py:FCMAE.forward (414-456) + autograd backward of the reference:
//   mask (fcmae.py:214-231) -> SparseConvNeXtV2 (convnextv2_sparse.py:191-220) -> proj + mask token
//   + shared decoder block + heads (fcmae.py:249-265) -> losses (fcmae.py:267-412, custom_loss.py:19-30)
// The step is a fixed launch sequence on the caller's stream; the library allocates no device
// memory, never synchronises and reports errors by return code (include/mpmae.h).
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <iterator>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/mpmae.h"
#include "common.cuh"
#include "dense_ops.cuh"
#include "dwconv.cuh"
#include "dwconv_tiled.cuh"
#include "dwconv_pipe.cuh"
#include "gemm_simt.cuh"
#include "gemm_tc.cuh"
#include "loss.cuh"
#include "misc_kernels.cuh"
#include "optim.cuh"
#include "raw_transform.cuh"
#include "stem.cuh"

using namespace mpmae;

namespace {

thread_local char g_err[512] = "";
int fail(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

struct ParamRef {
  std::string name;
  int64_t shape[4];
  int ndim;
  int64_t off, numel;
  int decay;  // AdamW weight decay applies (timm rule: ndim > 1 and not *.bias)
};
struct BlockP { int64_t dw_k, dw_b, ln_w, ln_b, w1, b1, gamma, beta, w2, b2; };
struct DsP { int64_t ln_w, ln_b, k, b; };
struct Tap { int64_t off_bytes, rows, cols; };

// A weight matrix as the GEMMs consume it, produced once per step by fold_slot(): Wf [N, K] and WfT [K, N] (LayerNorm /
// GRN affines folded in), each as a TF32-exact high part + remainder when the backend is 3xTF32, the folded bias, and
// (zeroed at the start of backward) the gradient of the folded weight / bias.
struct WSlot { int64_t wf, wf_lo, wft, wft_lo, bf, dwf, dbf; int N, K; };
// per-block saved activations / statistics (float offsets into the workspace)
struct BlockW { int64_t vhat, rstd, a, h, g, y, gsq, nx, scale, denom, dsv, cnt_b; WSlot s1, s2; };

}  // namespace

struct mpmae_plan {
  mpmae_cfg cfg;
  Geo geo;
  int S, Ppre, s_stem, P[4], D;
  int64_t R[4], Rpre, cells;
  std::vector<ParamRef> params;
  int64_t n_params = 0;
  int64_t ic_k, ic_b, ic_lnw, ic_lnb, st_k, st_b, st_lnw, st_lnb;
  DsP ds[3];
  std::vector<BlockP> blk[4];
  int64_t proj_w, proj_b, tok;
  std::vector<BlockP> dec;
  int64_t pixw = -1, pixb = -1, imgw = -1, imgb = -1, lnt_w = -1, lnt_b = -1, logv = -1;
  int npix = 0, nimg = 0;
  int col_off[MPMAE_MAX_MOD], col_len[MPMAE_MAX_MOD], is_img[MPMAE_MAX_MOD];
  // workspace
  std::map<std::string, Tap> taps;
  int64_t ws_floats = 0;
  int64_t o_slot, o_vis, o_chat, o_rstd_c, o_shat, o_rstd_s, o_x0;
  int64_t o_ds_xhat[3], o_ds_rstd[3], o_ds_out[3];
  std::vector<BlockW> bw[4];
  std::vector<BlockW> dw;
  int64_t o_z, o_xd, o_pooled, o_pool_rstd, o_dpix, o_dimg, o_acc, o_cs_pix, o_cs_img, o_dpooled;
  int64_t o_zero_begin, o_zero_end;  // statistics region cleared at the start of forward
  int64_t o_bzero_begin, o_bzero_end;  // folded-weight gradients + GRN statistic gradients, cleared at the start of backward
  WSlot ds_slot[3], proj_slot, pix_slot;
  int64_t o_wf, o_wft, o_wf_lo, o_wft_lo, o_bf, o_dwf, o_dbf, o_dsv, o_kg;
  int64_t o_g0, o_g1, o_gda, o_gdv, o_gdu;
  int64_t max_wf = 0, max_rc = 0, max_rd = 0, max_n = 0;
  int64_t o_unfoldjobs[3] = {0, 0, 0};   // device copies of the deferred un-fold jobs of each backward part
  const void *unfold_key[3][3] = {};
  int unfold_njobs[3] = {0, 0, 0}, unfold_ctas[3] = {0, 0, 0};
  int64_t o_foldjobs = 0;                 // device copy of the batched parameter-only fold jobs + tile prefix
  const void *fold_key_params = nullptr, *fold_key_ws = nullptr;
  int fold_njobs = 0, fold_tiles = 0;
  int launches_fwd = 0, launches_bwd = 0;
  int bwd_flip = 0;   // which ping-pong buffer holds the running activation gradient between backward parts
  // optional per-launch CUDA-event profile (bench.py roofline leg)
  bool prof_on = false;
  std::vector<cudaEvent_t> prof_ev;
  std::vector<const char *> prof_name;
  std::vector<double> prof_bytes, prof_flops;
  int prof_n = 0;
};

namespace {

// The fold / un-fold job tables live INSIDE the caller's workspace (the library owns no device memory) and are uploaded
// only when their cache key changes.  A workspace may be shared by several plans (one per batch size); their layouts
// differ, so a call of another plan overwrites this plan's tables with activations.  The registry remembers which plan
// last ran on a workspace: a plan that finds another owner forgets its keys and uploads again.
std::mutex g_ws_mu;
std::map<const void *, const mpmae_plan *> g_ws_owner;
void claim_workspace(mpmae_plan *pl, const void *ws) {
  std::lock_guard<std::mutex> lk(g_ws_mu);
  const mpmae_plan *&owner = g_ws_owner[ws];
  if (owner != pl) {
    owner = pl;
    pl->fold_key_params = pl->fold_key_ws = nullptr;
    pl->fold_njobs = 0;
    for (int i = 0; i < 3; ++i) {
      pl->unfold_njobs[i] = 0;
      for (int j = 0; j < 3; ++j) pl->unfold_key[i][j] = nullptr;
    }
  }
}
void release_workspaces(const mpmae_plan *pl) {
  std::lock_guard<std::mutex> lk(g_ws_mu);
  for (auto it = g_ws_owner.begin(); it != g_ws_owner.end();) it = (it->second == pl) ? g_ws_owner.erase(it) : std::next(it);
}

int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }
constexpr int kMaxFoldJobs = 192;   // base: 36 blocks x (pw1, pw2) + downsamples + decoder + heads = 79

int64_t add_param(mpmae_plan *pl, const std::string &name, std::initializer_list<int64_t> shape) {
  ParamRef r;
  r.name = name;
  r.ndim = (int)shape.size();
  r.numel = 1;
  int i = 0;
  for (int64_t s : shape) { r.shape[i++] = s; r.numel *= s; }
  for (; i < 4; ++i) r.shape[i] = 1;
  r.off = pl->n_params;
  const bool is_bias = name.size() >= 5 && name.compare(name.size() - 5, 5, ".bias") == 0;
  r.decay = (r.ndim > 1 && !is_bias) ? 1 : 0;
  pl->n_params = align_up(pl->n_params + r.numel, 4);
  pl->params.push_back(r);
  return r.off;
}

int64_t ws_alloc(mpmae_plan *pl, const char *name, int64_t rows, int64_t cols) {
  const int64_t off = pl->ws_floats;
  pl->ws_floats = align_up(pl->ws_floats + rows * cols, 64);  // 256-byte granularity
  if (name) pl->taps[name] = Tap{off * 4, rows, cols};
  return off;
}

void build_params(mpmae_plan *pl) {
  const mpmae_cfg &c = pl->cfg;
  const int *dm = c.dims;
  char nm[160];
  pl->ic_k = add_param(pl, "encoder.initial_conv.0.kernel", {9, c.in_chans, dm[0]});
  pl->ic_b = add_param(pl, "encoder.initial_conv.0.bias", {1, dm[0]});
  pl->ic_lnw = add_param(pl, "encoder.initial_conv.1.ln.weight", {dm[0]});
  pl->ic_lnb = add_param(pl, "encoder.initial_conv.1.ln.bias", {dm[0]});
  pl->st_k = add_param(pl, "encoder.stem.0.kernel", {pl->s_stem * pl->s_stem, dm[0]});
  pl->st_b = add_param(pl, "encoder.stem.0.bias", {1, dm[0]});
  pl->st_lnw = add_param(pl, "encoder.stem.1.ln.weight", {dm[0]});
  pl->st_lnb = add_param(pl, "encoder.stem.1.ln.bias", {dm[0]});
  for (int i = 0; i < 3; ++i) {
    snprintf(nm, sizeof nm, "encoder.downsample_layers.%d.0.ln.weight", i);
    pl->ds[i].ln_w = add_param(pl, nm, {dm[i]});
    snprintf(nm, sizeof nm, "encoder.downsample_layers.%d.0.ln.bias", i);
    pl->ds[i].ln_b = add_param(pl, nm, {dm[i]});
    snprintf(nm, sizeof nm, "encoder.downsample_layers.%d.1.kernel", i);
    pl->ds[i].k = add_param(pl, nm, {4, dm[i], dm[i + 1]});
    snprintf(nm, sizeof nm, "encoder.downsample_layers.%d.1.bias", i);
    pl->ds[i].b = add_param(pl, nm, {1, dm[i + 1]});
  }
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < c.depths[i]; ++j) {
      const int C = dm[i];
      BlockP b;
      auto N = [&](const char *leaf) {
        snprintf(nm, sizeof nm, "encoder.stages.%d.%d.%s", i, j, leaf);
        return std::string(nm);
      };
      b.dw_k = add_param(pl, N("dwconv.kernel"), {49, C});
      b.dw_b = add_param(pl, N("dwconv.bias"), {1, C});
      b.ln_w = add_param(pl, N("norm.ln.weight"), {C});
      b.ln_b = add_param(pl, N("norm.ln.bias"), {C});
      b.w1 = add_param(pl, N("pwconv1.linear.weight"), {4 * C, C});
      b.b1 = add_param(pl, N("pwconv1.linear.bias"), {4 * C});
      b.gamma = add_param(pl, N("grn.gamma"), {1, 4 * C});
      b.beta = add_param(pl, N("grn.beta"), {1, 4 * C});
      b.w2 = add_param(pl, N("pwconv2.linear.weight"), {C, 4 * C});
      b.b2 = add_param(pl, N("pwconv2.linear.bias"), {C});
      pl->blk[i].push_back(b);
    }
  const int D = c.dec_dim;
  pl->proj_w = add_param(pl, "proj.weight", {D, dm[3], 1, 1});
  pl->proj_b = add_param(pl, "proj.bias", {D});
  pl->tok = add_param(pl, "mask_token", {1, D, 1, 1});
  for (int k = 0; k < c.dec_depth; ++k) {
    BlockP b;
    auto N = [&](const char *leaf) {
      snprintf(nm, sizeof nm, "decoder.%d.%s", k, leaf);  // aliased decoder_dict.<mod>.<k>.* on the host side
      return std::string(nm);
    };
    b.dw_k = add_param(pl, N("dwconv.weight"), {D, 1, 7, 7});
    b.dw_b = add_param(pl, N("dwconv.bias"), {D});
    b.ln_w = add_param(pl, N("norm.weight"), {D});
    b.ln_b = add_param(pl, N("norm.bias"), {D});
    b.w1 = add_param(pl, N("pwconv1.weight"), {4 * D, D});
    b.b1 = add_param(pl, N("pwconv1.bias"), {4 * D});
    b.gamma = add_param(pl, N("grn.gamma"), {1, 1, 1, 4 * D});
    b.beta = add_param(pl, N("grn.beta"), {1, 1, 1, 4 * D});
    b.w2 = add_param(pl, N("pwconv2.weight"), {D, 4 * D});
    b.b2 = add_param(pl, N("pwconv2.bias"), {D});
    pl->dec.push_back(b);
  }
  // heads: weights of all pixel heads contiguous (one [npix, D] matrix), then their biases; same for image heads
  const int p2 = c.patch_size * c.patch_size;
  for (int pass = 0; pass < 4; ++pass) {
    const bool img = pass >= 2, bias = pass & 1;
    for (int m = 0; m < c.n_mod; ++m) {
      const bool m_img = c.mod_kind[m] == MPMAE_IMAGE_CATEGORICAL || c.mod_kind[m] == MPMAE_IMAGE_CONTINUOUS;
      if (m_img != img) continue;
      const int64_t n = m_img ? c.mod_chans[m] : (int64_t)p2 * c.mod_chans[m];
      snprintf(nm, sizeof nm, "pred_dict.#%d.%s", m, bias ? "bias" : "weight");
      int64_t off;
      if (bias) {
        off = add_param(pl, nm, {n});
      } else if (m_img) {
        off = add_param(pl, nm, {n, D});
      } else {
        off = add_param(pl, nm, {n, D, 1, 1});
      }
      // heads are packed without the 4-float alignment so that they form one matrix / one vector
      pl->params.back().off = off;
      if (!bias) {
        if (m_img) { if (pl->imgw < 0) pl->imgw = off; }
        else { if (pl->pixw < 0) pl->pixw = off; }
      } else {
        if (m_img) { if (pl->imgb < 0) pl->imgb = off; }
        else { if (pl->pixb < 0) pl->pixb = off; }
      }
      // undo the alignment padding inside a group (weights: n*D is a multiple of 4; biases may not be)
      pl->n_params = off + n * (bias ? 1 : D);
    }
    pl->n_params = align_up(pl->n_params, 4);
  }
  if (pl->nimg > 0) {
    pl->lnt_w = add_param(pl, "layer_norm_tmp.weight", {D});
    pl->lnt_b = add_param(pl, "layer_norm_tmp.bias", {D});
  }
  if (c.loss_aggr == 1) pl->logv = add_param(pl, "loss_fn.log_vars", {c.n_mod});
}

void note_wf(mpmae_plan *pl, int64_t n, int64_t k) {
  const int64_t need = std::max(bf16_pair_floats(n, (int)k), bf16_pair_floats(k, (int)n));   // either orientation, padded pairs
  if (need > pl->max_wf) pl->max_wf = need;
  if (n > pl->max_n) pl->max_n = n;
  if (k > pl->max_n) pl->max_n = k;
}

WSlot alloc_slot(mpmae_plan *pl, int N, int K) {
  WSlot s{};
  s.N = N; s.K = K;
  s.wf = ws_alloc(pl, nullptr, N, K);
  s.wf_lo = ws_alloc(pl, nullptr, 1, bf16_pair_floats(N, K));    // 3xTF32 remainder [N, K] or the (padded) 3xBF16 pair array
  s.wft = ws_alloc(pl, nullptr, K, N);
  s.wft_lo = ws_alloc(pl, nullptr, 1, bf16_pair_floats(K, N));
  s.bf = ws_alloc(pl, nullptr, 1, N);
  return s;
}
void alloc_slot_grads(mpmae_plan *pl, WSlot &s) {
  s.dwf = ws_alloc(pl, nullptr, s.N, s.K);
  s.dbf = ws_alloc(pl, nullptr, 1, s.N);
}

void build_workspace(mpmae_plan *pl) {
  const mpmae_cfg &c = pl->cfg;
  const int *dm = c.dims;
  const int64_t B = c.batch, L = pl->geo.L, V = pl->geo.V, D = c.dec_dim;
  char nm[96];
  pl->o_slot = ws_alloc(pl, "slot_of", B * L, 1);
  pl->o_vis = ws_alloc(pl, "vis_patch", B * V, 1);
  pl->o_chat = ws_alloc(pl, "initial.chat", pl->Rpre, dm[0]);
  pl->o_rstd_c = ws_alloc(pl, "initial.rstd", pl->Rpre, 1);
  pl->o_shat = ws_alloc(pl, "stem.shat", pl->R[0], dm[0]);
  pl->o_rstd_s = ws_alloc(pl, "stem.rstd", pl->R[0], 1);
  pl->o_x0 = ws_alloc(pl, "stem.out", pl->R[0], dm[0]);
  pl->max_rc = pl->Rpre * dm[0];
  for (int i = 0; i < 4; ++i) {
    const int64_t R = pl->R[i], C = dm[i];
    if (R * C > pl->max_rc) pl->max_rc = R * C;
    if (R * 4 * C > pl->max_rd) pl->max_rd = R * 4 * C;
    if (i > 0) {
      snprintf(nm, sizeof nm, "down%d.xhat", i);
      pl->o_ds_xhat[i - 1] = ws_alloc(pl, nm, pl->R[i - 1], dm[i - 1]);
      snprintf(nm, sizeof nm, "down%d.rstd", i);
      pl->o_ds_rstd[i - 1] = ws_alloc(pl, nm, pl->R[i - 1], 1);
      snprintf(nm, sizeof nm, "down%d.out", i);
      pl->o_ds_out[i - 1] = ws_alloc(pl, nm, R, C);
      pl->ds_slot[i - 1] = alloc_slot(pl, C, 4 * dm[i - 1]);
      note_wf(pl, C, 4 * dm[i - 1]);
    }
    note_wf(pl, 4 * C, C);
    for (int j = 0; j < c.depths[i]; ++j) {
      BlockW w;
      auto N = [&](const char *leaf) { snprintf(nm, sizeof nm, "stage%d.block%d.%s", i, j, leaf); return nm; };
      w.vhat = ws_alloc(pl, N("vhat"), R, C);
      w.rstd = ws_alloc(pl, N("rstd"), R, 1);
      w.a = ws_alloc(pl, N("a"), R, 4 * C);
      w.h = -1;   // never materialised: pw2 / dW2f apply GELU (and the GRN scale) to `a` on the way into the tensor core
      w.g = -1;
      w.y = ws_alloc(pl, N("y"), R, C);
      w.nx = ws_alloc(pl, N("nx"), 1, 4 * C);
      w.scale = ws_alloc(pl, N("scale"), 1, 4 * C);
      w.denom = ws_alloc(pl, N("denom"), 1, 1);
      w.s1 = alloc_slot(pl, 4 * C, C);
      w.s2 = alloc_slot(pl, C, 4 * C);
      pl->bw[i].push_back(w);
    }
  }
  pl->o_z = ws_alloc(pl, "proj.z", B * V, D);
  pl->o_xd = ws_alloc(pl, "decoder.in", pl->cells, D);
  pl->proj_slot = alloc_slot(pl, D, dm[3]);
  pl->pix_slot = alloc_slot(pl, pl->npix > 0 ? pl->npix : 8, D);
  note_wf(pl, D, dm[3]);
  note_wf(pl, 4 * D, D);
  if (pl->cells * D > pl->max_rc) pl->max_rc = pl->cells * D;
  if (pl->cells * 4 * D > pl->max_rd) pl->max_rd = pl->cells * 4 * D;
  for (int k = 0; k < c.dec_depth; ++k) {
    BlockW w;
    auto N = [&](const char *leaf) { snprintf(nm, sizeof nm, "decoder.block%d.%s", k, leaf); return nm; };
    w.vhat = ws_alloc(pl, N("vhat"), pl->cells, D);
    w.rstd = ws_alloc(pl, N("rstd"), pl->cells, 1);
    w.a = ws_alloc(pl, N("a"), pl->cells, 4 * D);
    w.h = ws_alloc(pl, N("h"), pl->cells, 4 * D);
    w.g = ws_alloc(pl, N("g"), pl->cells, 4 * D);
    w.y = ws_alloc(pl, N("y"), pl->cells, D);
    w.nx = ws_alloc(pl, N("nx"), B, 4 * D);
    w.scale = ws_alloc(pl, N("scale"), B, 4 * D);
    w.denom = ws_alloc(pl, N("denom"), B, 1);
    w.s1 = alloc_slot(pl, 4 * D, D);
    w.s2 = alloc_slot(pl, D, 4 * D);
    pl->dw.push_back(w);
  }
  pl->o_pooled = ws_alloc(pl, "pooled", B, D);
  pl->o_pool_rstd = ws_alloc(pl, "pool.rstd", pl->cells, 1);
  pl->o_dpooled = ws_alloc(pl, nullptr, B, D);
  pl->o_dpix = ws_alloc(pl, "dpix", pl->cells, pl->npix > 0 ? pl->npix : 1);
  pl->o_dimg = ws_alloc(pl, "dimg", B, pl->nimg > 0 ? pl->nimg : 1);
  pl->o_cs_pix = ws_alloc(pl, nullptr, 1, pl->npix + 4);
  pl->o_cs_img = ws_alloc(pl, nullptr, 1, pl->nimg + 4);
  if (pl->npix) note_wf(pl, pl->npix, D);
  if (pl->nimg) note_wf(pl, pl->nimg, D);
  // statistics accumulated with atomics in forward: one contiguous region, cleared by one memset
  pl->o_zero_begin = pl->ws_floats;
  pl->o_acc = ws_alloc(pl, "loss.acc", 1, 2 * MPMAE_MAX_MOD);
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < c.depths[i]; ++j) {
      snprintf(nm, sizeof nm, "stage%d.block%d.gsq", i, j);
      pl->bw[i][j].gsq = ws_alloc(pl, nm, 1, 4 * dm[i]);
    }
  for (int k = 0; k < c.dec_depth; ++k) {
    snprintf(nm, sizeof nm, "decoder.block%d.gsq", k);
    pl->dw[k].gsq = ws_alloc(pl, nm, B, 4 * D);
  }
  pl->o_zero_end = pl->ws_floats;
  // cleared by one memset at the start of backward
  pl->o_bzero_begin = pl->ws_floats;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < c.depths[i]; ++j) {
      alloc_slot_grads(pl, pl->bw[i][j].s1);
      alloc_slot_grads(pl, pl->bw[i][j].s2);
      pl->bw[i][j].dsv = ws_alloc(pl, nullptr, 1, 4 * dm[i]);
      pl->bw[i][j].cnt_b = ws_alloc(pl, nullptr, 1, 1);   // ticket counter of the GRN backward in the pw2 un-fold
    }
  for (int k = 0; k < c.dec_depth; ++k) {
    alloc_slot_grads(pl, pl->dw[k].s1);
    pl->dw[k].dsv = ws_alloc(pl, nullptr, B, 4 * D);
  }
  for (int i = 0; i < 3; ++i) alloc_slot_grads(pl, pl->ds_slot[i]);
  pl->o_bzero_end = pl->ws_floats;
  // weight-fold scratch + backward temporaries
  pl->o_wf = ws_alloc(pl, nullptr, 1, pl->max_wf);
  pl->o_wft = ws_alloc(pl, nullptr, 1, pl->max_wf);
  pl->o_wf_lo = ws_alloc(pl, nullptr, 1, pl->max_wf);
  pl->o_wft_lo = ws_alloc(pl, nullptr, 1, pl->max_wf);
  pl->o_bf = ws_alloc(pl, nullptr, 1, pl->max_n);
  pl->o_dwf = ws_alloc(pl, nullptr, 1, pl->max_wf);
  pl->o_dbf = ws_alloc(pl, nullptr, 1, pl->max_n);
  pl->o_dsv = ws_alloc(pl, nullptr, B, pl->max_n);
  pl->o_kg = ws_alloc(pl, nullptr, B, pl->max_n);
  pl->o_foldjobs = ws_alloc(pl, nullptr, 1, (int64_t)(kMaxFoldJobs * sizeof(FoldArgs) + (kMaxFoldJobs + 1) * sizeof(int) + 3) / 4);
  for (int i = 0; i < 3; ++i) pl->o_unfoldjobs[i] = ws_alloc(pl, nullptr, 1, (int64_t)(64 * sizeof(UnfoldArgs) + 65 * sizeof(int) + 3) / 4);
  pl->o_g0 = ws_alloc(pl, nullptr, 1, pl->max_rc);
  pl->o_g1 = ws_alloc(pl, nullptr, 1, pl->max_rc);
  pl->o_gdv = ws_alloc(pl, nullptr, 1, pl->max_rc);
  pl->o_gdu = ws_alloc(pl, nullptr, 1, pl->max_rc);
  pl->o_gda = ws_alloc(pl, nullptr, 1, pl->max_rd);
}

// ---------------------------------------------------------------------------------------- launch context
struct Ctx {
  mpmae_plan *pl;
  const mpmae_io *io;
  cudaStream_t st;
  float *ws;
  const float *P;  // params
  float *G;        // grads
  int launches = 0;
  std::vector<FoldArgs> *collect = nullptr;   // when set, fold_slot() records the job instead of launching it
  bool folds_done = false;                    // the parameter-only folds were launched as one batch: skip them
  std::vector<UnfoldArgs> deferred;           // un-folds whose results nothing in the backward pass reads
  cudaError_t err = cudaSuccess;
  const char *where = "";
  double pend_bytes = 0, pend_flops = 0;
  float *w(int64_t off) const { return ws + off; }
  const float *p(int64_t off) const { return P + off; }
  float *g(int64_t off) const { return G + off; }
  bool ok() const { return err == cudaSuccess; }
  // algorithmic bytes / flops of the NEXT launch (SURVEY.md 8d accounting; read by the profiler only)
  void acct(double bytes, double flops) { pend_bytes = bytes; pend_flops = flops; }
  void mark(const char *what) {
    if (!pl->prof_on) { pend_bytes = pend_flops = 0; return; }
    if (pl->prof_n >= (int)pl->prof_ev.size()) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      pl->prof_ev.push_back(e);
      pl->prof_name.push_back(what);
      pl->prof_bytes.push_back(0);
      pl->prof_flops.push_back(0);
    }
    pl->prof_name[pl->prof_n] = what;
    pl->prof_bytes[pl->prof_n] = pend_bytes;
    pl->prof_flops[pl->prof_n] = pend_flops;
    cudaEventRecord(pl->prof_ev[pl->prof_n], st);
    ++pl->prof_n;
    pend_bytes = pend_flops = 0;
  }
  void check(cudaError_t e, const char *what, bool kernel = true) {
    if (kernel) ++launches;
    if (err == cudaSuccess && e != cudaSuccess) { err = e; where = what; }
    mark(what);
  }
  void post(const char *what) { check(cudaGetLastError(), what); }
  void zero(float *ptr, int64_t n, const char *what) {
    acct(4.0 * n, 0);
    check(cudaMemsetAsync(ptr, 0, n * sizeof(float), st), "memset", false);
    (void)what;
  }
};

// register-tiled depthwise kernels, generic kernels for the shapes they do not take
// N <= 256 always lands in one N tile of the tcgen05 GEMM (bn = N rounded up to 16) unless shared memory forces a split;
// the launcher re-checks (num_n == 1) and refuses otherwise
bool pick_single_tile(const GemmArgs &a) { return a.N <= 256; }
cudaError_t dw_launch(const DwArgs &d, const int *vis, cudaStream_t st) {
  static const bool no_pipe = getenv("MPMAE_NO_DWPIPE") != nullptr;
  cudaError_t e = no_pipe ? cudaErrorInvalidConfiguration : launch_dwconv_pipe(d, vis, st);
  if (e != cudaErrorInvalidConfiguration) return e;   // the pipelined kernels also produce colsum_out
  (void)cudaGetLastError();
  e = launch_dwconv_tiled(d, vis, st);
  if (e == cudaErrorInvalidConfiguration) { (void)cudaGetLastError(); e = launch_dwconv_fwd(d, st); }
  if (e == cudaSuccess && d.colsum_out) {
    const int64_t R = (int64_t)d.geo.B * d.geo.V * d.P * d.P;
    launch_colsum(d.out, nullptr, d.colsum_out, R, d.C, st);
    e = cudaGetLastError();
  }
  return e;
}
cudaError_t dw_wgrad_launch(const DwWgradArgs &d, const int *vis, cudaStream_t st) {
  static const bool no_pipe = getenv("MPMAE_NO_DWPIPE") != nullptr;
  cudaError_t e = no_pipe ? cudaErrorInvalidConfiguration : launch_dwconv_wgrad_pipe(d, vis, st);
  if (e != cudaErrorInvalidConfiguration) return e;
  (void)cudaGetLastError();
  e = launch_dwconv_wgrad_tiled(d, vis, st);
  if (e == cudaErrorInvalidConfiguration) { (void)cudaGetLastError(); e = launch_dwconv_wgrad(d, st); }
  return e;
}

void fold(Ctx &c, FoldArgs a, const char *what);
int ew_grid(int64_t n4);
bool splitter_backend(Ctx &c);

// single_pass: with the 3xBF16 / 3xTF32 backends, run THIS product as one TF32 pass on the fp32 operands (no operand split):
// the prediction heads, whose outputs feed only the losses (experiment knob MPMAE_HEADS_TF32, see profiles/)
template <int MODE>
void gemm(Ctx &c, const GemmArgs &a_in, const char *what, bool single_pass = false) {
  if (!c.ok() || a_in.M <= 0) return;
  GemmArgs a = a_in;
  mpmae_plan *pl = c.pl;
  const bool use_tc = pl->cfg.gemm_backend != 0 && tc_gemm_supported(MODE, a);
  if (single_pass && use_tc && pl->cfg.gemm_backend != 2) {
    a.Bw_lo = nullptr; a.b16 = 0;
    c.acct(4.0 * ((double)a.M * a.K + (double)a.N * a.K + (double)a.M * a.N), 2.0 * (double)a.M * a.N * a.K);
    c.check(launch_gemm_rows_tc<MODE>(a, 2, c.st), what);
    return;
  }
  if (pl->cfg.gemm_backend == 3) a.b16 = a.Bw_lo ? 1 : 0;
  if (pl->cfg.gemm_backend == 1 || pl->cfg.gemm_backend == 3) {
    // 3xTF32: the weight operand is consumed as a (hi, lo) pair; folded weights were written that way by fold(),
    // raw parameter matrices are split here into the scratch buffers
    if (!a.Bw_lo && use_tc) {
      FoldArgs f{};
      f.W = a.Bw; f.s_n = a.K; f.s_k = 1; f.Wf = c.w(pl->o_wf); f.N = a.N; f.K = a.K; f.SL = a.K;
      fold(c, f, "split_w");
      a.Bw = c.w(pl->o_wf); a.Bw_lo = c.w(pl->o_wf_lo);
      a.b16 = pl->cfg.gemm_backend == 3 ? 1 : 0;
    }
  }
  {
    // SURVEY.md 8(d) accounting: the operands and the tensors this launch materialises / re-reads, each once:
    //   EPI_STORE   A + W + out (+ residual)          EPI_GELU_SQ  A + W + one [M, N] output (a; h only if it is stored)
    //   EPI_DG      A + W + out + h                   EPI_DH_GELU  A + W + out (da) + a
    const double mn = (double)a.M * a.N;
    const double io_mn = MODE == EPI_STORE ? 1 + (a.resid ? 1 : 0) : MODE == EPI_GELU_SQ ? (a.out2 ? 2 : 1) : 2;
    c.acct(4.0 * ((double)a.M * a.K + (double)a.N * a.K + mn * io_mn), 2.0 * mn * a.K);
  }
  if (use_tc) {
    c.check(launch_gemm_rows_tc<MODE>(a, pl->cfg.gemm_backend, c.st), what);
  } else {
    c.check(launch_gemm_rows_simt<MODE>(a, c.st), what);
  }
}
// does gemm<MODE>() take the tcgen05 path for these arguments (the callers that rely on splitter / tail features ask first)
template <int MODE>
bool gemm_uses_tc(Ctx &c, const GemmArgs &a) { return c.pl->cfg.gemm_backend != 0 && tc_gemm_supported(MODE, a); }
bool splitter_backend(Ctx &c) { return c.pl->cfg.gemm_backend == 1 || c.pl->cfg.gemm_backend == 3; }
void wgrad(Ctx &c, const WgradArgs &a_in, const char *what) {
  if (!c.ok()) return;
  WgradArgs a = a_in;
  const bool use_tc = c.pl->cfg.gemm_backend != 0 && tc_wgrad_supported(a);
  const bool exact = a.exact != 0 && splitter_backend(c);
  if (a.y_gelu && use_tc && !exact) {
    // tensor-core weight gradient without splitter warps: materialise gelu(Y) in the (still unused) da scratch
    float *h = c.w(c.pl->o_gda);
    c.acct(4.0 * 2.0 * (double)a.R * a.K, 0);
    pdl(gelu_scale_rows_kernel, ew_grid(a.R * (a.K / 4)), 256, 0, c.st)(a.Y, (const float *)nullptr, h, a.R, a.K);
    c.post("gelu_rows");
    a.Y = h; a.y_gelu = 0;
  }
  c.acct(4.0 * ((double)a.R * a.N + (double)a.R * a.K + (double)a.N * a.K), 2.0 * (double)a.R * a.N * a.K);
  if (use_tc) {
    c.check(launch_gemm_wgrad_tc(a, exact, c.st), what);
    if (a.db && c.ok()) {   // bias gradient = column sums of X (the tensor-core kernel produces dW only)
      c.acct(4.0 * (double)a.R * a.N, 0);
      launch_colsum(a.X, a.rs, a.db, a.R, a.N, c.st);
      c.post("colsum");
    }
  } else {
    c.check(launch_gemm_wgrad(a, c.st), what);
  }
}
void fold(Ctx &c, FoldArgs a, const char *what) {
  if (!c.ok()) return;
  if (c.pl->cfg.gemm_backend == 1 || c.pl->cfg.gemm_backend == 3) {
    if (a.Wf == c.w(c.pl->o_wf)) a.Wf_lo = c.w(c.pl->o_wf_lo);
    if (a.WfT == c.w(c.pl->o_wft)) a.WfT_lo = c.w(c.pl->o_wft_lo);
    a.b16 = c.pl->cfg.gemm_backend == 3 ? 1 : 0;
  }
  launch_fold(a, c.st);
  c.post(what);
}
// fold a weight into its slot: both orientations (+ hi/lo split for 3xTF32) and the folded bias in one launch
void fold_slot(Ctx &c, FoldArgs a, const WSlot &s, const char *what) {
  if (!c.ok()) return;
  const bool split = c.pl->cfg.gemm_backend == 1 || c.pl->cfg.gemm_backend == 3;
  a.b16 = c.pl->cfg.gemm_backend == 3 ? 1 : 0;
  a.Wf = c.w(s.wf); a.WfT = c.w(s.wft);
  a.Wf_lo = split ? c.w(s.wf_lo) : nullptr;
  a.WfT_lo = split ? c.w(s.wft_lo) : nullptr;
  if (a.bias || a.shift_k) a.bf = c.w(s.bf);
  a.N = s.N; a.K = s.K;
  const bool param_only = a.gsq == nullptr && a.scale_n == nullptr;   // depends on parameters alone
  if (param_only && c.collect) { c.collect->push_back(a); return; }
  if (param_only && c.folds_done) return;
  launch_fold(a, c.st);
  c.post(what);
}
// point a GEMM at a slot: out = A . Wf^T (transposed = false) or A . Wf (transposed = true)
void use_slot(Ctx &c, GemmArgs &g, const WSlot &s, bool transposed) {
  const bool split = c.pl->cfg.gemm_backend == 1 || c.pl->cfg.gemm_backend == 3;
  g.b16 = c.pl->cfg.gemm_backend == 3 ? 1 : 0;
  g.Bw = c.w(transposed ? s.wft : s.wf);
  g.Bw_lo = split ? c.w(transposed ? s.wft_lo : s.wf_lo) : nullptr;
}
void unfold(Ctx &c, UnfoldArgs a, const char *what, bool deferrable = false) {
  if (!c.ok()) return;
  if (deferrable && c.deferred.size() < 64) { c.deferred.push_back(a); return; }
  pdl(unfold_kernel, a.K, 256, 0, c.st)(a);
  c.post(what);
}
// the deferred un-folds of this call in one launch; slot selects the cached job table (0..2 = backward part)
void flush_unfolds(Ctx &c, int slot) {
  const int n = (int)c.deferred.size();
  if (n == 0 || !c.ok()) return;
  mpmae_plan *pl = c.pl;
  std::vector<int> start(n + 1, 0);
  for (int j = 0; j < n; ++j) start[j + 1] = start[j] + c.deferred[j].K;
  UnfoldArgs *d_jobs = reinterpret_cast<UnfoldArgs *>(c.w(pl->o_unfoldjobs[slot]));
  int *d_start = reinterpret_cast<int *>(d_jobs + 64);
  const void *key[3] = {c.P, c.G, c.ws};
  if (pl->unfold_key[slot][0] != key[0] || pl->unfold_key[slot][1] != key[1] || pl->unfold_key[slot][2] != key[2] ||
      pl->unfold_njobs[slot] != n || pl->unfold_ctas[slot] != start[n]) {
    c.check(cudaMemcpyAsync(d_jobs, c.deferred.data(), n * sizeof(UnfoldArgs), cudaMemcpyHostToDevice, c.st), "memcpy", false);
    c.check(cudaMemcpyAsync(d_start, start.data(), (n + 1) * sizeof(int), cudaMemcpyHostToDevice, c.st), "memcpy", false);
    for (int i = 0; i < 3; ++i) pl->unfold_key[slot][i] = key[i];
    pl->unfold_njobs[slot] = n; pl->unfold_ctas[slot] = start[n];
  }
  if (c.ok()) {
    pdl(unfold_batch_kernel, start[n], 256, 0, c.st)(d_jobs, d_start, n);
    c.post("unfold_batch");
  }
  c.deferred.clear();
}
int ew_grid(int64_t n4) {
  int64_t g = cdiv64(n4, 256);
  return (int)(g > 148 * 16 ? 148 * 16 : (g < 1 ? 1 : g));
}

// ---- fold jobs that depend on parameters alone (shared by the call sites and the batched launch)
FoldArgs pw1_fold_args(Ctx &c, const BlockP &bp, int C) {
  FoldArgs f{};
  f.W = c.p(bp.w1); f.s_n = C; f.s_k = 1; f.scale_k = c.p(bp.ln_w); f.shift_k = c.p(bp.ln_b);
  f.bias = c.p(bp.b1); f.SL = C;
  return f;
}
// sparse pw2: the raw weight (split for the tensor core) and b2f = b2 + W2 . beta; the GRN scale is NOT folded into the
// weight any more (it would make the fold depend on this step's statistic): it multiplies the A operand instead
FoldArgs sparse_pw2_fold_args(Ctx &c, const BlockP &bp, int D4) {
  FoldArgs f{};
  f.W = c.p(bp.w2); f.s_n = D4; f.s_k = 1; f.SL = D4; f.shift_k = c.p(bp.beta); f.bias = c.p(bp.b2);
  return f;
}
FoldArgs dense_pw2_fold_args(Ctx &c, const BlockP &bp, int D4) {   // per-sample GRN cannot be folded: plain (split) copy
  FoldArgs f{};
  f.W = c.p(bp.w2); f.s_n = D4; f.s_k = 1; f.SL = D4;
  return f;
}
FoldArgs ds_fold_args(Ctx &c, int i, int Ci, int Co) {
  FoldArgs f{};
  f.W = c.p(c.pl->ds[i].k); f.s_n = 1; f.s_k = Co; f.scale_k = c.p(c.pl->ds[i].ln_w);
  f.shift_k = c.p(c.pl->ds[i].ln_b); f.bias = c.p(c.pl->ds[i].b); f.SL = Ci;
  return f;
}
FoldArgs proj_fold_args(Ctx &c) {
  FoldArgs f{};
  f.W = c.p(c.pl->proj_w); f.s_n = c.pl->cfg.dims[3]; f.s_k = 1; f.SL = c.pl->cfg.dims[3];
  return f;
}
FoldArgs heads_fold_args(Ctx &c) {
  FoldArgs f{};
  f.W = c.p(c.pl->pixw); f.s_n = c.pl->cfg.dec_dim; f.s_k = 1; f.SL = c.pl->cfg.dec_dim;
  return f;
}
void fold_slot(Ctx &c, FoldArgs a, const WSlot &s, const char *what);
// All of them in ONE launch at the start of the forward pass (19 launches otherwise at cfg2)
void batched_param_folds(Ctx &c, bool encoder_only) {
  mpmae_plan *pl = c.pl;
  const mpmae_cfg &cf = pl->cfg;
  std::vector<FoldArgs> jobs;
  c.collect = &jobs;
  for (int i = 0; i < 4; ++i) {
    if (i > 0) fold_slot(c, ds_fold_args(c, i - 1, cf.dims[i - 1], cf.dims[i]), pl->ds_slot[i - 1], "fold_ds");
    for (int j = 0; j < cf.depths[i]; ++j) {
      fold_slot(c, pw1_fold_args(c, pl->blk[i][j], cf.dims[i]), pl->bw[i][j].s1, "fold_pw1");
      fold_slot(c, sparse_pw2_fold_args(c, pl->blk[i][j], 4 * cf.dims[i]), pl->bw[i][j].s2, "split_pw2");
    }
  }
  if (!encoder_only) {
    fold_slot(c, proj_fold_args(c), pl->proj_slot, "split_proj");
    for (int k = 0; k < cf.dec_depth; ++k) {
      fold_slot(c, pw1_fold_args(c, pl->dec[k], cf.dec_dim), pl->dw[k].s1, "fold_pw1");
      fold_slot(c, dense_pw2_fold_args(c, pl->dec[k], 4 * cf.dec_dim), pl->dw[k].s2, "split_pw2");
    }
    if (pl->npix > 0) fold_slot(c, heads_fold_args(c), pl->pix_slot, "split_heads");
  }
  c.collect = nullptr;
  const int n = (int)jobs.size();
  if (n == 0 || n > kMaxFoldJobs || !c.ok()) return;
  std::vector<int> start(n + 1, 0);
  for (int j = 0; j < n; ++j) start[j + 1] = start[j] + cdiv(jobs[j].K, 32) * cdiv(jobs[j].N, 32);
  FoldArgs *d_jobs = reinterpret_cast<FoldArgs *>(c.w(pl->o_foldjobs));
  int *d_start = reinterpret_cast<int *>(d_jobs + kMaxFoldJobs);
  if (pl->fold_key_params != (const void *)c.P || pl->fold_key_ws != (const void *)c.ws || pl->fold_njobs != n) {
    c.check(cudaMemcpyAsync(d_jobs, jobs.data(), n * sizeof(FoldArgs), cudaMemcpyHostToDevice, c.st), "memcpy", false);
    c.check(cudaMemcpyAsync(d_start, start.data(), (n + 1) * sizeof(int), cudaMemcpyHostToDevice, c.st), "memcpy", false);
    pl->fold_key_params = c.P; pl->fold_key_ws = c.ws; pl->fold_njobs = n; pl->fold_tiles = start[n];
  }
  if (!c.ok()) return;
  pdl(fold_batch_kernel, pl->fold_tiles, dim3(32, 8), 0, c.st)(d_jobs, d_start, n);
  c.post("fold_batch");
  c.folds_done = true;
}

// One ConvNeXt-V2 block forward (sparse: convnextv2_sparse.py:47-56; dense decoder: convnextv2.py:42-55).
//   dense = per-sample GRN over L cells (eps 1e-4), torch conv weight layout; sparse = batch-global GRN (eps 1e-6)
void block_forward(Ctx &c, const BlockP &bp, const BlockW &bw, const float *x, int64_t R, int C, int P, bool dense) {
  mpmae_plan *pl = c.pl;
  const int D4 = 4 * C;
  DwArgs d{};
  d.x = x; d.w = c.p(bp.dw_k); d.bias = c.p(bp.dw_b); d.resid = nullptr;
  if (dense) { d.w_sc = 49; d.w_skh = 7; d.w_skw = 1; } else { d.w_sc = 1; d.w_skh = C; d.w_skw = 7 * C; }
  d.out = c.w(bw.vhat); d.rstd = c.w(bw.rstd);
  d.slot_of = dense ? nullptr : reinterpret_cast<const int *>(c.w(pl->o_slot));
  d.geo = pl->geo;
  if (dense) d.geo.V = pl->geo.L;
  d.P = P; d.C = C; d.flip = 0; d.do_ln = 1; d.eps = 1e-6f;
  c.acct(4.0 * (2.0 * R * C + R + 50.0 * C), 2.0 * 49 * (double)R * C);
  if (c.ok()) c.check(dw_launch(d, reinterpret_cast<const int *>(c.w(pl->o_vis)), c.st), "dwconv_fwd");

  fold_slot(c, pw1_fold_args(c, bp, C), bw.s1, "fold_pw1");

  const int group_rows = dense ? pl->geo.L : (int)(R > 0x7fffffff ? 0x7fffffff : R);
  const int groups = dense ? pl->geo.B : 1;
  GemmArgs g1{};
  g1.A = c.w(bw.vhat); use_slot(c, g1, bw.s1, false); g1.bias = c.w(bw.s1.bf);
  g1.out = c.w(bw.a); g1.out2 = dense ? c.w(bw.h) : nullptr; g1.colsum = c.w(bw.gsq);
  g1.M = R; g1.N = D4; g1.K = C; g1.group_rows = group_rows;
  gemm<EPI_GELU_SQ>(c, g1, "pw1");

  if (dense && c.ok()) {   // per-sample statistic; the batch-global one of the sparse blocks is derived inside pw2
    pdl(grn_scale_kernel, groups, 256, 0, c.st)(c.w(bw.gsq), c.p(bp.gamma), c.w(bw.nx), c.w(bw.scale), c.w(bw.denom),
                                               D4, 1e-4f);
    c.post("grn_scale");
  }
  GemmArgs g2{};
  g2.out = c.w(bw.y); g2.resid = x; g2.M = R; g2.N = C; g2.K = D4; g2.group_rows = group_rows;
  if (dense) {
    if (c.ok()) {
      c.acct(4.0 * 2.0 * R * D4, 0);
      pdl(grn_apply_kernel, ew_grid(R * (D4 / 4)), 256, 0, c.st)(c.w(bw.h), c.w(bw.scale), c.p(bp.beta), c.w(bw.g), R, D4,
                                                               group_rows);
      c.post("grn_apply");
    }
    fold_slot(c, dense_pw2_fold_args(c, bp, D4), bw.s2, "split_pw2");
    g2.A = c.w(bw.g); g2.bias = c.p(bp.b2);
  } else {
    // GRN is affine in h per channel: y = x + (gelu(a) * s) . W2^T + (b2 + W2 . beta) with s = 1 + gamma * Nx.  The weight
    // split and the bias are parameter-only (batched fold); gelu and s are applied to the A operand by the splitter warps.
    fold_slot(c, sparse_pw2_fold_args(c, bp, D4), bw.s2, "split_pw2");
    g2.bias = c.w(bw.s2.bf);
    g2.A = c.w(bw.a); g2.a_gelu = 1;
    if (gemm_uses_tc<EPI_STORE>(c, g2) && splitter_backend(c)) {
      // the kernel derives the batch-global GRN scale from the finished statistic in its prologue (no launch, no pass)
      g2.grn_gsq = c.w(bw.gsq); g2.grn_gamma = c.p(bp.gamma); g2.grn_nx = c.w(bw.nx); g2.grn_scale = c.w(bw.scale);
      g2.grn_denom = c.w(bw.denom); g2.grn_eps = 1e-6f;
    } else {
      // no operand-splitter warps on this path: GRN scale by its own kernel, then gelu(a) * s materialised in the backward
      // scratch (unused in forward)
      float *hs = c.w(pl->o_gda);
      if (c.ok()) {
        pdl(grn_scale_kernel, 1, 256, 0, c.st)(c.w(bw.gsq), c.p(bp.gamma), c.w(bw.nx), c.w(bw.scale), c.w(bw.denom), D4, 1e-6f);
        c.post("grn_scale");
        c.acct(4.0 * 2.0 * R * D4, 0);
        pdl(gelu_scale_rows_kernel, ew_grid(R * (D4 / 4)), 256, 0, c.st)(c.w(bw.a), c.w(bw.scale), hs, R, D4);
        c.post("gelu_scale");
      }
      g2.A = hs; g2.a_gelu = 0;
    }
  }
  use_slot(c, g2, bw.s2, false);
  gemm<EPI_STORE>(c, g2, "pw2");
}

// Backward of one block.  dy -> dx (both [R, C]); parameter grads accumulate.
// dy_colsum_done: sum_rows dy was already accumulated into bw.s2.dbf by the kernel that produced dy;
// dx_colsum: where the column sums of dx go (the bias gradient of the layer that consumes dx), or null
void block_backward(Ctx &c, const BlockP &bp, const BlockW &bw, const float *x, const float *dy, float *dx, int64_t R,
                    int C, int P, bool dense, bool dy_colsum_done = false, float *dx_colsum = nullptr) {
  mpmae_plan *pl = c.pl;
  const int D4 = 4 * C;
  const int group_rows = dense ? pl->geo.L : (int)(R > 0x7fffffff ? 0x7fffffff : R);
  const int groups = dense ? pl->geo.B : 1;
  float *da = c.w(pl->o_gda);
  float *dsv = c.w(bw.dsv), *kg = c.w(pl->o_kg);
  float *dwf1 = c.w(bw.s1.dwf), *dbf1 = c.w(bw.s1.dbf);

  if (dense) {
    // dg = dy . W2 ; A[g,d] = sum dg*h ; dbeta = sum dg
    GemmArgs gd{};
    gd.A = dy; use_slot(c, gd, bw.s2, true); gd.out = da; gd.aux = c.w(bw.h); gd.colsum = dsv; gd.colsum2 = c.g(bp.beta);
    gd.M = R; gd.N = D4; gd.K = C; gd.group_rows = group_rows;
    gemm<EPI_DG>(c, gd, "dg");
    // dW2 += dy^T . g ; db2 += sum dy
    WgradArgs wg{};
    wg.X = dy; wg.Y = c.w(bw.g); wg.dW = c.g(bp.w2); wg.db = c.g(bp.b2); wg.R = R; wg.N = C; wg.K = D4;
    wgrad(c, wg, "dW2");
    if (c.ok()) {
      pdl(grn_bwd_scale_kernel, groups, 256, 0, c.st)(dsv, c.w(bw.nx), c.w(bw.denom), c.p(bp.gamma), c.g(bp.gamma), kg, D4);
      c.post("grn_bwd_scale");
      c.acct(4.0 * 4.0 * R * D4, 0);
      pdl(grn_gelu_bwd_kernel, ew_grid(R * (D4 / 4)), 256, 0, c.st)(da, c.w(bw.h), c.w(bw.a), c.w(bw.scale), kg, da, R, D4,
                                                                  group_rows);
      c.post("grn_gelu_bwd");
      launch_colsum(da, nullptr, dbf1, R, D4, c.st);   // db1f = sum da
      c.post("colsum");
    }
  } else {
    // pw2 with h' = gelu(a) * s:  dW2f = dy^T . gelu(a), db2f = sum dy  ->  dW2 = dW2f * s + db2f (x) beta, ds (= A) =
    // sum_n dW2f . W2, dbeta = W2^T db2f, db2 = db2f: the chain rule of (scale, shift) in front of a linear map
    float *dwf2 = c.w(bw.s2.dwf), *dbf2 = c.w(bw.s2.dbf);
    WgradArgs wg{};
    wg.X = dy; wg.Y = c.w(bw.a); wg.y_gelu = 1; wg.dW = dwf2; wg.db = dy_colsum_done ? nullptr : dbf2; wg.R = R; wg.N = C; wg.K = D4;
    wg.exact = 1;   // ds (GRN statistic gradient) is derived from dW2f and feeds every row's gradient
    wgrad(c, wg, "dW2f");
    UnfoldArgs u{};
    u.W = c.p(bp.w2); u.s_n = D4; u.s_k = 1; u.scale_k = c.w(bw.scale); u.shift_k = c.p(bp.beta);
    u.dWf = dwf2; u.dbf = dbf2; u.dW = c.g(bp.w2); u.dscale = dsv; u.dshift = c.g(bp.beta); u.dbias = c.g(bp.b2);
    u.N = C; u.K = D4; u.SL = D4;
    // da = ((dy . W2) * s + kg * gelu(a)) * gelu'(a) ; db1f = sum da rides on the epilogue
    GemmArgs gd{};
    gd.A = dy; use_slot(c, gd, bw.s2, true); gd.out = da; gd.aux = nullptr; gd.aux2 = c.w(bw.a); gd.kg = kg;
    gd.acc_scale = c.w(bw.scale);
    gd.colsum2 = dbf1;
    gd.M = R; gd.N = D4; gd.K = C; gd.group_rows = group_rows;
    if (gemm_uses_tc<EPI_DH_GELU>(c, gd)) {
      // plain un-fold; the backward of the GRN statistic (dgamma, kg) happens in the prologue of the da kernel
      unfold(c, u, "unfold_pw2");
      gd.kg = nullptr;
      gd.grnb_ds = dsv; gd.grnb_nx = c.w(bw.nx); gd.grnb_denom = c.w(bw.denom); gd.grnb_gamma = c.p(bp.gamma);
      gd.grnb_dgamma = c.g(bp.gamma);
    } else if (c.ok()) {   // un-fold + (last CTA) backward of the GRN statistic: dgamma, kg
      GrnBwdArgs q{c.w(bw.nx), c.w(bw.denom), c.p(bp.gamma), c.g(bp.gamma), kg, reinterpret_cast<unsigned int *>(c.w(bw.cnt_b))};
      pdl(unfold_grn_kernel, u.K, 256, 0, c.st)(u, q);
      c.post("unfold_pw2");
    }
    gemm<EPI_DH_GELU>(c, gd, "da");
  }
  // pw1 (LN affine folded): dW1f = da^T . vhat
  WgradArgs w1{};
  w1.X = da; w1.Y = c.w(bw.vhat); w1.dW = dwf1; w1.db = nullptr; w1.R = R; w1.N = D4; w1.K = C;
  wgrad(c, w1, "dW1f");
  UnfoldArgs u1{};
  u1.W = c.p(bp.w1); u1.s_n = C; u1.s_k = 1; u1.scale_k = c.p(bp.ln_w); u1.shift_k = c.p(bp.ln_b);
  u1.dWf = dwf1; u1.dbf = dbf1; u1.dW = c.g(bp.w1); u1.dscale = c.g(bp.ln_w); u1.dshift = c.g(bp.ln_b);
  u1.dbias = c.g(bp.b1); u1.N = D4; u1.K = C; u1.SL = C;
  unfold(c, u1, "unfold_pw1", true);
  // dvhat = da . W1f
  float *dv = c.w(pl->o_gdv), *du = c.w(pl->o_gdu);
  GemmArgs gv{};
  gv.A = da; use_slot(c, gv, bw.s1, true); gv.out = dv; gv.M = R; gv.N = C; gv.K = D4; gv.group_rows = group_rows;
  {
    GemmArgs gl = gv;   // LayerNorm backward fused into the epilogue when the whole row sits in one tile
    gl.out = du; gl.ln_xhat = c.w(bw.vhat); gl.ln_rstd = c.w(bw.rstd);
    static const bool no_fuse = getenv("MPMAE_NO_LNFUSE") != nullptr;
    if (!no_fuse && pl->cfg.gemm_backend != 0 && tc_gemm_supported(EPI_STORE, gl) && tc_ln_bwd_ok(gl) && pick_single_tile(gl)) {
      gemm<EPI_STORE>(c, gl, "dvhat_ln");
    } else {
      gemm<EPI_STORE>(c, gv, "dvhat");
      if (c.ok()) {
        c.acct(4.0 * (3.0 * R * C + R), 0);
        launch_ln_rows_bwd(dv, c.w(bw.vhat), c.w(bw.rstd), nullptr, du, R, C, c.st);
        c.post("ln_bwd");
      }
    }
  }
  // depthwise: dx = flipped stencil over du + dy (residual) ; dW, db
  DwArgs d{};
  d.x = du; d.w = c.p(bp.dw_k); d.bias = nullptr; d.resid = dy;
  if (dense) { d.w_sc = 49; d.w_skh = 7; d.w_skw = 1; } else { d.w_sc = 1; d.w_skh = C; d.w_skw = 7 * C; }
  d.out = dx; d.rstd = nullptr;
  d.slot_of = dense ? nullptr : reinterpret_cast<const int *>(c.w(pl->o_slot));
  d.geo = pl->geo;
  if (dense) d.geo.V = pl->geo.L;
  d.P = P; d.C = C; d.flip = 1; d.do_ln = 0; d.eps = 0.f;
  d.colsum_out = dense ? nullptr : dx_colsum;
  DwWgradArgs dwg{};
  dwg.x = x; dwg.du = du; dwg.dw = c.g(bp.dw_k); dwg.w_skh = d.w_skh; dwg.w_skw = d.w_skw; dwg.w_sc = d.w_sc;
  dwg.dbias = c.g(bp.dw_b); dwg.slot_of = d.slot_of; dwg.geo = d.geo; dwg.P = P; dwg.C = C;
  bool merged = false;
  if (c.ok() && !dense) {   // dX and dW from one read of the du halo window (patch stages)
    static const bool no_merge = getenv("MPMAE_NO_DWMERGE") != nullptr;
    pipe::DwBwdArgs mb{d, x, dwg.dw, dwg.dbias, reinterpret_cast<const int *>(c.w(pl->o_vis))};
    cudaError_t e = no_merge ? cudaErrorInvalidConfiguration : launch_dwconv_bwd_pipe(mb, c.st);
    if (e != cudaErrorInvalidConfiguration) {
      c.acct(4.0 * (4.0 * R * C + 99.0 * C), 2.0 * 99 * (double)R * C);
      c.check(e, "dwconv_bwd");
      merged = true;
    } else {
      (void)cudaGetLastError();
    }
  }
  if (!merged) {
    c.acct(4.0 * (3.0 * R * C + 49.0 * C), 2.0 * 49 * (double)R * C);
    if (c.ok()) c.check(dw_launch(d, reinterpret_cast<const int *>(c.w(pl->o_vis)), c.st), "dwconv_dx");
    c.acct(4.0 * (2.0 * R * C + 50.0 * C), 2.0 * 50 * (double)R * C);
    if (c.ok()) c.check(dw_wgrad_launch(dwg, reinterpret_cast<const int *>(c.w(pl->o_vis)), c.st), "dwconv_wgrad");
  }
}

InitConvArgs init_args(Ctx &c) {
  mpmae_plan *pl = c.pl;
  InitConvArgs a{};
  a.img = c.io->s2_input; a.kernel = c.p(pl->ic_k); a.bias = c.p(pl->ic_b);
  a.chat = c.w(pl->o_chat); a.rstd = c.w(pl->o_rstd_c);
  a.slot_of = reinterpret_cast<const int *>(c.w(pl->o_slot));
  a.vis_patch = reinterpret_cast<const int *>(c.w(pl->o_vis));
  a.flags = c.io->flags;
  a.geo = pl->geo; a.S = pl->S; a.Cin = pl->cfg.in_chans; a.C0 = pl->cfg.dims[0]; a.Ppre = pl->Ppre; a.eps = 1e-6f;
  return a;
}
StemArgs stem_args(Ctx &c) {
  mpmae_plan *pl = c.pl;
  StemArgs s{};
  s.chat = c.w(pl->o_chat); s.rstd_c = c.w(pl->o_rstd_c);
  s.ln0_w = c.p(pl->ic_lnw); s.ln0_b = c.p(pl->ic_lnb); s.kernel = c.p(pl->st_k); s.bias = c.p(pl->st_b);
  s.ln1_w = c.p(pl->st_lnw); s.ln1_b = c.p(pl->st_lnb);
  s.shat = c.w(pl->o_shat); s.rstd_s = c.w(pl->o_rstd_s); s.x0 = c.w(pl->o_x0);
  s.R0 = pl->R[0]; s.C0 = pl->cfg.dims[0]; s.s2 = pl->s_stem * pl->s_stem; s.eps = 1e-6f;
  return s;
}

LossArgs loss_args(Ctx &c) {
  mpmae_plan *pl = c.pl;
  LossArgs a{};
  a.n_mod = pl->cfg.n_mod;
  for (int m = 0; m < a.n_mod; ++m) {
    a.mod[m].kind = pl->cfg.mod_kind[m];
    a.mod[m].chans = pl->cfg.mod_chans[m];
    a.mod[m].col_off = pl->col_off[m];
    a.mod[m].norm_pix = pl->cfg.mod_norm_pix[m];
    a.mod[m].target = c.io->targets[m];
  }
  a.pred_pix = c.io->pred_pixel; a.npix = pl->npix;
  a.pred_img = c.io->pred_image; a.nimg = pl->nimg;
  a.mask = c.io->mask;
  a.dpix = c.w(pl->o_dpix); a.dimg = c.w(pl->o_dimg); a.acc = c.w(pl->o_acc);
  a.B = pl->geo.B; a.L = pl->geo.L; a.G = pl->geo.G; a.p = pl->cfg.patch_size; a.S = pl->S;
  return a;
}

// forward stages (mpmae.h: MPMAE_STAGE_*)
constexpr int ST_MASK = 1, ST_ENC = 2, ST_DEC = 4, ST_LOSS = 8, ST_ALL = 15;

int check_io(const mpmae_plan *pl, const mpmae_io *io, bool backward, int stages = ST_ALL) {
  if (!pl || !io) return fail(MPMAE_ERR_INVALID, "null plan/io");
  if (!io->params || !io->workspace || !io->mask || !io->flags)
    return fail(MPMAE_ERR_INVALID, "null device pointer in mpmae_io");
  if ((stages & ST_MASK) && !io->noise) return fail(MPMAE_ERR_INVALID, "noise is null");
  if ((stages & ST_ENC) && !io->s2_input) return fail(MPMAE_ERR_INVALID, "s2_input is null");
  if (stages & (ST_DEC | ST_LOSS)) {
    if (pl->npix > 0 && !io->pred_pixel) return fail(MPMAE_ERR_INVALID, "pred_pixel is null");
    if (pl->nimg > 0 && !io->pred_image) return fail(MPMAE_ERR_INVALID, "pred_image is null");
  }
  if (stages & ST_LOSS) {
    if (!io->losses) return fail(MPMAE_ERR_INVALID, "losses is null");
    for (int m = 0; m < pl->cfg.n_mod; ++m)
      if (!io->targets[m]) return fail(MPMAE_ERR_INVALID, "target %d is null", m);
  }
  if (io->workspace_bytes < (size_t)pl->ws_floats * 4) return fail(MPMAE_ERR_WORKSPACE, "workspace too small");
  if ((reinterpret_cast<uintptr_t>(io->workspace) & 255) != 0) return fail(MPMAE_ERR_INVALID, "workspace not 256-byte aligned");
  if ((reinterpret_cast<uintptr_t>(io->params) & 15) != 0) return fail(MPMAE_ERR_INVALID, "params not 16-byte aligned");
  if (backward && (!io->grads || (reinterpret_cast<uintptr_t>(io->grads) & 15) != 0))
    return fail(MPMAE_ERR_INVALID, "grads null or misaligned");
  return MPMAE_OK;
}

int finish(Ctx &c, int *count_slot) {
  if (!c.ok()) return fail(MPMAE_ERR_CUDA, "%s: %s", c.where, cudaGetErrorString(c.err));
  *count_slot = c.launches;
  return MPMAE_OK;
}

}  // namespace

// =================================================================================================
extern "C" {

const char *mpmae_last_error(void) { return g_err; }
int mpmae_version(void) { return 100; }

int mpmae_plan_create(const mpmae_cfg *cfg, mpmae_plan **out) {
  if (!cfg || !out) return fail(MPMAE_ERR_INVALID, "null cfg/out");
  const mpmae_cfg &c = *cfg;
  if (c.batch <= 0 || c.patch_size <= 0 || c.patch_size % 8 != 0 || c.img_size % c.patch_size != 0)
    return fail(MPMAE_ERR_INVALID, "bad geometry: batch %d img %d patch %d", c.batch, c.img_size, c.patch_size);
  if (c.patch_size != 8 && c.patch_size != 16)
    return fail(MPMAE_ERR_UNSUPPORTED, "patch_size %d: kernels are instantiated for 8 and 16", c.patch_size);
  if (c.n_mod <= 0 || c.n_mod > MPMAE_MAX_MOD) return fail(MPMAE_ERR_INVALID, "n_mod %d", c.n_mod);
  if (c.dec_depth < 1 || c.dec_dim % 128 != 0 || c.dec_dim > 1024)
    return fail(MPMAE_ERR_INVALID, "decoder depth/dim (dec_dim must be a multiple of 128, at most 1024)");
  for (int i = 0; i < 4; ++i) {
    if (c.depths[i] < 1 || c.dims[i] % 8 != 0) return fail(MPMAE_ERR_UNSUPPORTED, "dims must be multiples of 8");
    if (i > 0 && c.dims[i] < c.dims[i - 1]) return fail(MPMAE_ERR_UNSUPPORTED, "dims must be non-decreasing");
  }
  if (c.dims[0] > 128) return fail(MPMAE_ERR_UNSUPPORTED, "dims[0] > 128: stem kernels hold <= 4 channels per lane");
  if (c.in_chans < 1 || c.in_chans * (c.dims[0] / 4) > 512)
    return fail(MPMAE_ERR_UNSUPPORTED, "in_chans * dims[0] / 4 > 512: the patch-embedding weight-gradient kernel maps one thread per (ci, 4 co)");
  auto *pl = new mpmae_plan();
  pl->cfg = c;
  pl->S = c.img_size;
  pl->geo.B = c.batch;
  pl->geo.G = c.img_size / c.patch_size;
  pl->geo.L = pl->geo.G * pl->geo.G;
  pl->geo.V = (int)(pl->geo.L * (1.0 - (double)c.mask_ratio));  // int(L * (1 - mask_ratio)), fcmae.py:216-217
  {  // python evaluates L * (1 - mask_ratio) in double with mask_ratio a python float
    const double mr = (double)c.mask_ratio;
    // the float -> double round trip of e.g. 0.6f is 0.60000002384; recover the decimal the caller meant
    const double mr_dec = (double)((long long)(mr * 1e6 + 0.5)) / 1e6;
    pl->geo.V = (int)(pl->geo.L * (1.0 - mr_dec));
  }
  if (pl->geo.V < 1 || pl->geo.V > pl->geo.L) { delete pl; return fail(MPMAE_ERR_INVALID, "no visible patches"); }
  if (pl->geo.L > 1024) { delete pl; return fail(MPMAE_ERR_UNSUPPORTED, "patch grid too large"); }
  pl->Ppre = c.patch_size;
  pl->s_stem = c.patch_size / 8;
  pl->D = c.dec_dim;
  int P = 8;
  for (int i = 0; i < 4; ++i) { pl->P[i] = P; pl->R[i] = (int64_t)c.batch * pl->geo.V * P * P; P /= 2; }
  pl->Rpre = (int64_t)c.batch * pl->geo.V * pl->Ppre * pl->Ppre;
  pl->cells = (int64_t)c.batch * pl->geo.L;
  const int p2 = c.patch_size * c.patch_size;
  for (int m = 0; m < c.n_mod; ++m) {
    const int k = c.mod_kind[m];
    if (k < 0 || k > 3 || c.mod_chans[m] <= 0) { delete pl; return fail(MPMAE_ERR_INVALID, "modality %d", m); }
    pl->is_img[m] = (k == MPMAE_IMAGE_CATEGORICAL || k == MPMAE_IMAGE_CONTINUOUS);
    if (pl->is_img[m]) { pl->col_off[m] = pl->nimg; pl->col_len[m] = c.mod_chans[m]; pl->nimg += c.mod_chans[m]; }
    else { pl->col_off[m] = pl->npix; pl->col_len[m] = p2 * c.mod_chans[m]; pl->npix += p2 * c.mod_chans[m]; }
  }
  build_params(pl);
  build_workspace(pl);
  *out = pl;
  return MPMAE_OK;
}

void mpmae_plan_destroy(mpmae_plan *plan) {
  if (plan) release_workspaces(plan);
  delete plan;
}

int64_t mpmae_param_total(const mpmae_plan *plan) { return plan ? plan->n_params : 0; }
int32_t mpmae_param_count(const mpmae_plan *plan) { return plan ? (int32_t)plan->params.size() : 0; }
int mpmae_param_info(const mpmae_plan *plan, int32_t index, char *name, int32_t name_cap, int64_t shape[4],
                     int32_t *ndim, int64_t *offset) {
  if (!plan || index < 0 || index >= (int32_t)plan->params.size()) return fail(MPMAE_ERR_INVALID, "param index");
  const ParamRef &r = plan->params[index];
  if (name && name_cap > 0) { strncpy(name, r.name.c_str(), name_cap - 1); name[name_cap - 1] = 0; }
  if (shape) for (int i = 0; i < 4; ++i) shape[i] = r.shape[i];
  if (ndim) *ndim = r.ndim;
  if (offset) *offset = r.off;
  return MPMAE_OK;
}
int32_t mpmae_param_decay(const mpmae_plan *plan, int32_t index) {
  if (!plan || index < 0 || index >= (int32_t)plan->params.size()) return -1;
  return plan->params[index].decay;
}

size_t mpmae_workspace_bytes(const mpmae_plan *plan) { return plan ? (size_t)plan->ws_floats * 4 : 0; }
int32_t mpmae_pred_pixel_cols(const mpmae_plan *plan) { return plan ? plan->npix : 0; }
int32_t mpmae_pred_image_cols(const mpmae_plan *plan) { return plan ? plan->nimg : 0; }
int32_t mpmae_pred_col_offset(const mpmae_plan *plan, int32_t mod) {
  if (!plan || mod < 0 || mod >= plan->cfg.n_mod) return -1;
  return plan->col_off[mod];
}
int32_t mpmae_visible_patches(const mpmae_plan *plan) { return plan ? plan->geo.V : 0; }

int mpmae_tap_info(const mpmae_plan *plan, const char *name, int64_t *byte_offset, int64_t *rows, int64_t *cols) {
  if (!plan || !name) return fail(MPMAE_ERR_INVALID, "null");
  auto it = plan->taps.find(name);
  if (it == plan->taps.end()) return fail(MPMAE_ERR_INVALID, "no tap named %s", name);
  if (byte_offset) *byte_offset = it->second.off_bytes;
  if (rows) *rows = it->second.rows;
  if (cols) *cols = it->second.cols;
  return MPMAE_OK;
}
int32_t mpmae_tap_count(const mpmae_plan *plan) { return plan ? (int32_t)plan->taps.size() : 0; }
int mpmae_tap_name(const mpmae_plan *plan, int32_t index, char *name, int32_t name_cap) {
  if (!plan || index < 0 || index >= (int32_t)plan->taps.size() || !name || name_cap <= 0)
    return fail(MPMAE_ERR_INVALID, "tap index");
  auto it = plan->taps.begin();
  std::advance(it, index);
  strncpy(name, it->first.c_str(), name_cap - 1);
  name[name_cap - 1] = 0;
  return MPMAE_OK;
}

int32_t mpmae_launch_count(const mpmae_plan *plan, int32_t backward) {
  if (!plan) return 0;
  return backward ? plan->launches_bwd : plan->launches_fwd;
}

// -------------------------------------------------------------------------------------------------
static int forward_impl(mpmae_plan *pl, const mpmae_io *io, void *cuda_stream, int stages) {
  int rc = check_io(pl, io, false, stages);
  if (rc) return rc;
  const bool encoder_only = !(stages & ST_DEC);
  claim_workspace(pl, io->workspace);
  Ctx c{pl, io, static_cast<cudaStream_t>(cuda_stream), static_cast<float *>(io->workspace), io->params, io->grads};
  c.mark("start");
  const mpmae_cfg &cf = pl->cfg;
  const int *dm = cf.dims;
  const Geo geo = pl->geo;
  int *slot_of = reinterpret_cast<int *>(c.w(pl->o_slot));
  int *vis = reinterpret_cast<int *>(c.w(pl->o_vis));

  c.zero(c.w(pl->o_zero_begin), pl->o_zero_end - pl->o_zero_begin, "zero_stats");
  if (stages & ST_MASK) {
    c.check(cudaMemsetAsync(io->flags, 0, 4 * sizeof(int32_t), c.st), "memset", false);
    pdl(mask_kernel, geo.B, 64, (size_t)geo.L * 8, c.st)(io->noise, io->mask, slot_of, vis, geo.L, geo.V);
    c.post("mask");
  }
  if (stages & (ST_ENC | ST_DEC)) batched_param_folds(c, encoder_only);

  const float *x = c.w(pl->bw[3][cf.depths[3] - 1].y);   // encoder output rows (written by the caller when ST_ENC is off)
  if (stages & ST_ENC) {
  {  // patch embedding: 3x3 conv + LN (+GELU, stem depthwise, LN); the stem is fused into the conv kernel when k = s = 1
    InitConvArgs a = init_args(c);
    StemArgs s = stem_args(c);
    const int f4 = (a.C0 >> 2) / ln_parts(a.C0);
    if (a.C0 % 8 != 0 || f4 < 1 || f4 > 6) return fail(MPMAE_ERR_UNSUPPORTED, "dims[0] = %d: unsupported by the patch-embedding kernel", a.C0);
    const int fuse = (s.s2 == 1) ? 1 : 0;
    const size_t sm = ((size_t)a.Cin * 100 + 9 * (size_t)a.Cin * a.C0 + 64 * (size_t)(a.C0 + 4) + 6 * (size_t)a.C0) * 4;
    static size_t configured = 0;
    if (sm > 48 * 1024 && sm > configured) {
      c.check(cudaFuncSetAttribute(initial_conv_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm), "attr");
      configured = sm;
    }
    const int tiles = (pl->Ppre / 8) * (pl->Ppre / 8);
    const int64_t units = (int64_t)geo.B * geo.V * tiles;
    const int threads = 32 * (a.C0 / 8);
    static int occ_threads = 0, occ_per_sm = 1;
    static size_t occ_sm = 0;
    if (occ_threads != threads || occ_sm != sm) {   // queried once per (block size, shared memory) pair
      int v = 1;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, initial_conv_fwd_kernel, threads, sm) != cudaSuccess || v < 1) {
        (void)cudaGetLastError();
        v = 1;
      }
      occ_threads = threads; occ_sm = sm; occ_per_sm = v;
    }
    const int per_sm = occ_per_sm;
    int64_t grid = (int64_t)148 * per_sm;
    if (grid > units) grid = units;
    pdl(initial_conv_fwd_kernel, (unsigned)grid, threads, sm, c.st)(a, units, s.ln0_w, s.ln0_b, s.kernel, s.bias, s.ln1_w, s.ln1_b,
                                                                  s.shat, s.rstd_s, s.x0, fuse);
    c.post("initial_conv");
    if (!fuse) {
      const unsigned sg = (unsigned)cdiv64(s.R0, 8);
      switch (cdiv(s.C0, 32)) {
        case 1: pdl(stem_fwd_kernel<1>, sg, 256, 0, c.st)(s); break;
        case 2: pdl(stem_fwd_kernel<2>, sg, 256, 0, c.st)(s); break;
        case 3: pdl(stem_fwd_kernel<3>, sg, 256, 0, c.st)(s); break;
        default: pdl(stem_fwd_kernel<4>, sg, 256, 0, c.st)(s); break;
      }
      c.post("stem");
    }
  }
  x = c.w(pl->o_x0);
  for (int i = 0; i < 4; ++i) {
    if (i > 0) {  // downsample: LN + 2x2 stride-2 conv == [R/4, 4Cin] x [4Cin, Cout] on Z-ordered rows
      const int Ci = dm[i - 1], Co = dm[i];
      launch_ln_rows_fwd(x, c.w(pl->o_ds_xhat[i - 1]), c.w(pl->o_ds_rstd[i - 1]), pl->R[i - 1], Ci, 1e-6f, c.st);
      c.post("ds_ln");
      fold_slot(c, ds_fold_args(c, i - 1, Ci, Co), pl->ds_slot[i - 1], "fold_ds");
      GemmArgs g{};
      g.A = c.w(pl->o_ds_xhat[i - 1]); use_slot(c, g, pl->ds_slot[i - 1], false); g.bias = c.w(pl->ds_slot[i - 1].bf);
      g.out = c.w(pl->o_ds_out[i - 1]);
      g.M = pl->R[i]; g.N = Co; g.K = 4 * Ci; g.group_rows = 0x7fffffff;
      gemm<EPI_STORE>(c, g, "ds_conv");
      x = c.w(pl->o_ds_out[i - 1]);
    }
    for (int j = 0; j < cf.depths[i]; ++j) {
      block_forward(c, pl->blk[i][j], pl->bw[i][j], x, pl->R[i], dm[i], pl->P[i], false);
      x = c.w(pl->bw[i][j].y);
    }
  }
  }  // ST_ENC
  if (stages & ST_DEC) {
  // decoder entry: proj on visible rows, mask token elsewhere (fcmae.py:251-255)
  const int D = cf.dec_dim;
  {
    fold_slot(c, proj_fold_args(c), pl->proj_slot, "split_proj");
    GemmArgs g{};
    g.A = x; use_slot(c, g, pl->proj_slot, false); g.bias = c.p(pl->proj_b); g.out = c.w(pl->o_z);
    g.M = (int64_t)geo.B * geo.V; g.N = D; g.K = dm[3]; g.group_rows = 0x7fffffff;
    gemm<EPI_STORE>(c, g, "proj");
    pdl(scatter_token_kernel, ew_grid(pl->cells * (D / 4)), 256, 0, c.st)(c.w(pl->o_z), c.p(pl->tok), slot_of, c.w(pl->o_xd),
                                                                      pl->cells, geo.L, geo.V, D);
    c.post("scatter_token");
  }
  const float *d = c.w(pl->o_xd);
  for (int k = 0; k < cf.dec_depth; ++k) {
    block_forward(c, pl->dec[k], pl->dw[k], d, pl->cells, D, 1, true);
    d = c.w(pl->dw[k].y);
  }
  if (pl->npix > 0) {
    fold_slot(c, heads_fold_args(c), pl->pix_slot, "split_heads");
    GemmArgs g{};
    g.A = d; use_slot(c, g, pl->pix_slot, false); g.bias = c.p(pl->pixb); g.out = io->pred_pixel;
    g.M = pl->cells; g.N = pl->npix; g.K = D; g.group_rows = 0x7fffffff;
    static const bool heads_tf32 = getenv("MPMAE_HEADS_TF32") != nullptr;
    gemm<EPI_STORE>(c, g, "pixel_heads", heads_tf32);
  }
  if (pl->nimg > 0) {
    pdl(pool_ln_fwd_kernel, geo.B, 256, (size_t)8 * D * 4, c.st)(d, c.p(pl->lnt_w), c.p(pl->lnt_b), c.w(pl->o_pooled),
                                                          c.w(pl->o_pool_rstd), geo.L, D, 1e-6f);
    c.post("pool_ln");
    if (c.ok()) {
      pdl(small_gemm_nn_kernel, dim3(cdiv(pl->nimg, 32), cdiv(geo.B, 32)), 256, 0, c.st)(c.w(pl->o_pooled), c.p(pl->imgw), c.p(pl->imgb),
                                                                                       io->pred_image, geo.B, pl->nimg, D);
      c.post("image_heads");
    }
  }
  }  // ST_DEC
  if (stages & ST_LOSS) {
    LossArgs a = loss_args(c);
    if (pl->npix > 0) {
      int npm = 0;
      for (int m = 0; m < cf.n_mod; ++m) npm += pl->is_img[m] ? 0 : 1;
      // the continuous targets of a cell are staged in shared memory (read in memory order); measured against the
      // register-cached gather path: 0.124 vs 0.173 ms at cfg2 (patch 8), 0.26 vs 0.52 ms at cfg3 (patch 16)
      size_t cache = 0;
      bool big = false;
      for (int m = 0; m < cf.n_mod; ++m)
        if (cf.mod_kind[m] == MPMAE_PIXEL_CONTINUOUS) {
          cache += (size_t)cf.patch_size * cf.patch_size * cf.mod_chans[m] * sizeof(float);
          big = true;
        }
      static const int env_cache = getenv("MPMAE_LOSS_SMEM") ? atoi(getenv("MPMAE_LOSS_SMEM")) : -1;   // experiment: force off / on
      if (env_cache >= 0) big = env_cache != 0;
      a.smem_cache = (big && cache <= 200 * 1024) ? 1 : 0;
      if (a.smem_cache && cache > 48 * 1024) {
        static size_t configured = 0;
        if (cache > configured) {
          c.check(cudaFuncSetAttribute(pixel_loss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cache), "attr", false);
          configured = cache;
        }
      }
      pdl(pixel_loss_kernel, (unsigned)pl->cells, 32 * npm, a.smem_cache ? cache : 0, c.st)(a);
      c.post("pixel_loss");
    }
    if (pl->nimg > 0) {
      int nim = 0;
      for (int m = 0; m < cf.n_mod; ++m) nim += pl->is_img[m] ? 1 : 0;
      pdl(image_loss_kernel, geo.B, 32 * nim, 0, c.st)(a);
      c.post("image_loss");
    }
    pdl(loss_finalize_kernel, 1, 32, 0, c.st)(c.w(pl->o_acc), pl->logv >= 0 ? c.p(pl->logv) : nullptr, cf.n_mod,
                                             cf.loss_aggr, io->losses);
    c.post("loss_finalize");
  }
  int partial = 0;
  return finish(c, stages == ST_ALL ? &pl->launches_fwd : &partial);
}

int mpmae_forward(mpmae_plan *pl, const mpmae_io *io, void *cuda_stream) { return forward_impl(pl, io, cuda_stream, ST_ALL); }
int mpmae_forward_encoder(mpmae_plan *pl, const mpmae_io *io, void *cuda_stream) {
  return forward_impl(pl, io, cuda_stream, ST_MASK | ST_ENC);
}
int mpmae_forward_stages(mpmae_plan *pl, const mpmae_io *io, int32_t stages, void *cuda_stream) {
  if (stages <= 0 || (stages & ~ST_ALL)) return fail(MPMAE_ERR_INVALID, "stages %d: expected a non-empty MPMAE_STAGE_* mask", stages);
  return forward_impl(pl, io, cuda_stream, stages);
}

// -------------------------------------------------------------------------------------------------
// parts: bit 0 = seeds + heads + decoder + proj ; bit 1 = stages 3, 2 ; bit 2 = stages 1, 0 + patch embedding
// ext: step-wise backward (mpmae_backward_step).  which = 1: the decoder part starts from GIVEN prediction gradients
// (dpred_pixel / dpred_image, unit loss seeds) and hands the encoder-output gradient rows out through d_x3; which = 2: the
// encoder parts start from d_x3.
struct StepBwd { int which; const float *dpred_pixel, *dpred_image; float *d_x3; };
static int backward_impl(mpmae_plan *pl, const mpmae_io *io, void *cuda_stream, int parts, const StepBwd *ext = nullptr) {
  int rc = check_io(pl, io, true, ext ? (ext->which == 2 ? ST_ENC : 0) : ST_ALL);
  if (rc) return rc;
  claim_workspace(pl, io->workspace);
  Ctx c{pl, io, static_cast<cudaStream_t>(cuda_stream), static_cast<float *>(io->workspace), io->params, io->grads};
  c.mark("start");
  const mpmae_cfg &cf = pl->cfg;
  const int *dm = cf.dims;
  const Geo geo = pl->geo;
  const int D = cf.dec_dim;
  const int *slot_of = reinterpret_cast<const int *>(c.w(pl->o_slot));
  float *g0 = c.w(pl->o_g0), *g1 = c.w(pl->o_g1);

  float *cur = pl->bwd_flip ? g1 : g0, *nxt = pl->bwd_flip ? g0 : g1;
  const int64_t BV = (int64_t)geo.B * geo.V;
  if (parts & 1) {
  cur = g0; nxt = g1;
  c.zero(c.w(pl->o_bzero_begin), pl->o_bzero_end - pl->o_bzero_begin, "zero_bwd");
  if (!ext) {  // seeds: d total / d L_i / denominator_i per prediction column ; d total / d log_vars
    SeedArgs s{};
    s.acc = c.w(pl->o_acc); s.log_vars = pl->logv >= 0 ? c.p(pl->logv) : nullptr; s.losses = io->losses;
    s.grad_out = io->grad_out; s.d_log_vars = pl->logv >= 0 ? c.g(pl->logv) : nullptr;
    s.colscale_pix = c.w(pl->o_cs_pix); s.colscale_img = c.w(pl->o_cs_img);
    s.n_mod = cf.n_mod; s.uncertainty = cf.loss_aggr;
    for (int m = 0; m < cf.n_mod; ++m) { s.col_off[m] = pl->col_off[m]; s.col_len[m] = pl->col_len[m]; s.is_img[m] = pl->is_img[m]; }
    pdl(loss_seed_kernel, cf.n_mod, 256, 0, c.st)(s);
    c.post("loss_seed");
  } else {     // given prediction gradients: unit seeds
    if (pl->npix > 0) {
      pdl(fill_kernel, 8, 256, 0, c.st)(c.w(pl->o_cs_pix), 1.f, (int64_t)pl->npix);
      c.post("fill");
      c.check(cudaMemcpyAsync(c.w(pl->o_dpix), ext->dpred_pixel, (size_t)pl->cells * pl->npix * 4, cudaMemcpyDeviceToDevice, c.st), "memcpy", false);
    }
    if (pl->nimg > 0) {
      pdl(fill_kernel, 8, 256, 0, c.st)(c.w(pl->o_cs_img), 1.f, (int64_t)pl->nimg);
      c.post("fill");
      c.check(cudaMemcpyAsync(c.w(pl->o_dimg), ext->dpred_image, (size_t)geo.B * pl->nimg * 4, cudaMemcpyDeviceToDevice, c.st), "memcpy", false);
    }
  }
  const float *dec_out = c.w(pl->dw[cf.dec_depth - 1].y);
  float *dd = g0;  // gradient at the decoder output [cells, D]
  if (pl->npix > 0) {
    FoldArgs f{};
    f.W = c.p(pl->pixw); f.s_n = D; f.s_k = 1; f.scale_n = c.w(pl->o_cs_pix); f.SL = D;
    fold_slot(c, f, pl->pix_slot, "fold_pixT");   // loss seeds (known only now) scale the rows of the head matrix
    GemmArgs g{};
    g.A = c.w(pl->o_dpix); use_slot(c, g, pl->pix_slot, true); g.out = dd; g.M = pl->cells; g.N = D; g.K = pl->npix; g.group_rows = 0x7fffffff;
    static const bool heads_tf32 = getenv("MPMAE_HEADS_TF32") != nullptr;
    gemm<EPI_STORE>(c, g, "d_dec_pix", heads_tf32);
    WgradArgs w{};
    w.X = c.w(pl->o_dpix); w.Y = dec_out; w.rs = c.w(pl->o_cs_pix); w.dW = c.g(pl->pixw); w.db = c.g(pl->pixb);
    w.R = pl->cells; w.N = pl->npix; w.K = D;
    wgrad(c, w, "dW_pix");
  } else {
    c.zero(dd, pl->cells * D, "zero_dd");
  }
  if (pl->nimg > 0) {
    if (c.ok()) {
      pdl(small_gemm_nt_kernel, dim3(cdiv(D, 32), cdiv(geo.B, 32)), 256, 0, c.st)(
          c.w(pl->o_dimg), c.p(pl->imgw), c.w(pl->o_cs_img), c.w(pl->o_dpooled), geo.B, D, pl->nimg);
      c.post("d_pooled");
    }
    WgradArgs w{};
    w.X = c.w(pl->o_dimg); w.Y = c.w(pl->o_pooled); w.rs = c.w(pl->o_cs_img); w.dW = c.g(pl->imgw); w.db = c.g(pl->imgb);
    w.R = geo.B; w.N = pl->nimg; w.K = D;
    wgrad(c, w, "dW_img");
    if (c.ok()) {
      pdl(pool_ln_bwd_kernel, geo.B, 256, (size_t)8 * D * 4, c.st)(dec_out, c.w(pl->o_pool_rstd), c.p(pl->lnt_w), c.w(pl->o_dpooled),
                                                            dd, c.g(pl->lnt_w), c.g(pl->lnt_b), geo.L, D, 1e-6f);
      c.post("pool_ln_bwd");
    }
  }
  for (int k = cf.dec_depth - 1; k >= 0; --k) {
    const float *xin = k > 0 ? c.w(pl->dw[k - 1].y) : c.w(pl->o_xd);
    block_backward(c, pl->dec[k], pl->dw[k], xin, cur, nxt, pl->cells, D, 1, true);
    std::swap(cur, nxt);
  }
  // decoder entry backward: visible cells -> proj rows, masked cells -> mask token
  const float *x3 = c.w(pl->bw[3][cf.depths[3] - 1].y);
  {
    float *dz = c.w(pl->o_gdv);
    if (c.ok()) {
      pdl(gather_token_bwd_kernel, 148 * 4, ((D / 4 + 31) / 32) * 32, 0, c.st)(cur, slot_of, dz, c.g(pl->tok), pl->cells, geo.L, geo.V, D);
      c.post("gather_token_bwd");
    }
    WgradArgs w{};
    w.X = dz; w.Y = x3; w.dW = c.g(pl->proj_w); w.db = c.g(pl->proj_b); w.R = BV; w.N = D; w.K = dm[3];
    wgrad(c, w, "dW_proj");
    GemmArgs g{};
    g.A = dz; use_slot(c, g, pl->proj_slot, true); g.out = nxt; g.M = BV; g.N = dm[3]; g.K = D; g.group_rows = 0x7fffffff;
    gemm<EPI_STORE>(c, g, "d_x3");
    std::swap(cur, nxt);
    if (ext && ext->d_x3)
      c.check(cudaMemcpyAsync(ext->d_x3, cur, (size_t)BV * dm[3] * 4, cudaMemcpyDeviceToDevice, c.st), "memcpy", false);
  }
  }  // part 0
  if (ext && ext->which == 2) {   // the encoder parts start from the given gradient of the encoder output rows
    cur = g0; nxt = g1;
    c.zero(c.w(pl->o_bzero_begin), pl->o_bzero_end - pl->o_bzero_begin, "zero_bwd");
    c.check(cudaMemcpyAsync(cur, ext->d_x3, (size_t)BV * dm[3] * 4, cudaMemcpyDeviceToDevice, c.st), "memcpy", false);
  }
  for (int i = 3; i >= 0; --i) {
    if (!(parts & (i >= 2 ? 2 : 4))) continue;
    for (int j = cf.depths[i] - 1; j >= 0; --j) {
      const float *xin = j > 0 ? c.w(pl->bw[i][j - 1].y) : (i > 0 ? c.w(pl->o_ds_out[i - 1]) : c.w(pl->o_x0));
      // the depthwise dX kernel of this block also sums its output columns: the bias gradient of the pw2 of the block
      // below, or of the downsample conv when this is the first block of the stage
      float *dx_cs = j > 0 ? c.w(pl->bw[i][j - 1].s2.dbf) : (i > 0 ? c.w(pl->ds_slot[i - 1].dbf) : nullptr);
      block_backward(c, pl->blk[i][j], pl->bw[i][j], xin, cur, nxt, pl->R[i], dm[i], pl->P[i], false,
                     /*dy_colsum_done=*/j < cf.depths[i] - 1, dx_cs);
      std::swap(cur, nxt);
    }
    if (i > 0) {
      const int Ci = dm[i - 1], Co = dm[i];
      float *dwf = c.w(pl->ds_slot[i - 1].dwf), *dbf = c.w(pl->ds_slot[i - 1].dbf);
      WgradArgs w{};
      w.X = cur; w.Y = c.w(pl->o_ds_xhat[i - 1]); w.dW = dwf; w.db = nullptr; w.R = pl->R[i]; w.N = Co; w.K = 4 * Ci;
      wgrad(c, w, "dW_ds");
      UnfoldArgs u{};
      u.W = c.p(pl->ds[i - 1].k); u.s_n = 1; u.s_k = Co; u.scale_k = c.p(pl->ds[i - 1].ln_w); u.shift_k = c.p(pl->ds[i - 1].ln_b);
      u.dWf = dwf; u.dbf = dbf; u.dW = c.g(pl->ds[i - 1].k); u.dscale = c.g(pl->ds[i - 1].ln_w);
      u.dshift = c.g(pl->ds[i - 1].ln_b); u.dbias = c.g(pl->ds[i - 1].b); u.N = Co; u.K = 4 * Ci; u.SL = Ci;
      unfold(c, u, "unfold_ds", true);
      float *dxh = c.w(pl->o_gdv);
      GemmArgs g{};
      g.A = cur; use_slot(c, g, pl->ds_slot[i - 1], true); g.out = dxh; g.M = pl->R[i]; g.N = 4 * Ci; g.K = Co; g.group_rows = 0x7fffffff;
      gemm<EPI_STORE>(c, g, "d_ds_in");
      if (c.ok()) {
        launch_ln_rows_bwd(dxh, c.w(pl->o_ds_xhat[i - 1]), c.w(pl->o_ds_rstd[i - 1]), nullptr, nxt, pl->R[i - 1], Ci, c.st);
        c.post("ds_ln_bwd");
      }
      std::swap(cur, nxt);
    }
  }
  if (parts & 4) {  // stem + initial conv
    StemBwdArgs sb{};
    sb.f = stem_args(c);
    sb.dx0 = cur; sb.dc = nxt;
    sb.d_ln0_w = c.g(pl->ic_lnw); sb.d_ln0_b = c.g(pl->ic_lnb); sb.d_kernel = c.g(pl->st_k); sb.d_bias = c.g(pl->st_b);
    sb.d_ln1_w = c.g(pl->st_lnw); sb.d_ln1_b = c.g(pl->st_lnb);
    if (c.ok()) {
      const size_t ssm = (size_t)(5 + sb.f.s2) * sb.f.C0 * 4;
      c.acct(4.0 * ((2.0 + 2.0 * sb.f.s2) * sb.f.R0 * sb.f.C0), 0);
      static const bool no_vec = getenv("MPMAE_NO_STEMVEC") != nullptr;
      if (no_vec || !launch_stem_bwd_vec(sb, c.st)) {
        switch (cdiv(sb.f.C0, 32)) {
          case 1: pdl(stem_bwd_kernel<1>, 148 * 8, 256, ssm, c.st)(sb); break;
          case 2: pdl(stem_bwd_kernel<2>, 148 * 6, 256, ssm, c.st)(sb); break;
          case 3: pdl(stem_bwd_kernel<3>, 148 * 4, 256, ssm, c.st)(sb); break;
          default: pdl(stem_bwd_kernel<4>, 148 * 4, 256, ssm, c.st)(sb); break;
        }
      }
      c.post("stem_bwd");
    }
    InitConvWgradArgs iw{};
    iw.f = init_args(c);
    iw.dc = nxt; iw.dkernel = c.g(pl->ic_k); iw.dbias = c.g(pl->ic_b);
    const size_t sm = ((size_t)iw.f.Cin * 100 + 64 * (size_t)iw.f.C0 + (9 * (size_t)iw.f.Cin + 1) * iw.f.C0) * 4;
    static size_t configured = 0;
    if (sm > 48 * 1024 && sm > configured) {
      c.check(cudaFuncSetAttribute(initial_conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm), "attr");
      configured = sm;
    }
    if (c.ok()) {
      const int owners = iw.f.Cin * (iw.f.C0 / 4);
      const int parts = owners * 4 <= 512 ? 4 : (owners * 2 <= 512 ? 2 : 1);
      const int per_sm = (owners * parts <= 256) ? 4 : 2;
      pdl(initial_conv_wgrad_kernel, 148 * per_sm, owners * parts, sm, c.st)(iw);
      c.post("initial_conv_wgrad");
    }
  }
  flush_unfolds(c, parts == 7 ? 0 : (parts == 1 ? 0 : (parts == 2 ? 1 : 2)));
  pl->bwd_flip = (cur == g1) ? 1 : 0;
  int n = 0;
  rc = finish(c, &n);
  pl->launches_bwd = (parts & 1) ? n : pl->launches_bwd + n;
  return rc;
}

int mpmae_backward(mpmae_plan *pl, const mpmae_io *io, void *cuda_stream) { return backward_impl(pl, io, cuda_stream, 7); }

int mpmae_backward_part(mpmae_plan *pl, const mpmae_io *io, int32_t part, void *cuda_stream) {
  if (part < 0 || part > 2) return fail(MPMAE_ERR_INVALID, "backward part %d", part);
  return backward_impl(pl, io, cuda_stream, 1 << part);
}

int mpmae_backward_step(mpmae_plan *pl, const mpmae_io *io, int32_t which, float *dpred_pixel, float *dpred_image, float *d_x3,
                        void *cuda_stream) {
  if (!pl || !io) return fail(MPMAE_ERR_INVALID, "null plan/io");
  if (which == MPMAE_BWD_LOSS) {
    if (!io->workspace || !io->params || !io->grads || !io->losses) return fail(MPMAE_ERR_INVALID, "backward_step(loss): null pointer in mpmae_io");
    if ((pl->npix > 0 && !dpred_pixel) || (pl->nimg > 0 && !dpred_image)) return fail(MPMAE_ERR_INVALID, "backward_step(loss): null output");
    claim_workspace(pl, io->workspace);
    Ctx c{pl, io, static_cast<cudaStream_t>(cuda_stream), static_cast<float *>(io->workspace), io->params, io->grads};
    const mpmae_cfg &cf = pl->cfg;
    SeedArgs s{};
    s.acc = c.w(pl->o_acc); s.log_vars = pl->logv >= 0 ? c.p(pl->logv) : nullptr; s.losses = io->losses;
    s.grad_out = io->grad_out; s.d_log_vars = pl->logv >= 0 ? c.g(pl->logv) : nullptr;
    s.colscale_pix = c.w(pl->o_cs_pix); s.colscale_img = c.w(pl->o_cs_img);
    s.n_mod = cf.n_mod; s.uncertainty = cf.loss_aggr;
    for (int m = 0; m < cf.n_mod; ++m) { s.col_off[m] = pl->col_off[m]; s.col_len[m] = pl->col_len[m]; s.is_img[m] = pl->is_img[m]; }
    pdl(loss_seed_kernel, cf.n_mod, 256, 0, c.st)(s);
    c.post("loss_seed");
    if (pl->npix > 0 && c.ok()) {
      pdl(scale_cols_kernel, ew_grid(pl->cells * pl->npix / 4), 256, 0, c.st)(c.w(pl->o_dpix), c.w(pl->o_cs_pix), dpred_pixel, pl->cells, pl->npix);
      c.post("scale_cols");
    }
    if (pl->nimg > 0 && c.ok()) {
      pdl(scale_cols_kernel, ew_grid((int64_t)pl->geo.B * pl->nimg / 4 + 1), 256, 0, c.st)(c.w(pl->o_dimg), c.w(pl->o_cs_img), dpred_image, pl->geo.B, pl->nimg);
      c.post("scale_cols");
    }
    int n = 0;
    return finish(c, &n);
  }
  if (which == MPMAE_BWD_DECODER) {
    if ((pl->npix > 0 && !dpred_pixel) || (pl->nimg > 0 && !dpred_image) || !d_x3) return fail(MPMAE_ERR_INVALID, "backward_step(decoder): null pointer");
    StepBwd e{1, dpred_pixel, dpred_image, d_x3};
    return backward_impl(pl, io, cuda_stream, 1, &e);
  }
  if (which == MPMAE_BWD_ENCODER) {
    if (!d_x3) return fail(MPMAE_ERR_INVALID, "backward_step(encoder): d_x3 is null");
    StepBwd e{2, nullptr, nullptr, d_x3};
    return backward_impl(pl, io, cuda_stream, 6, &e);
  }
  return fail(MPMAE_ERR_INVALID, "backward_step: which = %d", which);
}

// [lo, hi) of the flat gradient buffer that is final once backward part `part` has run (reverse layer order)
int mpmae_backward_part_range(const mpmae_plan *pl, int32_t part, int64_t *lo, int64_t *hi) {
  if (!pl || !lo || !hi || part < 0 || part > 2) return fail(MPMAE_ERR_INVALID, "backward part %d", part);
  const int64_t mid_begin = pl->blk[2][0].dw_k, tail_begin = pl->proj_w;
  if (part == 0) { *lo = tail_begin; *hi = pl->n_params; }
  else if (part == 1) { *lo = mid_begin; *hi = tail_begin; }
  else { *lo = 0; *hi = mid_begin; }
  return MPMAE_OK;
}

int mpmae_profile_begin(mpmae_plan *pl) {
  if (!pl) return fail(MPMAE_ERR_INVALID, "null plan");
  pl->prof_on = true;
  pl->prof_n = 0;
  return MPMAE_OK;
}

int mpmae_profile_report(mpmae_plan *pl, char *buf, int32_t cap) {
  if (!pl || !buf || cap <= 0) return fail(MPMAE_ERR_INVALID, "null");
  pl->prof_on = false;
  struct Agg { int count = 0; double ms = 0, bytes = 0, flops = 0; };
  std::map<std::string, Agg> agg;
  std::vector<std::string> order;
  if (pl->prof_n > 0) {
    cudaError_t e = cudaEventSynchronize(pl->prof_ev[pl->prof_n - 1]);
    if (e != cudaSuccess) return fail(MPMAE_ERR_CUDA, "profile sync: %s", cudaGetErrorString(e));
  }
  for (int i = 1; i < pl->prof_n; ++i) {
    if (strcmp(pl->prof_name[i], "start") == 0) continue;   // interval spans host time between calls
    float ms = 0.f;
    cudaEventElapsedTime(&ms, pl->prof_ev[i - 1], pl->prof_ev[i]);
    auto it = agg.find(pl->prof_name[i]);
    if (it == agg.end()) { order.push_back(pl->prof_name[i]); it = agg.emplace(pl->prof_name[i], Agg()).first; }
    it->second.count += 1; it->second.ms += ms; it->second.bytes += pl->prof_bytes[i]; it->second.flops += pl->prof_flops[i];
  }
  int pos = snprintf(buf, cap, "name,launches,ms,alg_bytes,alg_flops\n");
  for (const auto &n : order) {
    const Agg &a = agg[n];
    if (pos >= cap - 1) break;
    pos += snprintf(buf + pos, cap - pos, "%s,%d,%.6f,%.0f,%.0f\n", n.c_str(), a.count, a.ms, a.bytes, a.flops);
  }
  pl->prof_n = 0;
  return MPMAE_OK;
}

int mpmae_encoder_features(mpmae_plan *pl, const mpmae_io *io, float *out_nchw, void *cuda_stream) {
  if (!pl || !io || !io->workspace || !out_nchw) return fail(MPMAE_ERR_INVALID, "null");
  float *ws = static_cast<float *>(io->workspace);
  const int C3 = pl->cfg.dims[3];
  const float *x3 = ws + pl->bw[3][pl->cfg.depths[3] - 1].y;
  const int64_t total = (int64_t)pl->geo.B * C3 * pl->geo.L;
  pdl(densify_kernel, ew_grid(total), 256, 0, static_cast<cudaStream_t>(cuda_stream))(
      x3, reinterpret_cast<const int *>(ws + pl->o_slot), out_nchw, pl->geo.B, pl->geo.L, pl->geo.V, C3);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(MPMAE_ERR_CUDA, "densify: %s", cudaGetErrorString(e));
  return MPMAE_OK;
}

int mpmae_gemm_rows(int32_t backend, const float *a, const float *b, const float *bias, float *out, int64_t M, int32_t N,
                    int32_t K, float *scratch, void *cuda_stream) {
  if (!a || !b || !out || M < 0 || N <= 0 || K <= 0 || K % 8 != 0) return fail(MPMAE_ERR_INVALID, "gemm args");
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  GemmArgs g{};
  g.A = a; g.Bw = b; g.bias = bias; g.out = out; g.M = M; g.N = N; g.K = K; g.group_rows = 0x7fffffff;
  cudaError_t e;
  if (backend != 0) {
    if (!tc_gemm_supported(EPI_STORE, g)) return fail(MPMAE_ERR_UNSUPPORTED, "shape not taken by the tcgen05 path");
    if (backend == 1 || backend == 3) {  // split the weight (TF32 hi + remainder, or a bf16 pair) in the caller's scratch
      if (!scratch) return fail(MPMAE_ERR_INVALID, "backends 1 and 3 need scratch of 2*N*K floats");
      FoldArgs f{};
      f.W = b; f.s_n = K; f.s_k = 1; f.Wf = scratch; f.Wf_lo = scratch + (int64_t)N * K; f.N = N; f.K = K; f.SL = K;
      f.b16 = backend == 3 ? 1 : 0;
      launch_fold(f, st);
      g.Bw = f.Wf; g.Bw_lo = f.Wf_lo; g.b16 = f.b16;
    }
    e = launch_gemm_rows_tc<EPI_STORE>(g, backend, st);
  } else {
    e = launch_gemm_rows_simt<EPI_STORE>(g, st);
  }
  if (e != cudaSuccess) return fail(MPMAE_ERR_CUDA, "gemm: %s", cudaGetErrorString(e));
  return MPMAE_OK;
}

int mpmae_gemm_epi(int32_t mode, int32_t backend, const mpmae_gemm_desc *d, void *cuda_stream) {
  if (!d || !d->a || !d->b || !d->out || d->M <= 0 || d->N <= 0 || d->K <= 0 || d->K % 8 != 0 || mode < 0 || mode > 3)
    return fail(MPMAE_ERR_INVALID, "gemm_epi args");
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  GemmArgs g{};
  g.A = d->a; g.Bw = d->b; g.bias = d->bias; g.resid = d->resid; g.out = d->out; g.out2 = d->out2; g.aux = d->aux;
  g.aux2 = d->aux2; g.kg = d->kg; g.colsum = d->colsum; g.colsum2 = d->colsum2; g.M = d->M; g.N = d->N; g.K = d->K;
  g.group_rows = d->group_rows > 0 ? d->group_rows : 0x7fffffff;
  g.a_gelu = d->a_gelu; g.a_scale = d->a_scale; g.acc_scale = d->acc_scale;
  g.grn_gsq = d->grn_gsq; g.grn_gamma = d->grn_gamma; g.grn_nx = d->grn_nx; g.grn_scale = d->grn_scale; g.grn_denom = d->grn_denom;
  g.grn_eps = d->grn_eps;
  if (g.grn_gsq && (!(backend == 1 || backend == 3) || mode != 0 || !g.a_gelu || !g.grn_gamma || !g.grn_nx || !g.grn_scale || !g.grn_denom))
    return fail(MPMAE_ERR_UNSUPPORTED, "the in-kernel GRN scale needs mode 0 with a_gelu on backend 1 or 3 and all grn_* pointers");
  if (g.a_gelu && backend == 2) return fail(MPMAE_ERR_UNSUPPORTED, "a_gelu needs operand-splitter warps (backends 1, 3) or backend 0");
  const bool tc_ok = backend != 0 && tc_gemm_supported(mode, g);
  if (backend != 0 && !tc_ok) return fail(MPMAE_ERR_UNSUPPORTED, "shape not taken by the tcgen05 path");
  if (backend == 1 || backend == 3) {
    if (!d->scratch) return fail(MPMAE_ERR_INVALID, "backends 1 and 3 need scratch of 2*N*K floats");
    FoldArgs f{};
    f.W = d->b; f.s_n = d->K; f.s_k = 1; f.Wf = d->scratch; f.Wf_lo = d->scratch + (int64_t)d->N * d->K; f.N = d->N;
    f.K = d->K; f.SL = d->K;
    f.b16 = backend == 3 ? 1 : 0;
    launch_fold(f, st);
    g.Bw = f.Wf; g.Bw_lo = f.Wf_lo; g.b16 = f.b16;
  }
  cudaError_t e = cudaSuccess;
  switch (mode) {
    case 0: e = tc_ok ? launch_gemm_rows_tc<EPI_STORE>(g, backend, st) : launch_gemm_rows_simt<EPI_STORE>(g, st); break;
    case 1: e = tc_ok ? launch_gemm_rows_tc<EPI_GELU_SQ>(g, backend, st) : launch_gemm_rows_simt<EPI_GELU_SQ>(g, st); break;
    case 2: e = tc_ok ? launch_gemm_rows_tc<EPI_DG>(g, backend, st) : launch_gemm_rows_simt<EPI_DG>(g, st); break;
    default: e = tc_ok ? launch_gemm_rows_tc<EPI_DH_GELU>(g, backend, st) : launch_gemm_rows_simt<EPI_DH_GELU>(g, st); break;
  }
  if (e != cudaSuccess) return fail(MPMAE_ERR_CUDA, "gemm_epi: %s", cudaGetErrorString(e));
  return MPMAE_OK;
}

#ifdef MPMAE_TC_KNOBS
// knob builds only (tools/tc_trace.py): the event trace of the last traced GEMM launch, 8 roles x 256 words
extern "C" int mpmae_debug_tc_trace(unsigned long long *dst_host) {
  if (!tc::tc_trace_buffer()) return 1;
  cudaDeviceSynchronize();
  return cudaMemcpy(dst_host, tc::tc_trace_buffer(), 8 * 256 * sizeof(unsigned long long), cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : 2;
}
#endif

int mpmae_gemm_wgrad(int32_t backend, const float *x, const float *y, float *dw, int64_t R, int32_t N, int32_t K,
                     void *cuda_stream) {
  return mpmae_gemm_wgrad_act(backend, x, y, dw, R, N, K, 0, cuda_stream);
}

int mpmae_gemm_wgrad_act(int32_t backend, const float *x, const float *y, float *dw, int64_t R, int32_t N, int32_t K,
                         int32_t y_gelu, void *cuda_stream) {
  if (!x || !y || !dw || R <= 0 || N <= 0 || K <= 0) return fail(MPMAE_ERR_INVALID, "wgrad args");
  if (y_gelu && backend == 2) return fail(MPMAE_ERR_UNSUPPORTED, "y_gelu needs the splitter warps of backends 1 / 3 (or backend 0)");
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  WgradArgs w{};
  w.X = x; w.Y = y; w.dW = dw; w.R = R; w.N = N; w.K = K; w.y_gelu = y_gelu ? 1 : 0;
  cudaError_t e;
  if (backend != 0) {
    if (!tc_wgrad_supported(w)) return fail(MPMAE_ERR_UNSUPPORTED, "shape not taken by the tcgen05 path");
    e = launch_gemm_wgrad_tc(w, backend == 1 || backend == 3, st);
  } else {
    e = launch_gemm_wgrad(w, st);
  }
  if (e != cudaSuccess) return fail(MPMAE_ERR_CUDA, "wgrad: %s", cudaGetErrorString(e));
  return MPMAE_OK;
}

int mpmae_dense_im2col(const float *x, float *out, int32_t B, int32_t C, int32_t H, int32_t W, int32_t k, int32_t s, int32_t kpad,
                       int32_t nchw, void *cuda_stream) {
  if (!x || !out || B <= 0 || C <= 0 || k <= 0 || s <= 0 || H < k || W < k || kpad < C * k * k || kpad % 8 != 0)
    return fail(MPMAE_ERR_INVALID, "dense_im2col args");
  const int Ho = (H - k) / s + 1, Wo = (W - k) / s + 1;
  pdl(dense_im2col_kernel, ew_grid((int64_t)B * Ho * Wo * kpad / 4), 256, 0, static_cast<cudaStream_t>(cuda_stream))(x, out, B, C, H, W, k, s,
                                                                                                                Ho, Wo, kpad, nchw);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? MPMAE_OK : fail(MPMAE_ERR_CUDA, "dense_im2col: %s", cudaGetErrorString(e));
}

int mpmae_ln_rows(const float *x, const float *w, const float *b, float *out, int64_t R, int32_t C, float eps, int32_t gelu,
                  void *cuda_stream) {
  if (!x || !out || R <= 0 || C <= 0 || (w && !b)) return fail(MPMAE_ERR_INVALID, "ln_rows args");
  pdl(ln_affine_rows_kernel, ew_grid(R * 8), 256, 0, static_cast<cudaStream_t>(cuda_stream))(x, w, b, out, R, C, eps, gelu);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? MPMAE_OK : fail(MPMAE_ERR_CUDA, "ln_rows: %s", cudaGetErrorString(e));
}

int mpmae_dense_dwconv(const float *x, const float *w, const float *bias, float *out, int32_t B, int32_t H, int32_t W, int32_t C,
                       int32_t k, int32_t s, int32_t p, int32_t ln, float eps, void *cuda_stream) {
  if (!x || !w || !out || B <= 0 || C <= 0 || C > 1024 || k <= 0 || s <= 0 || p < 0 || H + 2 * p < k || W + 2 * p < k)
    return fail(MPMAE_ERR_INVALID, "dense_dwconv args (C <= 1024)");
  const int Ho = (H + 2 * p - k) / s + 1, Wo = (W + 2 * p - k) / s + 1;
  pdl(dense_dwconv_kernel, ew_grid((int64_t)B * Ho * Wo * 8), 256, 0, static_cast<cudaStream_t>(cuda_stream))(x, w, bias, out, B, H, W, C, k,
                                                                                                         s, p, Ho, Wo, ln, eps);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? MPMAE_OK : fail(MPMAE_ERR_CUDA, "dense_dwconv: %s", cudaGetErrorString(e));
}

int mpmae_grn_apply(const float *h, const float *gsq, const float *gamma, const float *beta, float *out, int64_t R, int32_t D,
                    int32_t group_rows, float eps, float *scratch, void *cuda_stream) {
  if (!h || !gsq || !gamma || !beta || !out || !scratch || R <= 0 || D <= 0 || D % 4 != 0 || group_rows <= 0 || R % group_rows != 0)
    return fail(MPMAE_ERR_INVALID, "grn_apply args (D % 4 == 0, R a multiple of group_rows)");
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  const int groups = (int)(R / group_rows);
  float *nx = scratch, *scale = scratch + (int64_t)groups * D, *denom = scale + (int64_t)groups * D;
  pdl(grn_scale_kernel, groups, 256, 0, st)(gsq, gamma, nx, scale, denom, D, eps);
  pdl(grn_apply_kernel, ew_grid(R * (D / 4)), 256, 0, st)(h, scale, beta, out, R, D, group_rows);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? MPMAE_OK : fail(MPMAE_ERR_CUDA, "grn_apply: %s", cudaGetErrorString(e));
}

int mpmae_raw_transform(const mpmae_raw_desc *d, void *cuda_stream) {
  if (!d || !d->src || !d->out || d->B <= 0 || d->inner <= 0 || d->n_bands <= 0 || d->src_bands <= 0 || d->src_type < 0 || d->src_type > 2)
    return fail(MPMAE_ERR_INVALID, "raw_transform args");
  if (d->n_bands > MPMAE_RAW_MAX_BANDS && d->inner != 1)
    return fail(MPMAE_ERR_UNSUPPORTED, "raw_transform: at most %d selected bands", MPMAE_RAW_MAX_BANDS);
  RawArgs a{};
  a.src = d->src; a.out = d->out; a.l2a = d->l2a; a.lut = d->lut; a.inner = d->inner; a.B = d->B; a.src_bands = d->src_bands;
  a.n_bands = d->n_bands; a.src_type = d->src_type; a.out_int64 = d->out_int64; a.has_nodata = d->has_nodata;
  a.normalize = d->normalize; a.nodata = d->nodata;
  if (d->n_bands > MPMAE_RAW_MAX_BANDS) {   // a long vector taken whole (one-hot rows): one band of n_bands elements
    if (d->normalize || d->src_bands != d->n_bands) return fail(MPMAE_ERR_UNSUPPORTED, "raw_transform: long rows are taken whole");
    a.inner = d->n_bands; a.n_bands = 1; a.src_bands = 1; a.band[0] = 0;
  } else {
    for (int b = 0; b < d->n_bands; ++b) {
      if (d->band[b] < 0 || d->band[b] >= d->src_bands) return fail(MPMAE_ERR_INVALID, "raw_transform: band index");
      a.band[b] = d->band[b];
      for (int s = 0; s < 2; ++s) { a.mean[s][b] = d->mean[s][b]; a.stdv[s][b] = d->std[s][b]; }
    }
  }
  cudaError_t e = launch_raw_transform(a, static_cast<cudaStream_t>(cuda_stream));
  if (e != cudaSuccess) return fail(MPMAE_ERR_CUDA, "raw_transform: %s", cudaGetErrorString(e));
  return MPMAE_OK;
}

int mpmae_adamw_step(float *params, const float *grads, float *exp_avg, float *exp_avg_sq, const uint8_t *decay_mask,
                     int64_t n, float lr, float beta1, float beta2, float eps, float weight_decay, int64_t step,
                     float grad_scale_inv, void *cuda_stream) {
  if (!params || !grads || !exp_avg || !exp_avg_sq || n <= 0 || step < 1) return fail(MPMAE_ERR_INVALID, "adamw args");
  cudaError_t e = launch_adamw(params, grads, exp_avg, exp_avg_sq, decay_mask, n, lr, beta1, beta2, eps, weight_decay, step,
                               grad_scale_inv, static_cast<cudaStream_t>(cuda_stream));
  if (e != cudaSuccess) return fail(MPMAE_ERR_CUDA, "adamw: %s", cudaGetErrorString(e));
  return MPMAE_OK;
}

int mpmae_adamw_step_dev(float *params, const float *grads, float *exp_avg, float *exp_avg_sq, const uint8_t *decay_mask,
                         int64_t n, float lr, float beta1, float beta2, float eps, float weight_decay,
                         const float *dev_state, void *cuda_stream) {
  if (!params || !grads || !exp_avg || !exp_avg_sq || !dev_state || n <= 0) return fail(MPMAE_ERR_INVALID, "adamw args");
  cudaError_t e = launch_adamw_dev(params, grads, exp_avg, exp_avg_sq, decay_mask, n, lr, beta1, beta2, eps, weight_decay,
                                   dev_state, static_cast<cudaStream_t>(cuda_stream));
  if (e == cudaErrorInvalidValue) return fail(MPMAE_ERR_UNSUPPORTED, "adamw_step_dev needs n % 4 == 0 and 16-byte aligned buffers");
  if (e != cudaSuccess) return fail(MPMAE_ERR_CUDA, "adamw: %s", cudaGetErrorString(e));
  return MPMAE_OK;
}

}  // extern "C"
