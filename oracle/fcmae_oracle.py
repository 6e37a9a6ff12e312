"""Stand-alone CPU restatement of the reference MP-MAE (FCMAE) pretraining forward.

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this file; the product
path (``mmearth_train_b200``) never does and fails loudly without its CUDA library.

Why a restatement: the reference's sparse path needs MinkowskiEngine (un-buildable here, GPU-only
depthwise op) and ``/root/reference`` does not exist on the GPU box.  This file restates the
algorithm in pure torch (CPU, fp32 or fp64, autograd-differentiable) with the reference's
state-dict keys and shapes, so weights interchange with the reference and with the CUDA path.

Parity pinning (tests/test_oracle_*.py, oracle/make_golden.py):
  * against the UNMODIFIED reference ``FCMAE(sparse=True)`` running on ``oracle/me_shim.py``
    (build container only) -> golden vectors in ``tests/golden/``;
  * ``me_shim`` itself against the reference's depthwise known-answer vectors
    (``MinkowskiEngine/MinkowskiEngine/MinkowskiDepthwiseConvolution.py:200-263``) and the
    sparse<->dense weight-layout identity of ``helpers.py:676-690``.

Formulation: every sparse op equals the dense op on the zero-filled masked image
with zero padding, evaluated only at active sites, with the active mask re-applied after every op;
the sparse GRN statistic runs over all active rows of the whole per-GPU batch.

Reference lines followed:
  encoder graph          models/convnextv2_sparse.py:26-56,99-152,191-220
  sparse LN / GRN        models/sparse_norm_layers.py:16-33,61-77
  active set             MinkowskiEngine/MinkowskiEngine/MinkowskiOps.py:308-317
  kernel index <-> offset  MinkowskiEngine/src/kernel_region.hpp:199-221
  strided coordinates    MinkowskiEngine/src/coordinate_map.hpp:59-66
  densify                MinkowskiEngine/MinkowskiEngine/MinkowskiSparseTensor.py:512-554
  mask                   models/fcmae.py:214-231
  decoder                models/fcmae.py:249-265, models/convnextv2.py:42-55, models/norm_layers.py:23-44
  losses                 models/fcmae.py:267-412, custom_loss.py:19-30
  forward                models/fcmae.py:414-456
"""
from __future__ import annotations

from argparse import Namespace
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

PIXEL_CONTINUOUS = ("sentinel2", "sentinel1", "aster", "canopy_height_eth")
PIXEL_CATEGORICAL = ("dynamic_world", "esa_worldcover")
IMAGE_CATEGORICAL = ("biome", "eco_region")
IMAGE_CONTINUOUS = ("lat", "lon", "month", "era5")
PIXEL_MODALITIES = PIXEL_CONTINUOUS + PIXEL_CATEGORICAL
IMAGE_MODALITIES = IMAGE_CATEGORICAL + IMAGE_CONTINUOUS
N_CLASSES = {"dynamic_world": 9, "esa_worldcover": 11, "biome": 14, "eco_region": 846}

MODEL_ZOO = {
    "convnextv2_atto": ([2, 2, 6, 2], [40, 80, 160, 320]),
    "convnextv2_femto": ([2, 2, 6, 2], [48, 96, 192, 384]),
    "convnextv2_pico": ([2, 2, 6, 2], [64, 128, 256, 512]),
    "convnextv2_nano": ([2, 2, 8, 2], [80, 160, 320, 640]),
    "convnextv2_tiny": ([3, 3, 9, 3], [96, 192, 384, 768]),
    "convnextv2_base": ([3, 3, 27, 3], [128, 256, 512, 1024]),
    "convnextv2_large": ([3, 3, 27, 3], [192, 384, 768, 1536]),
    "convnextv2_huge": ([3, 3, 27, 3], [352, 704, 1408, 2816]),
}


def out_channels(args: Namespace) -> Dict[str, int]:
    """Per-modality channel count, as ``models/fcmae.py:70-91``."""
    oc = {}
    for m, bands in args.modalities.items():
        if m in N_CLASSES:
            oc[m] = N_CLASSES[m]
        else:
            oc[m] = len(args.modalities_full[m]) if bands == "all" else len(bands)
    return oc


# --------------------------------------------------------------------------- parameter holders
class MEConvParams(nn.Module):
    """kernel [K, Cin, Cout] (or [K, C] depthwise), bias [1, Cout] -- ME layouts."""

    def __init__(self, kshape, cout):
        super().__init__()
        self.kernel = nn.Parameter(torch.zeros(*kshape))
        self.bias = nn.Parameter(torch.zeros(1, cout))


class LNWrap(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.ln = nn.LayerNorm(c, eps=1e-6)


class LinWrap(nn.Module):
    def __init__(self, i, o):
        super().__init__()
        self.linear = nn.Linear(i, o)


class SparseGRNParams(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.gamma = nn.Parameter(torch.zeros(1, d))
        self.beta = nn.Parameter(torch.zeros(1, d))


class SparseBlockParams(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.dwconv = MEConvParams((49, c), c)
        self.norm = LNWrap(c)
        self.pwconv1 = LinWrap(c, 4 * c)
        self.pwconv2 = LinWrap(4 * c, c)
        self.grn = SparseGRNParams(4 * c)


class _AffineLN(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(c))
        self.bias = nn.Parameter(torch.zeros(c))


class _DenseGRN(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.gamma = nn.Parameter(torch.zeros(1, 1, 1, d))
        self.beta = nn.Parameter(torch.zeros(1, 1, 1, d))


class DenseBlockParams(nn.Module):
    """Decoder block parameters (``models/convnextv2.py:26-40`` key names)."""

    def __init__(self, c):
        super().__init__()
        self.dwconv = nn.Conv2d(c, c, 7, padding=3, groups=c)
        self.norm = _AffineLN(c)
        self.pwconv1 = nn.Linear(c, 4 * c)
        self.grn = _DenseGRN(4 * c)
        self.pwconv2 = nn.Linear(4 * c, c)


def me_to_torch_conv_weight(kernel: torch.Tensor, K: int) -> torch.Tensor:
    """ME kernel -> torch conv weight; kernel index k = kh + K*kw (axis 0 fastest)."""
    if kernel.dim() == 3:  # [K*K, Cin, Cout] -> [Cout, Cin, kh, kw]
        kv, ci, co = kernel.shape
        return kernel.reshape(K, K, ci, co).permute(3, 2, 1, 0)
    kv, c = kernel.shape  # depthwise [K*K, C] -> [C, 1, kh, kw]
    return kernel.reshape(K, K, c).permute(2, 1, 0).unsqueeze(1)


def _ln_cl(x_nchw, ln: nn.LayerNorm):
    y = F.layer_norm(x_nchw.permute(0, 2, 3, 1), ln.normalized_shape, ln.weight, ln.bias, ln.eps)
    return y.permute(0, 3, 1, 2)


class OracleEncoder(nn.Module):
    def __init__(self, in_chans, depths, dims, patch_size, img_size):
        super().__init__()
        self.depths, self.dims = list(depths), list(dims)
        self.patch_size, self.img_size = patch_size, img_size
        self.stem_stride = patch_size // 8
        self.initial_conv = nn.Sequential(MEConvParams((9, in_chans, dims[0]), dims[0]), LNWrap(dims[0]), nn.GELU())
        self.stem = nn.Sequential(MEConvParams((self.stem_stride ** 2, dims[0]), dims[0]), LNWrap(dims[0]))
        self.downsample_layers = nn.ModuleList(
            nn.Sequential(LNWrap(dims[i]), MEConvParams((4, dims[i], dims[i + 1]), dims[i + 1])) for i in range(3))
        self.stages = nn.ModuleList(
            nn.Sequential(*[SparseBlockParams(dims[i]) for _ in range(depths[i])]) for i in range(4))

    def _block(self, x, act, p: SparseBlockParams, taps=None):
        c = x.shape[1]
        u = F.conv2d(x, me_to_torch_conv_weight(p.dwconv.kernel, 7), p.dwconv.bias.reshape(-1), padding=3, groups=c)
        v = F.layer_norm(u.permute(0, 2, 3, 1), (c,), p.norm.ln.weight, p.norm.ln.bias, 1e-6)
        a = p.pwconv1.linear(v)
        h = F.gelu(a) * act.permute(0, 2, 3, 1)            # rows that do not exist contribute nothing
        gx = torch.sqrt((h * h).sum(dim=(0, 1, 2)))          # over every active row of the batch
        nx = gx / (gx.mean() + 1e-6)
        g = p.grn.gamma.reshape(-1) * (h * nx) + p.grn.beta.reshape(-1) + h
        y = p.pwconv2.linear(g).permute(0, 3, 1, 2)
        out = (x + y) * act
        if taps is not None:
            taps.append(dict(u=u * act, v=v.permute(0, 3, 1, 2) * act, a=a.permute(0, 3, 1, 2) * act,
                             h=h.permute(0, 3, 1, 2), y=out))
        return out

    def forward(self, imgs: torch.Tensor, mask: torch.Tensor, taps: Optional[dict] = None) -> torch.Tensor:
        B = imgs.shape[0]
        g = int(round(mask.shape[1] ** 0.5))
        scale = self.img_size // g
        m = mask.reshape(B, g, g).repeat_interleave(scale, 1).repeat_interleave(scale, 2).unsqueeze(1).to(imgs.dtype)
        x = imgs * (1.0 - m)
        act = (x.abs().sum(1, keepdim=True) != 0).to(imgs.dtype)   # to_sparse active set
        ic, ln0, _ = self.initial_conv
        x = F.conv2d(x, me_to_torch_conv_weight(ic.kernel, 3), ic.bias.reshape(-1), padding=1)
        x = F.gelu(_ln_cl(x, ln0.ln)) * act
        if taps is not None:
            taps["initial"] = x
        st, ln1 = self.stem
        s = self.stem_stride
        c0 = self.dims[0]
        x = F.conv2d(x, me_to_torch_conv_weight(st.kernel, s), st.bias.reshape(-1), stride=s, groups=c0)
        act = F.max_pool2d(act, s) if s > 1 else act
        x = _ln_cl(x, ln1.ln) * act
        if taps is not None:
            taps["stem"] = x
            taps["blocks"] = []
        for i in range(4):
            if i > 0:
                lnd, cv = self.downsample_layers[i - 1]
                x = _ln_cl(x, lnd.ln) * act
                x = F.conv2d(x, me_to_torch_conv_weight(cv.kernel, 2), cv.bias.reshape(-1), stride=2)
                act = F.max_pool2d(act, 2)
                x = x * act
                if taps is not None:
                    taps.setdefault("down", []).append(x)
            for blk in self.stages[i]:
                x = self._block(x, act, blk, taps["blocks"] if taps is not None else None)
        return x  # [B, C3, g, g], zeros at masked cells


class OracleFCMAE(nn.Module):
    """Same constructor meaning, state-dict keys and forward tuple as ``models/fcmae.py:FCMAE``."""

    def __init__(self, img_size=112, depths=None, dims=None, decoder_depth=1, decoder_embed_dim=512,
                 patch_size=16, mask_ratio=0.6, norm_pix_loss=False, args: Namespace = None, loss_fn=None,
                 sparse=True):
        super().__init__()
        self.args = args
        self.img_size, self.patch_size, self.mask_ratio = img_size, patch_size, mask_ratio
        self.depths = depths or [3, 3, 9, 3]
        self.dims = dims or [96, 192, 384, 768]
        self.norm_pix_loss = norm_pix_loss
        self.decoder_embed_dim, self.decoder_depth = decoder_embed_dim, decoder_depth
        self.loss_fn = loss_fn
        self.out_chans = out_channels(args)
        s2 = args.modalities["sentinel2"]
        self.in_chans = len(args.modalities_full["sentinel2"]) if s2 == "all" else len(s2)
        self.encoder = OracleEncoder(self.in_chans, self.depths, self.dims, patch_size, img_size)
        self.proj = nn.Conv2d(self.dims[-1], decoder_embed_dim, 1)
        self.mask_token = nn.Parameter(torch.zeros(1, decoder_embed_dim, 1, 1))
        shared = [DenseBlockParams(decoder_embed_dim) for _ in range(decoder_depth)]   # ONE block set, aliased
        self.decoder_dict = nn.ModuleDict()
        self.pred_dict = nn.ModuleDict()
        for m in args.out_modalities:
            self.decoder_dict[m] = nn.Sequential(*shared)
            if m in PIXEL_MODALITIES:
                self.pred_dict[m] = nn.Conv2d(decoder_embed_dim, patch_size ** 2 * self.out_chans[m], 1)
            else:
                self.layer_norm_tmp = _AffineLN(decoder_embed_dim)
                self.pred_dict[m] = nn.Linear(decoder_embed_dim, self.out_chans[m])

    # ---------------------------------------------------------------- pieces
    @staticmethod
    def mask_from_noise(noise: torch.Tensor, mask_ratio: float) -> torch.Tensor:
        """``fcmae.py:214-231``: a patch is kept iff its rank in ascending noise order < len_keep."""
        L = noise.shape[1]
        keep = int(L * (1 - mask_ratio))
        rank = torch.argsort(torch.argsort(noise, dim=1, stable=True), dim=1, stable=True)
        return (rank >= keep).to(torch.float32)

    def patchify(self, imgs, channels):
        p = self.patch_size
        B = imgs.shape[0]
        g = imgs.shape[2] // p
        x = imgs.reshape(B, channels, g, p, g, p).permute(0, 2, 4, 3, 5, 1)   # n h w p q c
        return x.reshape(B, g * g, p * p * channels)

    def decoder_block(self, x, p: DenseBlockParams):
        u = p.dwconv(x).permute(0, 2, 3, 1)
        v = F.layer_norm(u, (u.shape[-1],), p.norm.weight, p.norm.bias, 1e-6)
        h = F.gelu(p.pwconv1(v))
        gx = torch.sqrt((h * h).sum(dim=(1, 2), keepdim=True))               # per sample
        nx = gx / (gx.mean(dim=-1, keepdim=True) + 1e-4)
        g = p.grn.gamma * (h * nx) + p.grn.beta + h
        return x + p.pwconv2(g).permute(0, 3, 1, 2)

    def forward_decoder(self, x, mask):
        z = self.proj(x)
        B, c, h, w = z.shape
        m = mask.reshape(B, 1, h, w).to(z.dtype)
        z = z * (1.0 - m) + self.mask_token * m
        first = next(iter(self.args.out_modalities))
        d = z
        for blk in self.decoder_dict[first]:
            d = self.decoder_block(d, blk)      # the 12 decoders are one aliased block: evaluate once
        pred = {}
        pooled = None
        for mname in self.args.out_modalities:
            if mname in PIXEL_MODALITIES:
                pred[mname] = self.pred_dict[mname](d)
            else:
                if pooled is None:
                    u = d.mean(1, keepdim=True)
                    s = (d - u).pow(2).mean(1, keepdim=True)
                    n = (d - u) / torch.sqrt(s + 1e-6)
                    n = self.layer_norm_tmp.weight[:, None, None] * n + self.layer_norm_tmp.bias[:, None, None]
                    pooled = n.mean(dim=(-2, -1))
                pred[mname] = self.pred_dict[mname](pooled)
        return pred, d

    def forward_loss(self, imgs_dict, preds, mask):
        p2 = self.patch_size ** 2
        losses = {}
        for mname in self.args.out_modalities:
            pr, tg = preds[mname], imgs_dict[mname]
            if mname in IMAGE_CATEGORICAL:
                losses[mname] = F.cross_entropy(pr, tg.argmax(-1))
            elif mname in IMAGE_CONTINUOUS:
                ok = ~torch.isnan(tg)
                losses[mname] = ((pr[ok] - tg[ok]) ** 2).mean()
            elif mname in PIXEL_CATEGORICAL:
                B, c = pr.shape[:2]
                K = self.out_chans[mname]
                logits = pr.reshape(B, c, -1).transpose(1, 2).reshape(B, -1, p2, K)      # [B, L, p2, K]
                t = self.patchify(tg, 1)                                                  # [B, L, p2]
                sel = (mask.unsqueeze(-1) == 1) & (t != -1)
                losses[mname] = F.cross_entropy(logits[sel], t[sel].long())
            else:
                B, c = pr.shape[:2]
                q = pr.reshape(B, c, -1).transpose(1, 2)                                  # [B, L, p2*c]
                t = self.patchify(tg, self.out_chans[mname])
                if self.norm_pix_loss and mname == "sentinel2":
                    t = (t - t.mean(-1, keepdim=True)) / (t.var(-1, keepdim=True) + 1.0e-6) ** 0.5
                e = (q - t) ** 2
                bad = torch.isnan(e)
                cnt = (~bad).sum(-1)
                per_patch = torch.where(bad, torch.zeros_like(e), e).sum(-1) / cnt
                tmp = per_patch * mask
                tmp = torch.where(torch.isnan(tmp), torch.zeros_like(tmp), tmp)
                losses[mname] = tmp.sum() / torch.count_nonzero(tmp)
        lst = list(losses.values())
        if self.args.loss_aggr == "uncertainty":
            lt = torch.stack(lst)
            s = self.loss_fn.log_vars
            weighted = (torch.exp(-s) * lt + s) * (lt != 0.0)
            return weighted.sum(), losses, s.tolist(), weighted
        return sum(lst), losses, None, None

    def forward(self, imgs_dict, labels=None, mask_ratio=0.6, noise: Optional[torch.Tensor] = None,
                taps: Optional[dict] = None):
        imgs_dict = dict(imgs_dict)
        imgs = imgs_dict["sentinel2"]              # bound before nan_to_num (fcmae.py:439-449)
        for m in PIXEL_CONTINUOUS:
            if m in imgs_dict:
                imgs_dict[m] = torch.nan_to_num(imgs_dict[m], nan=0.0, posinf=0.0, neginf=0.0)
        B = imgs.shape[0]
        L = (imgs.shape[2] // self.patch_size) ** 2
        if noise is None:
            noise = torch.randn(B, L, device=imgs.device)
        mask = self.mask_from_noise(noise, mask_ratio)
        x = self.encoder(imgs, mask.to(imgs.dtype), taps)
        pred, dec = self.forward_decoder(x, mask.to(imgs.dtype))
        if taps is not None:
            taps["encoder_out"], taps["decoder_out"] = x, dec
        loss, loss_dict, log_vars, weighted = self.forward_loss(imgs_dict, pred, mask.to(imgs.dtype))
        return loss, pred, mask, loss_dict, log_vars, weighted


class UncertaintyWeights(nn.Module):
    """Holder with the reference key ``loss_fn.log_vars`` (``custom_loss.py:10-17``)."""

    def __init__(self, tasks: int):
        super().__init__()
        self.tasks = tasks
        self.log_vars = nn.Parameter(torch.zeros(tasks))


# --------------------------------------------------------------------------- config + synthetic data
S2_BANDS = ["B1", "B2", "B3", "B4", "B5", "B6", "B7", "B8A", "B8", "B9", "B11", "B12"]
FULL_BANDS = {"sentinel2": 13, "sentinel1": 8, "aster": 2, "era5": 12, "dynamic_world": 1,
              "canopy_height_eth": 2, "lat": 2, "lon": 2, "biome": 1, "eco_region": 1, "month": 2,
              "esa_worldcover": 1}
ALL_OUT = ["sentinel2", "sentinel1", "aster", "era5", "dynamic_world", "canopy_height_eth", "lat", "lon",
           "biome", "eco_region", "month", "esa_worldcover"]          # MODALITIES.py:75-101 order


def make_args(out_modalities: Optional[List[str]] = None, loss_aggr="uncertainty") -> Namespace:
    outs = ALL_OUT if out_modalities is None else list(out_modalities)
    out = {m: (S2_BANDS if m == "sentinel2" else "all") for m in outs}
    mods = {"sentinel2": S2_BANDS}
    mods.update(out)
    full = {m: [f"{m}_{i}" for i in range(n)] for m, n in FULL_BANDS.items()}
    return Namespace(inp_modalities={"sentinel2": S2_BANDS}, out_modalities=out, modalities=mods,
                     modalities_full=full, use_orig_stem=False, loss_aggr=loss_aggr)


def build_oracle(model="convnextv2_atto", img_size=56, patch_size=8, out_modalities=None,
                 loss_aggr="uncertainty", norm_pix_loss=True, mask_ratio=0.6, decoder_depth=1,
                 decoder_embed_dim=512, args: Optional[Namespace] = None) -> OracleFCMAE:
    args = args or make_args(out_modalities, loss_aggr)
    depths, dims = MODEL_ZOO[model]
    lf = UncertaintyWeights(len(args.out_modalities)) if loss_aggr == "uncertainty" else None
    return OracleFCMAE(img_size=img_size, depths=depths, dims=dims, decoder_depth=decoder_depth,
                       decoder_embed_dim=decoder_embed_dim, patch_size=patch_size, mask_ratio=mask_ratio,
                       norm_pix_loss=norm_pix_loss, args=args, loss_fn=lf)


def init_like_reference(model: nn.Module, seed: int = 0, std_scale: float = 1.0) -> None:
    """Deterministic non-degenerate weights for parity tests.

    Not the reference initialiser (that needs timm): every tensor gets seeded N(0, s) values with a
    scale that keeps activations O(1), *including* GRN gamma/beta, biases and log_vars (zero in the
    reference init) so that every gradient path is exercised.
    """
    g = torch.Generator().manual_seed(seed)
    seen = set()
    with torch.no_grad():
        for name, p in model.named_parameters():
            if id(p) in seen:
                continue
            seen.add(id(p))
            leaf = name.split(".")[-1]
            if leaf == "log_vars":
                p.copy_(0.3 * torch.randn(p.shape, generator=g))
            elif leaf == "weight" and p.dim() == 1:            # LN scales
                p.copy_(1.0 + 0.2 * torch.randn(p.shape, generator=g))
            elif leaf in ("gamma",):
                p.copy_(0.3 * torch.randn(p.shape, generator=g))
            elif leaf in ("beta", "bias"):
                p.copy_(0.1 * torch.randn(p.shape, generator=g))
            elif name.endswith("dwconv.kernel") or name.endswith("dwconv.weight") or "stem.0.kernel" in name:
                p.copy_(0.15 * std_scale * torch.randn(p.shape, generator=g))
            elif leaf == "mask_token":
                p.copy_(0.5 * torch.randn(p.shape, generator=g))
            else:
                fan_in = p[0].numel() if p.dim() != 3 else p.shape[0] * p.shape[1]
                if leaf == "kernel" and p.dim() == 3:
                    fan_in = p.shape[0] * p.shape[1]
                p.copy_(std_scale * torch.randn(p.shape, generator=g) / fan_in ** 0.5)


def synthetic_batch(B: int, img_size: int, out_modalities: Optional[List[str]] = None, seed: int = 1234,
                    nan_frac: float = 0.0) -> Dict[str, torch.Tensor]:
    """SURVEY.md §8(d) synthetic tensors (dtypes/shapes of ``mmearth_dataset.py:58-153``)."""
    g = torch.Generator().manual_seed(seed)
    outs = ALL_OUT if out_modalities is None else list(out_modalities)
    S = img_size
    d = {"sentinel2": torch.randn(B, 12, S, S, generator=g)}
    for m, c in (("sentinel1", 8), ("aster", 2), ("canopy_height_eth", 2)):
        if m in outs:
            t = torch.randn(B, c, S, S, generator=g)
            if nan_frac > 0:
                t[torch.rand(t.shape, generator=g) < nan_frac] = float("nan")
            d[m] = t
    if "dynamic_world" in outs:
        d["dynamic_world"] = torch.randint(-1, 9, (B, 1, S, S), generator=g)
    if "esa_worldcover" in outs:
        d["esa_worldcover"] = torch.randint(-1, 11, (B, 1, S, S), generator=g)
    if "biome" in outs:
        d["biome"] = F.one_hot(torch.randint(0, 14, (B,), generator=g), 14)
    if "eco_region" in outs:
        d["eco_region"] = F.one_hot(torch.randint(0, 846, (B,), generator=g), 846)
    for m, c in (("lat", 2), ("lon", 2), ("month", 2), ("era5", 12)):
        if m in outs:
            t = torch.randn(B, c, generator=g)
            if m == "era5" and nan_frac > 0:
                t[torch.rand(t.shape, generator=g) < nan_frac / 2] = float("nan")
            d[m] = t
    return d
