"""CPU restatement of the reference's DENSE ConvNeXt-V2 forward (TEST INFRASTRUCTURE ONLY).

The network finetuning / linear probing runs on a pretrained encoder: ``models/convnextv2.py:59-207`` (block ``:18-56``,
custom LayerNorm / per-sample GRN ``models/norm_layers.py:7-44``), with the dense state-dict layout that
``helpers.remap_checkpoint_keys`` produces from a pretraining checkpoint.  Pure functional torch over a state dict, so that the
GPU box (which has no ``/root/reference``) can check the CUDA path; pinned against the unmodified reference module by
``tests/golden/dense_*.npz`` (``oracle/make_dense_golden.py``).
"""
from __future__ import annotations

from typing import Dict, List

import torch
import torch.nn.functional as F

ZOO = {"convnextv2_atto": ([2, 2, 6, 2], [40, 80, 160, 320]), "convnextv2_femto": ([2, 2, 6, 2], [48, 96, 192, 384]),
       "convnextv2_pico": ([2, 2, 6, 2], [64, 128, 256, 512]), "convnextv2_nano": ([2, 2, 8, 2], [80, 160, 320, 640]),
       "convnextv2_tiny": ([3, 3, 9, 3], [96, 192, 384, 768]), "convnextv2_base": ([3, 3, 27, 3], [128, 256, 512, 1024])}


def state_dict_shapes(depths: List[int], dims: List[int], in_chans: int, patch_size: int, num_classes: int) -> Dict[str, tuple]:
    """Keys and shapes of the reference's dense model (``use_orig_stem=False``)."""
    k = patch_size // 8
    sd = {"initial_conv.0.weight": (dims[0], in_chans, 3, 3), "initial_conv.0.bias": (dims[0],),
          "initial_conv.1.weight": (dims[0],), "initial_conv.1.bias": (dims[0],),
          "stem.0.weight": (dims[0], 1, k, k), "stem.0.bias": (dims[0],), "stem.1.weight": (dims[0],), "stem.1.bias": (dims[0],)}
    for i in range(3):
        sd[f"downsample_layers.{i}.0.weight"] = (dims[i],)
        sd[f"downsample_layers.{i}.0.bias"] = (dims[i],)
        sd[f"downsample_layers.{i}.1.weight"] = (dims[i + 1], dims[i], 2, 2)
        sd[f"downsample_layers.{i}.1.bias"] = (dims[i + 1],)
    for i in range(4):
        C = dims[i]
        for j in range(depths[i]):
            p = f"stages.{i}.{j}."
            sd.update({p + "dwconv.weight": (C, 1, 7, 7), p + "dwconv.bias": (C,), p + "norm.weight": (C,), p + "norm.bias": (C,),
                       p + "pwconv1.weight": (4 * C, C), p + "pwconv1.bias": (4 * C,), p + "grn.gamma": (1, 1, 1, 4 * C),
                       p + "grn.beta": (1, 1, 1, 4 * C), p + "pwconv2.weight": (C, 4 * C), p + "pwconv2.bias": (C,)})
    sd.update({"norm.weight": (dims[3],), "norm.bias": (dims[3],), "head.weight": (num_classes, dims[3]), "head.bias": (num_classes,)})
    return sd


def seeded_state_dict(shapes: Dict[str, tuple], seed: int = 0) -> Dict[str, torch.Tensor]:
    """Non-degenerate weights from the torch CPU generator (every tensor drawn in sorted key order): LayerNorm scales
    around 1, everything else N(0, s) with s keeping activations O(1); GRN gamma / beta non-zero so that the GRN path counts."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k in sorted(shapes):
        shape = shapes[k]
        leaf = k.split(".")[-1]
        if len(shape) == 1 and leaf == "weight":
            out[k] = 1.0 + 0.2 * torch.randn(shape, generator=g)
        elif leaf in ("bias", "beta"):
            out[k] = 0.1 * torch.randn(shape, generator=g)
        elif leaf == "gamma":
            out[k] = 0.3 * torch.randn(shape, generator=g)
        else:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            out[k] = torch.randn(shape, generator=g) * fan_in ** -0.5
    return out


def _ln_cf(x, w, b, eps=1e-6):           # channels_first LayerNorm, models/norm_layers.py:26-31
    u = x.mean(1, keepdim=True)
    s = (x - u).pow(2).mean(1, keepdim=True)
    return w[:, None, None] * ((x - u) / torch.sqrt(s + eps)) + b[:, None, None]


def forward_features(sd: Dict[str, torch.Tensor], x: torch.Tensor, depths: List[int], patch_size: int, pooled: bool = True):
    """``ConvNeXtV2.forward_features`` (``models/convnextv2.py:160-172``); ``pooled=False`` returns the last feature map."""
    k = patch_size // 8
    x = F.conv2d(x, sd["initial_conv.0.weight"], sd["initial_conv.0.bias"])                       # 3x3, stride 1, NO padding
    x = F.gelu(_ln_cf(x, sd["initial_conv.1.weight"], sd["initial_conv.1.bias"]))
    x = F.conv2d(x, sd["stem.0.weight"], sd["stem.0.bias"], stride=k, padding=k // 2, groups=x.shape[1])
    x = _ln_cf(x, sd["stem.1.weight"], sd["stem.1.bias"])
    for i in range(4):
        if i > 0:
            x = _ln_cf(x, sd[f"downsample_layers.{i - 1}.0.weight"], sd[f"downsample_layers.{i - 1}.0.bias"])
            x = F.conv2d(x, sd[f"downsample_layers.{i - 1}.1.weight"], sd[f"downsample_layers.{i - 1}.1.bias"], stride=2)
        for j in range(depths[i]):
            p = f"stages.{i}.{j}."
            C = x.shape[1]
            y = F.conv2d(x, sd[p + "dwconv.weight"], sd[p + "dwconv.bias"], padding=3, groups=C).permute(0, 2, 3, 1)
            y = F.layer_norm(y, (C,), sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-6)
            y = F.gelu(F.linear(y, sd[p + "pwconv1.weight"], sd[p + "pwconv1.bias"]))
            Gx = torch.norm(y, p=2, dim=(1, 2), keepdim=True)                                      # per sample, norm_layers.py:41-44
            Nx = Gx / (Gx.mean(dim=-1, keepdim=True) + 1e-4)
            y = sd[p + "grn.gamma"] * (y * Nx) + sd[p + "grn.beta"] + y
            y = F.linear(y, sd[p + "pwconv2.weight"], sd[p + "pwconv2.bias"])
            x = x + y.permute(0, 3, 1, 2)
    if not pooled:
        return x
    return F.layer_norm(x.mean([-2, -1]), (x.shape[1],), sd["norm.weight"], sd["norm.bias"], 1e-6)


def forward(sd, x, depths, patch_size):
    return F.linear(forward_features(sd, x, depths, patch_size), sd["head.weight"], sd["head.bias"])
