"""Checkpoint / weight interchange.

``FCMAE.state_dict()`` already uses the reference's key names and shapes (MinkowskiEngine kernel layout included), so a
checkpoint written by either side loads on the other.  This module adds the two pieces the reference keeps in
``helpers.py``: the sparse -> dense key / layout conversion that finetuning applies to a pretraining checkpoint
(``helpers.remap_checkpoint_keys``, ``helpers.py:668-707``) and a save / load pair for the model + ``FlatAdamW`` state
(``helpers.save_model`` / ``auto_load_model``, ``helpers.py:541-610``).
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, Optional

import torch


def _dense_conv_weight(kernel: torch.Tensor) -> torch.Tensor:
    """MinkowskiEngine kernel -> torch conv weight.  ME orders the taps with the FIRST spatial axis fastest
    (k = kh + ks * kw, ``MinkowskiEngine/src/kernel_region.hpp:199-221``): dense[o, i, kh, kw] = kernel[kh + ks*kw, i, o]."""
    if kernel.dim() == 3:                       # [ks*ks, Cin, Cout] standard convolution
        kv, cin, cout = kernel.shape
        ks = math.isqrt(kv)
        return kernel.reshape(ks, ks, cin, cout).permute(3, 2, 1, 0).contiguous()      # [kw, kh, i, o] -> [o, i, kh, kw]
    if kernel.dim() == 2:                       # [ks*ks, C] depthwise convolution
        kv, c = kernel.shape
        ks = math.isqrt(kv)
        return kernel.reshape(ks, ks, c).permute(2, 1, 0).unsqueeze(1).contiguous()    # [kw, kh, c] -> [c, 1, kh, kw]
    raise ValueError(f"unexpected kernel rank {kernel.dim()}")


def to_dense_state_dict(ckpt: Dict[str, torch.Tensor]) -> "OrderedDict[str, torch.Tensor]":
    """Sparse pretraining checkpoint -> keys / layouts of the dense ``ConvNeXtV2`` (``models/convnextv2.py``).

    Same mapping as ``helpers.remap_checkpoint_keys``: the ``encoder.`` prefix goes, ``*.kernel`` becomes ``*.weight`` in
    torch layout, the ``ln`` / ``linear`` wrapper level of the Minkowski modules goes, biases flatten to 1-D and the GRN
    affine parameters become ``[1, 1, 1, C]``.
    """
    out: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for key, value in ckpt.items():
        parts = key.split(".")
        if parts[0] == "encoder":
            parts = parts[1:]
        if parts[-1] == "kernel":
            out[".".join(parts[:-1] + ["weight"])] = _dense_conv_weight(value)
            continue
        name = ".".join(parts)
        if "ln" in name or "linear" in name:
            parts = parts[:-2] + parts[-1:]                 # drop the wrapper level (norm.ln.weight -> norm.weight)
        elif "backbone.resnet." in name:
            parts = name.split("backbone.resnet.")[1].split(".")
        out[".".join(parts)] = value
    for key in list(out.keys()):
        value = out[key]
        if key.endswith("bias") and value.dim() != 1:
            out[key] = value.reshape(-1)
        elif "grn" in key:
            out[key] = value.unsqueeze(0).unsqueeze(1)
    return out


def load_pretrained_encoder(dense_model, checkpoint: dict, linear_probe: bool = True) -> list:
    """``hubconf.load_custom_checkpoint`` (``hubconf.py:20-75``): put the encoder of a PRETRAINING checkpoint into the dense
    ``convnextv2.ConvNeXtV2``.  A ``head`` of another shape, the decoder, mask token, projection, prediction heads (and the
    loss / pooled-head parameters that only exist in pretraining) are dropped, the rest goes through the sparse -> dense
    remapping; with ``linear_probe=False`` the classifier is re-initialised like the reference does for finetuning.  Returns
    the keys of the dense model the checkpoint did not provide (normally the final ``norm`` and the ``head``)."""
    ck = dict(checkpoint["model"] if "model" in checkpoint else checkpoint)
    own = dense_model.state_dict()
    for k in ("head.weight", "head.bias"):
        if k in ck and ck[k].shape != own[k].shape:
            del ck[k]
    for k in list(ck.keys()):
        if any(s in k for s in ("decoder", "mask_token", "proj", "pred", "loss_fn", "layer_norm_tmp")):
            del ck[k]
    res = dense_model.load_state_dict(to_dense_state_dict(ck), strict=False)
    if res.unexpected_keys:
        raise KeyError(f"checkpoint keys the dense model does not have: {res.unexpected_keys[:5]}")
    if not linear_probe:
        torch.nn.init.trunc_normal_(dense_model.head.weight, std=2e-5)
        torch.nn.init.constant_(dense_model.head.bias, 0.0)
    return list(res.missing_keys)


def save_checkpoint(path: str, model, optimizer=None, epoch: Optional[int] = None, extra: Optional[dict] = None) -> None:
    """``{"model": state_dict, "optimizer": ..., "epoch": ...}`` -- the layout ``helpers.save_model`` writes (``helpers.py:541-547``)."""
    blob = {"model": {k: v.detach().cpu() for k, v in model.state_dict().items()}}
    if optimizer is not None:
        blob["optimizer"] = {k: (v.detach().cpu() if torch.is_tensor(v) else v) for k, v in optimizer.state_dict().items()}
    if epoch is not None:
        blob["epoch"] = int(epoch)
    if extra:
        blob.update(extra)
    torch.save(blob, path)


def load_checkpoint(path: str, model, optimizer=None, strict: bool = True) -> dict:
    """Loads what :func:`save_checkpoint` (or the reference's ``save_model``) wrote; returns the remaining entries."""
    blob = torch.load(path, map_location="cpu", weights_only=False)
    model.load_state_dict(blob["model"], strict=strict)
    if optimizer is not None and "optimizer" in blob and hasattr(optimizer, "load_state_dict"):
        optimizer.load_state_dict(blob["optimizer"])
    return {k: v for k, v in blob.items() if k not in ("model", "optimizer")}


def _is_main_process() -> bool:
    dist = torch.distributed
    return not (dist.is_available() and dist.is_initialized()) or dist.get_rank() == 0


def save_model(args, epoch, model, model_without_ddp, optimizer, loss_scaler, model_ema=None, best: bool = False) -> None:
    """``helpers.save_model`` (``helpers.py:529-565``): rank 0 writes ``<output_dir>/checkpoint-<epoch>.pth`` with the
    reference's entries (``model``, ``optimizer``, ``epoch``, ``scaler``, ``args``) and removes the checkpoint that falls out
    of the ``save_ckpt_num x save_ckpt_freq`` window.  Tensors are moved to the host first (the flat AdamW state is two
    device buffers)."""
    import os

    def host(obj):
        if torch.is_tensor(obj):
            return obj.detach().cpu()
        if isinstance(obj, dict):
            return {k: host(v) for k, v in obj.items()}
        if isinstance(obj, (list, tuple)):
            return type(obj)(host(v) for v in obj)
        return obj

    if not _is_main_process():
        return
    os.makedirs(args.output_dir, exist_ok=True)
    blob = {"model": host(model_without_ddp.state_dict()), "optimizer": host(optimizer.state_dict()), "epoch": epoch,
            "scaler": host(loss_scaler.state_dict()), "args": args}
    if model_ema is not None:
        blob["model_ema"] = host(model_ema.state_dict())
    torch.save(blob, os.path.join(args.output_dir, "checkpoint-%s.pth" % str(epoch)))
    if isinstance(epoch, int):
        old = os.path.join(args.output_dir, "checkpoint-%s.pth" % (epoch - args.save_ckpt_num * args.save_ckpt_freq))
        if os.path.exists(old):
            os.remove(old)


def auto_load_model(args, model, model_without_ddp, optimizer, loss_scaler, model_ema=None) -> None:
    """``helpers.auto_load_model`` (``helpers.py:568-610``): with ``args.auto_resume`` and no explicit ``args.resume`` the
    newest ``checkpoint-<int>.pth`` of ``args.output_dir`` is taken; model, optimizer, scaler are restored and
    ``args.start_epoch`` is set to the epoch after the checkpoint's."""
    import glob
    import os

    if getattr(args, "auto_resume", False) and len(getattr(args, "resume", "") or "") == 0:
        latest = -1
        for ckpt in glob.glob(os.path.join(args.output_dir, "checkpoint-*.pth")):
            t = ckpt.split("-")[-1].split(".")[0]
            if t.isdigit():
                latest = max(int(t), latest)
        if latest >= 0:
            args.resume = os.path.join(args.output_dir, "checkpoint-%d.pth" % latest)
    if not getattr(args, "resume", ""):
        return
    if args.resume.startswith("https"):
        blob = torch.hub.load_state_dict_from_url(args.resume, map_location="cpu", check_hash=True)
    else:
        blob = torch.load(args.resume, map_location="cpu", weights_only=False)
    model_without_ddp.load_state_dict(blob["model"])
    if "optimizer" in blob and "epoch" in blob:
        optimizer.load_state_dict(blob["optimizer"])
        if not isinstance(blob["epoch"], str):
            args.start_epoch = blob["epoch"] + 1
        if "scaler" in blob:
            loss_scaler.load_state_dict(blob["scaler"])
