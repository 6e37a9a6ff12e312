"""Import the UNMODIFIED reference (``/root/reference``) in the build container.

TEST INFRASTRUCTURE ONLY -- used by ``oracle/make_golden.py`` and by the
``-m "not gpu"`` tests that are skipped when ``/root/reference`` is absent
(it does not exist on the GPU box).

The reference imports packages that are not installed here (timm, kornia,
torchmetrics, tensorboardX, MinkowskiEngine).  They are replaced by minimal
stand-ins *before* import (recipe: SURVEY.md Appendix B):

* ``MinkowskiEngine`` / ``MinkowskiOps``  -> ``oracle.me_shim`` (CPU restatement
  of the six ME ops), so ``models/convnextv2_sparse.py`` and
  ``FCMAE(sparse=True)`` run unmodified;
* ``timm.models.layers.trunc_normal_`` -> ``torch.nn.init.trunc_normal_``;
  ``DropPath`` -> identity (drop_path is 0 on the pretraining path);
* ``kornia.augmentation.RandomCrop`` -> identity crop (inputs are generated at
  model size, ``models/fcmae.py:419-434``);
* ``torchmetrics.Dice`` / ``tensorboardX.SummaryWriter`` -> names only.
"""
from __future__ import annotations

import os
import sys
import types
from argparse import Namespace

import torch
import torch.nn as nn

REFERENCE_ROOT = os.environ.get("MPMAE_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "fcmae.py"))


def _stub(name: str, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


class _IdentityCrop:
    def __init__(self, size, *a, **k):
        self.size = size

    def generate_parameters(self, shape):
        return {}

    def apply_transform(self, x, params, transform=None):
        return x


class _DropPath(nn.Identity):
    def __init__(self, drop_prob=0.0, *a, **k):
        super().__init__()


_loaded = None


def load_reference():
    """Returns a namespace with the reference modules (fcmae, custom_loss, MODALITIES, ...)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    from . import me_shim

    me_shim.install_as_minkowski()
    if "timm" not in sys.modules:
        layers = _stub("timm.models.layers", trunc_normal_=torch.nn.init.trunc_normal_, DropPath=_DropPath)
        models = _stub("timm.models", layers=layers)
        utils = _stub("timm.utils", get_state_dict=lambda m, *a, **k: m.state_dict())
        _stub("timm", models=models, utils=utils)
    if "kornia" not in sys.modules:
        aug = _stub("kornia.augmentation", RandomCrop=_IdentityCrop)
        _stub("kornia", augmentation=aug)
    if "torchmetrics" not in sys.modules:
        _stub("torchmetrics", Dice=object)
    if "tensorboardX" not in sys.modules:
        _stub("tensorboardX", SummaryWriter=object)

    parent = os.path.dirname(REFERENCE_ROOT.rstrip("/"))
    pkg = os.path.basename(REFERENCE_ROOT.rstrip("/"))
    if parent not in sys.path:
        sys.path.insert(0, parent)
    import importlib

    fcmae = importlib.import_module(f"{pkg}.models.fcmae")
    custom_loss = importlib.import_module(f"{pkg}.custom_loss")
    modalities = importlib.import_module(f"{pkg}.MODALITIES")
    _loaded = Namespace(fcmae=fcmae, custom_loss=custom_loss, MODALITIES=modalities, pkg=pkg)
    return _loaded


def import_toplevel(name: str):
    """Import one of the reference's top-level scripts (``helpers``, ``engine_pretrain``: they import each other by bare
    name) with the reference root on ``sys.path`` only for the duration of the import -- the root also holds a ``tests``
    package that would otherwise shadow this repository's in spawned worker processes."""
    load_reference()
    import importlib
    if name in sys.modules:
        return sys.modules[name]
    sys.path.append(REFERENCE_ROOT)
    try:
        return importlib.import_module(name)
    finally:
        sys.path.remove(REFERENCE_ROOT)


def make_args(out_modalities=None, loss_aggr="uncertainty", use_orig_stem=False) -> Namespace:
    """The fields ``main_pretrain.py:175-180`` puts on ``args`` and ``FCMAE`` reads."""
    ref = load_reference()
    M = ref.MODALITIES
    inp = dict(M.INP_MODALITIES)
    out = dict(M.OUT_MODALITIES) if out_modalities is None else {k: M.OUT_MODALITIES[k] for k in out_modalities}
    mods = dict(inp)
    mods.update(out)
    return Namespace(inp_modalities=inp, out_modalities=out, modalities=mods,
                     modalities_full=dict(M.MODALITIES_FULL), use_orig_stem=use_orig_stem,
                     loss_aggr=loss_aggr)


def build_reference_model(model="convnextv2_atto", img_size=56, patch_size=8, out_modalities=None,
                          loss_aggr="uncertainty", norm_pix_loss=True, mask_ratio=0.6, sparse=True,
                          decoder_depth=1, decoder_embed_dim=512, seed=0):
    ref = load_reference()
    args = make_args(out_modalities, loss_aggr)
    loss_fn = ref.custom_loss.UncertaintyWeightingStrategy(len(args.out_modalities)) \
        if loss_aggr == "uncertainty" else None
    torch.manual_seed(seed)
    m = ref.fcmae.__dict__[model](mask_ratio=mask_ratio, decoder_depth=decoder_depth,
                                  decoder_embed_dim=decoder_embed_dim, norm_pix_loss=norm_pix_loss,
                                  patch_size=patch_size, img_size=img_size, args=args, loss_fn=loss_fn,
                                  sparse=sparse)
    return m, args
