"""Drop-in for the reference's ``models.fcmae.FCMAE`` backed by the B200-native step library.

Mirrors the reference module surface (``models/fcmae.py:27-496``): same constructor arguments, the
same ``forward`` return tuple ``(loss, pred, mask, loss_dict, log_vars, normalized_loss_list)``, the
same state-dict keys and shapes (``SURVEY.md`` section 8a "State-dict surface"), the same factory
functions ``convnextv2_atto ... convnextv2_huge`` (``models/fcmae.py:459-496``), so that
``main_pretrain.py:270-281`` / ``engine_pretrain.train_one_epoch`` can use it unchanged.

What is different underneath: one call into ``libmpmae.so`` runs mask -> sparse encoder -> decoder ->
heads -> losses as a fixed sequence of hand-written sm_100a kernels, and ``loss.backward()`` runs
the hand-derived backward (no autograd graph over the model).  All parameters are views into ONE
flat fp32 buffer and all gradients views into ONE flat gradient buffer (the unit of the NCCL
all-reduce, ``SURVEY.md`` section 8e).  There is no PyTorch fallback: without the library, import fails.
"""
from __future__ import annotations

import ctypes as C
from argparse import Namespace
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from . import _native as nat
from .dist import FlatGradReducer

PIXEL_CONTINUOUS = ("sentinel2", "sentinel1", "aster", "canopy_height_eth")
PIXEL_CATEGORICAL = ("dynamic_world", "esa_worldcover")
IMAGE_CATEGORICAL = ("biome", "eco_region")
IMAGE_CONTINUOUS = ("lat", "lon", "month", "era5")
# class counts, models/fcmae.py:70-91
N_CLASSES = {"dynamic_world": 9, "esa_worldcover": 11, "biome": 14, "eco_region": 846}


def modality_kind(name: str) -> int:
    if name in PIXEL_CONTINUOUS:
        return nat.PIXEL_CONTINUOUS
    if name in PIXEL_CATEGORICAL:
        return nat.PIXEL_CATEGORICAL
    if name in IMAGE_CATEGORICAL:
        return nat.IMAGE_CATEGORICAL
    if name in IMAGE_CONTINUOUS:
        return nat.IMAGE_CONTINUOUS
    raise ValueError(f"unsupported output modality {name!r} (models/fcmae.py:126-151 lists the supported ones)")


class UncertaintyWeightingStrategy(nn.Module):
    """Parameter holder with the reference's key ``loss_fn.log_vars`` (``custom_loss.py:10-30``).

    The weighting itself ``(exp(-s) * L + s) * [L != 0]`` runs inside the fused loss kernel; calling
    this module on a tensor of losses evaluates the same formula in torch for users who want it.
    """

    def __init__(self, tasks: int):
        super().__init__()
        self.tasks = tasks
        self.log_vars = nn.Parameter(torch.zeros(tasks))

    def forward(self, task_losses):
        lt = torch.stack(list(task_losses)) if not torch.is_tensor(task_losses) else task_losses
        w = (torch.exp(-self.log_vars) * lt + self.log_vars) * (lt != 0.0)
        return w, self.log_vars.tolist()


class _Node(nn.Module):
    """Name-only container: the module tree exists to reproduce the reference's state-dict keys."""


class _StepFunction(torch.autograd.Function):
    """loss = f(parameters): forward and backward are single calls into the C ABI."""

    @staticmethod
    def forward(ctx, model, run, token, *params):
        ctx.model, ctx.run = model, run
        model._native_forward(run)
        T = run["T"]
        losses = run["losses"]
        total = losses[2 * T].clone()
        ctx.mark_non_differentiable(losses)
        return total, losses

    @staticmethod
    def backward(ctx, grad_total, _grad_losses):
        ctx.model._native_backward(ctx.run, grad_total)
        # parameter gradients were written straight into the flat buffer (p.grad are views of it); the only
        # gradient handed to autograd is the zero of the DDP token
        return (None, None, grad_total.new_zeros(1)) + (None,) * len(ctx.model._param_list)


class _EncoderFunction(torch.autograd.Function):
    """``forward_encoder`` with autograd (``models/fcmae.py:242-247``): features = f(parameters)."""

    @staticmethod
    def forward(ctx, model, run, token, *params):
        ctx.model, ctx.run = model, run
        model._native_forward_encoder(run)
        feats = model.encoder_features(run)
        ctx.mark_non_differentiable(run["mask"])
        return feats, run["mask"]

    @staticmethod
    def backward(ctx, dfeats, _dmask):
        model, run = ctx.model, ctx.run
        B, C3 = dfeats.shape[0], dfeats.shape[1]
        rows = dfeats.float().permute(0, 2, 3, 1).reshape(B * model.num_patches, C3)[run["mask"].reshape(-1) == 0].contiguous()
        model._stepwise_backward(run, nat.BWD_ENCODER, d_x3=rows)
        return (None, None, dfeats.new_zeros(1)) + (None,) * len(model._param_list)


class _DecoderFunction(torch.autograd.Function):
    """``forward_decoder`` with autograd (``models/fcmae.py:249-265``): predictions = f(encoder features, parameters)."""

    @staticmethod
    def forward(ctx, model, run, x, token, *params):
        ctx.model, ctx.run = model, run
        B, C3 = x.shape[0], x.shape[1]
        rows = x.detach().float().permute(0, 2, 3, 1).reshape(B * model.num_patches, C3)[run["mask"].reshape(-1) == 0]
        model.tap(f"stage3.block{model.depths[3] - 1}.y", run).copy_(rows)      # [B*V, C3], ascending patch index
        model._stages(run, nat.STAGE_MASK | nat.STAGE_DECODER)
        return run["pred_pixel"], run["pred_image"]

    @staticmethod
    def backward(ctx, dpix, dimg):
        model, run = ctx.model, ctx.run
        B, L, C3 = run["B"], model.num_patches, model.dims[3]
        dpix = torch.zeros_like(run["pred_pixel"]) if dpix is None else dpix.float().contiguous()
        dimg = torch.zeros_like(run["pred_image"]) if dimg is None else dimg.float().contiguous()
        vis = run["mask"].reshape(-1) == 0
        d_x3 = torch.empty(int(vis.sum()), C3, device=dpix.device)
        model._stepwise_backward(run, nat.BWD_DECODER, dpred_pixel=dpix, dpred_image=dimg, d_x3=d_x3)
        dx = torch.zeros(B * L, C3, device=dpix.device)
        dx[vis] = d_x3
        G = model.img_size // model.patch_size
        return (None, None, dx.view(B, G, G, C3).permute(0, 3, 1, 2), dpix.new_zeros(1)) + (None,) * len(model._param_list)


class _LossFunction(torch.autograd.Function):
    """``forward_loss`` with autograd (``models/fcmae.py:267-412``): total loss = f(predictions, log_vars)."""

    @staticmethod
    def forward(ctx, model, run, token, *preds):
        ctx.model, ctx.run = model, run
        model._pack_preds(run, preds)
        model._stages(run, nat.STAGE_MASK | nat.STAGE_LOSS)
        losses = run["losses"]
        ctx.mark_non_differentiable(losses)
        return losses[2 * run["T"]].clone(), losses

    @staticmethod
    def backward(ctx, grad_total, _grad_losses):
        model, run = ctx.model, ctx.run
        dpp, dpi = torch.empty_like(run["pred_pixel"]), torch.empty_like(run["pred_image"])
        model._stepwise_backward(run, nat.BWD_LOSS, dpred_pixel=dpp, dpred_image=dpi, grad_out=grad_total)
        grads = model._pred_dict({"plan": run["plan"], "B": run["B"], "pred_pixel": dpp, "pred_image": dpi})
        return (None, None, grad_total.new_zeros(1)) + tuple(grads[m] for m in model.out_modalities)


class FCMAE(nn.Module):
    """Fully convolutional multi-pretext masked autoencoder, native B200 step (``models/fcmae.py:27``)."""

    #: DDP must not reduce the flat-buffer views itself; the module all-reduces the flat gradient
    #: buffer inside backward (``main_pretrain.py:306-310`` wraps whatever it is given).
    _ddp_params_and_buffers_to_ignore: List[str] = []

    def __init__(self, img_size: int = 112, depths: List[int] = None, dims: List[int] = None,
                 decoder_depth: int = 1, decoder_embed_dim: int = 512, patch_size: int = 16,
                 mask_ratio: float = 0.6, norm_pix_loss: bool = False, args: Namespace = None,
                 loss_fn=None, sparse: bool = True, gemm_backend: Optional[int] = None):
        super().__init__()
        if not sparse:
            raise NotImplementedError("the native step implements the sparse (masked) encoder only; the reference's "
                                      "dense CPU path (sparse=False) is a different network (SURVEY.md section 0.1)")
        if args is None:
            raise ValueError("args with inp/out modalities is required (main_pretrain.py:175-180)")
        if getattr(args, "use_orig_stem", False):
            raise NotImplementedError("use_orig_stem=True is not on the pretraining path (convnextv2_sparse.py:99-111)")
        depths = list(depths) if depths is not None else [3, 3, 9, 3]
        dims = list(dims) if dims is not None else [96, 192, 384, 768]
        self.args = args
        self.img_size, self.patch_size, self.mask_ratio = img_size, patch_size, mask_ratio
        self.depths, self.dims = depths, dims
        self.decoder_depth, self.decoder_embed_dim = decoder_depth, decoder_embed_dim
        self.norm_pix_loss = norm_pix_loss
        self.num_patches = (img_size // patch_size) ** 2
        self.imgs_size = img_size
        self.loss_aggr = args.loss_aggr
        self.out_modalities = list(args.out_modalities.keys())
        s2 = args.modalities["sentinel2"]
        self.in_chans = len(args.modalities_full["sentinel2"]) if s2 == "all" else len(s2)
        self.out_chans: Dict[str, int] = {}
        for m, bands in args.modalities.items():          # models/fcmae.py:70-91
            if m in N_CLASSES:
                self.out_chans[m] = N_CLASSES[m]
            else:
                self.out_chans[m] = len(args.modalities_full[m]) if bands == "all" else len(bands)
        if self.loss_aggr == "uncertainty" and loss_fn is None:
            raise ValueError("loss_aggr='uncertainty' needs loss_fn with a log_vars parameter (custom_loss.py:10-17)")
        self.gemm_backend = 3 if gemm_backend is None else int(gemm_backend)

        self._plans: Dict[int, nat.Plan] = {}
        plan = self._plan(1)
        self._layout = plan.params()
        self._n_flat = plan.param_total
        self._flat = torch.zeros(self._n_flat, dtype=torch.float32)
        self._gacc: Optional[torch.Tensor] = None
        self._gstep: Optional[torch.Tensor] = None
        self._grad_views: Optional[list] = None      # persistent views of _gacc, one per parameter (built once)
        self._grads_cleared = False                  # zero_grad(set_to_none=True) was called: next backward overwrites
        self._workspace: Optional[torch.Tensor] = None
        self._flags: Optional[torch.Tensor] = None
        self._param_list: List[nn.Parameter] = []
        self._param_slices: List[Tuple[int, int, Tuple[int, ...]]] = []
        self.allreduce_chunks = 4
        self.noise_override: Optional[torch.Tensor] = None
        self.backward_in_parts = False      # world_size > 1 always runs the backward in parts (overlapped all-reduce)
        self.reduce_gradients = True        # False: the backward stays rank-local even under torch.distributed (checks)
        self.last_run: Optional[dict] = None

        # ---- module tree with the reference's names
        self.encoder = _Node()
        self.proj = _Node()
        self.decoder_dict = nn.ModuleDict()
        self.pred_dict = nn.ModuleDict()
        self.loss_fn = loss_fn
        decoder_blocks = [_Node() for _ in range(decoder_depth)]
        for m in self.out_modalities:
            modality_kind(m)
            seq = _Node()
            for k, blk in enumerate(decoder_blocks):
                seq.add_module(str(k), blk)
            self.decoder_dict[m] = seq
            self.pred_dict[m] = _Node()
        for name, shape, off, _decay in self._layout:
            numel = 1
            for s in shape:
                numel *= s
            p = nn.Parameter(self._flat[off:off + numel].view(shape))
            self._param_list.append(p)
            self._param_slices.append((off, numel, shape))
            self._register(name, p, decoder_blocks)
        # DistributedDataParallel (main_pretrain.py:306-310) needs one parameter to manage; everything else is
        # ignored by it and reduced here as one flat buffer.  The token is kept out of the state dict.
        self._ddp_token = nn.Parameter(torch.zeros(1))
        self._ddp_params_and_buffers_to_ignore = [k for k in self.state_dict().keys() if k != "_ddp_token"]
        self._register_state_dict_hook(_drop_token)
        self._register_load_state_dict_pre_hook(_add_token)
        self.reset_parameters()

    # ------------------------------------------------------------------ construction helpers
    def _cfg(self, batch: int) -> nat.Cfg:
        c = nat.Cfg()
        c.batch, c.img_size, c.patch_size, c.in_chans = batch, self.img_size, self.patch_size, self.in_chans
        for i in range(4):
            c.depths[i], c.dims[i] = self.depths[i], self.dims[i]
        c.dec_dim, c.dec_depth = self.decoder_embed_dim, self.decoder_depth
        c.mask_ratio = float(self.mask_ratio)
        c.loss_aggr = 1 if self.loss_aggr == "uncertainty" else 0
        c.n_mod = len(self.out_modalities)
        for i, m in enumerate(self.out_modalities):
            c.mod_kind[i] = modality_kind(m)
            c.mod_chans[i] = self.out_chans[m]
            c.mod_norm_pix[i] = 1 if (self.norm_pix_loss and m == "sentinel2") else 0
        c.gemm_backend = self.gemm_backend
        return c

    def _plan(self, batch: int) -> nat.Plan:
        pl = self._plans.get(batch)
        if pl is None or pl.cfg.gemm_backend != self.gemm_backend or abs(pl.cfg.mask_ratio - self.mask_ratio) > 1e-7:
            pl = nat.Plan(self._cfg(batch))
            self._plans[batch] = pl
        return pl

    def _register(self, name: str, p: nn.Parameter, decoder_blocks) -> None:
        parts = name.split(".")
        if parts[0] == "decoder":                       # shared block, aliased under every decoder_dict[mod]
            node = decoder_blocks[int(parts[1])]
            parts = parts[2:]
        elif parts[0] == "pred_dict":
            node = self.pred_dict[self.out_modalities[int(parts[1][1:])]]
            parts = parts[2:]
        elif parts[0] == "mask_token":
            self.mask_token = p
            return
        elif parts[0] == "loss_fn":
            # the caller's loss module keeps its key; its parameter becomes a view of the flat buffer
            with torch.no_grad():
                p.copy_(self.loss_fn.log_vars.detach().reshape(p.shape))
            self.loss_fn.log_vars = p
            return
        else:
            node = self
        for part in parts[:-1]:
            child = getattr(node, part, None) if part in node._modules else None
            if child is None:
                child = _Node()
                node.add_module(part, child)
            node = child
        node.register_parameter(parts[-1], p)

    @torch.no_grad()
    def reset_parameters(self) -> None:
        """The reference initialiser, ``models/fcmae.py:157-175`` (trunc_normal via torch.nn.init)."""
        tn = torch.nn.init.trunc_normal_
        for (name, shape, _o, _d), p in zip(self._layout, self._param_list):
            leaf = name.split(".")[-1]
            if name.startswith("loss_fn"):
                continue
            if name == "mask_token":
                p.normal_(std=0.02)
            elif leaf in ("gamma", "beta"):
                p.zero_()
            elif leaf == "bias":
                p.zero_()
            elif leaf == "kernel":
                tn(p, std=0.02 if p.dim() == 3 else 1.0)      # MinkowskiConvolution vs depthwise
            elif leaf == "weight" and p.dim() == 1:
                p.fill_(1.0)                                   # LayerNorm scales
            elif name.startswith("encoder") and leaf == "weight":
                tn(p)                                          # MinkowskiLinear: default std 1.0
            elif p.dim() == 4:
                tn(p.view(p.shape[0], -1))                     # nn.Conv2d: std 1.0 on the flattened view
            else:
                tn(p, std=0.02)                                # nn.Linear

    # ------------------------------------------------------------------ flat storage management
    def _apply(self, fn, recurse=True):
        new_flat = fn(self._flat)
        if new_flat.dtype != torch.float32:
            raise TypeError("the native step is fp32; parameter dtype conversion is not supported")
        super()._apply(fn, recurse)
        self._flat = new_flat.contiguous()
        for p, (off, numel, shape) in zip(self._param_list, self._param_slices):
            p.data = self._flat[off:off + numel].view(shape)
            p.grad = None
        self._gacc = self._gstep = self._workspace = self._flags = None
        self._grad_views, self._grads_cleared = None, False
        return self

    @property
    def flat_params(self) -> torch.Tensor:
        return self._flat

    @property
    def flat_grads(self) -> Optional[torch.Tensor]:
        return self._gacc

    def decay_mask(self) -> torch.Tensor:
        """uint8 per flat element: 1 where AdamW weight decay applies (main_pretrain.py:312-319 rule)."""
        m = torch.zeros(self._n_flat, dtype=torch.uint8)
        for (name, shape, off, decay), (_o, numel, _s) in zip(self._layout, self._param_slices):
            if decay:
                m[off:off + numel] = 1
        return m.to(self._flat.device)

    def zero_grad(self, set_to_none: bool = True) -> None:
        """Gradients are views of ONE flat buffer: ``set_to_none=True`` (the reference's default, ``helpers.py:487`` /
        ``optimizer.zero_grad()``) only marks that buffer as cleared -- the next backward overwrites it -- instead of
        dropping and re-creating hundreds of view tensors every step (2 ms of host time at cfg2).  ``p.grad`` keeps
        pointing at the (stale until the next backward) view."""
        if set_to_none:
            if self._gacc is None:
                super().zero_grad(set_to_none=True)
            self._grads_cleared = True
        else:
            if self._gacc is not None:
                self._gacc.zero_()
            self._grads_cleared = False

    def _grad_views_fresh(self) -> bool:
        """True when the next backward starts a new accumulation: zero_grad(set_to_none=True), the first step, or an
        external optimizer that set every p.grad to None."""
        if self._grads_cleared or self._gacc is None:
            return True
        first, last = self._param_list[0], self._param_list[-1]
        if first.grad is None and last.grad is None:        # torch.optim's own zero_grad(set_to_none=True)
            return all(p.grad is None for p in self._param_list)
        return False

    def _bind_grads(self) -> None:
        if self._grad_views is None or self._grad_views[0].untyped_storage().data_ptr() != self._gacc.untyped_storage().data_ptr():
            self._grad_views = [self._gacc[off:off + numel].view(shape) for (off, numel, shape) in self._param_slices]
        for p, g in zip(self._param_list, self._grad_views):
            if p.grad is not g:
                p.grad = g
        self._grads_cleared = False

    # ------------------------------------------------------------------ native calls
    def _device_check(self, t: torch.Tensor) -> None:
        if not t.is_cuda:
            raise RuntimeError("the native MP-MAE step runs on CUDA (sm_100a) tensors only; there is no CPU path")
        if self._flat.device != t.device:
            raise RuntimeError(f"model is on {self._flat.device}, input on {t.device}: call model.to(device) first")

    def _targets(self, imgs_dict: Dict[str, torch.Tensor]) -> List[torch.Tensor]:
        targets = []
        for m in self.out_modalities:
            t = imgs_dict[m]
            kind = modality_kind(m)
            t = t.contiguous()
            if kind in (nat.PIXEL_CATEGORICAL, nat.IMAGE_CATEGORICAL):
                t = t.long() if t.dtype != torch.int64 else t
            else:
                t = t.float() if t.dtype != torch.float32 else t
            targets.append(t)
        return targets

    def _prepare(self, imgs_dict: Optional[Dict[str, torch.Tensor]], with_targets: bool = True, with_preds: bool = True,
                 mask: Optional[torch.Tensor] = None) -> dict:
        """Buffers of one native call.  ``mask`` (step-wise methods): a given {0,1} mask is passed as the noise -- the
        stable ranking of the mask kernel reproduces it -- instead of drawing new noise."""
        if imgs_dict is not None and "sentinel2" in imgs_dict:
            imgs = imgs_dict["sentinel2"]
            self._device_check(imgs)
            B, S = imgs.shape[0], imgs.shape[2]
            if S != self.img_size:
                imgs_dict = self._random_crop(imgs_dict)
                imgs = imgs_dict["sentinel2"]
            dev = imgs.device
        else:
            imgs = None
            self._device_check(mask)
            B, dev = mask.shape[0], mask.device
        plan = self._plan(B)
        T = len(self.out_modalities)
        run = {"B": B, "T": T, "plan": plan, "dev": dev}
        run["s2"] = imgs.contiguous().float() if imgs is not None else None
        run["targets"] = self._targets(imgs_dict) if with_targets else []
        L = self.num_patches
        if mask is not None:
            run["noise"] = mask.to(dev).float().contiguous()
        elif self.noise_override is not None:                 # parity tests inject the oracle's noise
            run["noise"] = self.noise_override.to(dev).float().contiguous()
        else:
            run["noise"] = torch.randn(B, L, device=dev)      # the reference's RNG call, fcmae.py:220
        if run["noise"].shape != (B, L):
            raise ValueError(f"mask / noise shape {tuple(run['noise'].shape)}, expected {(B, L)}")
        run["mask"] = torch.empty(B, L, device=dev)
        run["pred_pixel"] = torch.empty(B * L, max(plan.npix, 1), device=dev) if with_preds else None
        run["pred_image"] = torch.empty(B, max(plan.nimg, 1), device=dev) if with_preds else None
        run["losses"] = torch.zeros(2 * T + 1, device=dev)
        need = plan.workspace_bytes
        if self._workspace is None or self._workspace.numel() * 4 < need or self._workspace.device != dev:
            self._workspace = torch.empty((need + 3) // 4, dtype=torch.float32, device=dev)
        if self._flags is None or self._flags.device != dev:
            self._flags = torch.zeros(4, dtype=torch.int32, device=dev)
        return run

    def _io(self, run: dict, grads: Optional[torch.Tensor] = None, grad_out: Optional[torch.Tensor] = None) -> nat.IO:
        io = nat.IO()
        io.params = self._flat.data_ptr()
        io.grads = grads.data_ptr() if grads is not None else None
        io.workspace = self._workspace.data_ptr()
        io.workspace_bytes = self._workspace.numel() * 4
        io.noise = run["noise"].data_ptr()
        io.s2_input = run["s2"].data_ptr() if run["s2"] is not None else None
        for i, t in enumerate(run["targets"]):
            io.targets[i] = t.data_ptr()
        io.mask = run["mask"].data_ptr()
        io.pred_pixel = run["pred_pixel"].data_ptr() if run["pred_pixel"] is not None else None
        io.pred_image = run["pred_image"].data_ptr() if run["pred_image"] is not None else None
        io.losses = run["losses"].data_ptr()
        io.grad_out = grad_out.data_ptr() if grad_out is not None else None
        io.flags = self._flags.data_ptr()
        return io

    def _native_forward(self, run: dict) -> None:
        io = self._io(run)
        stream = torch.cuda.current_stream(run["dev"]).cuda_stream
        with torch.cuda.device(run["dev"]):
            nat.check(nat.lib.mpmae_forward(run["plan"].handle, C.byref(io), C.c_void_p(stream)), "mpmae_forward")

    def _native_backward(self, run: dict, grad_total: torch.Tensor) -> None:
        dev = run["dev"]
        world = 1
        dist = torch.distributed
        if dist.is_available() and dist.is_initialized() and self.reduce_gradients:
            world = dist.get_world_size()
        go = grad_total.detach().reshape(1).float().contiguous()
        if world > 1:
            go = go / world                                   # mean over ranks folded into the backward seed
        fresh = self._grad_views_fresh()
        if self._gacc is None or self._gacc.device != dev:
            self._gacc = torch.zeros(self._n_flat, device=dev)
            fresh = True
        if fresh:
            target = self._gacc
        else:
            if self._gstep is None:
                self._gstep = torch.empty(self._n_flat, device=dev)
            target = self._gstep
        target.zero_()
        io = self._io(run, grads=target, grad_out=go)
        stream = torch.cuda.current_stream(dev).cuda_stream
        plan = run["plan"]
        with torch.cuda.device(dev):
            if world == 1 and not self.backward_in_parts:
                nat.check(nat.lib.mpmae_backward(plan.handle, C.byref(io), C.c_void_p(stream)), "mpmae_backward")
            else:
                # three parts in reverse layer order; the finished slice of the flat buffer is all-reduced (NCCL over
                # NVLink / NVSwitch, on the backend's stream) while the next part computes
                reducer = FlatGradReducer(plan.backward_ranges(), self._n_flat)
                for part in range(3):
                    nat.check(nat.lib.mpmae_backward_part(plan.handle, C.byref(io), part, C.c_void_p(stream)),
                              "mpmae_backward_part")
                    reducer.reduce_part(target, part)
                reducer.wait()
        if not fresh:
            self._gacc.add_(self._gstep)
        self._bind_grads()

    def _native_forward_encoder(self, run: dict) -> None:
        io = self._io(run)
        stream = torch.cuda.current_stream(run["dev"]).cuda_stream
        with torch.cuda.device(run["dev"]):
            nat.check(nat.lib.mpmae_forward_encoder(run["plan"].handle, C.byref(io), C.c_void_p(stream)), "mpmae_forward_encoder")

    def _stepwise_backward(self, run: dict, which: int, dpred_pixel=None, dpred_image=None, d_x3=None, grad_out=None) -> None:
        """One of the three step-wise backward calls (``mpmae_backward_step``).  Each accumulates its parameter gradients
        into a scratch flat buffer that is then added to (first call after ``zero_grad``: copied into) the flat gradient
        buffer, all-reduced first under ``torch.distributed`` (mean over ranks, DDP semantics)."""
        dev = run["dev"]
        if self._gstep is None or self._gstep.device != dev:
            self._gstep = torch.empty(self._n_flat, device=dev)
        tmp = self._gstep
        tmp.zero_()
        go = grad_out.detach().reshape(1).float().contiguous() if grad_out is not None else None
        io = self._io(run, grads=tmp, grad_out=go)
        ptr = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
        stream = torch.cuda.current_stream(dev).cuda_stream
        with torch.cuda.device(dev):
            nat.check(nat.lib.mpmae_backward_step(run["plan"].handle, C.byref(io), which, ptr(dpred_pixel), ptr(dpred_image),
                                                  ptr(d_x3), C.c_void_p(stream)), "mpmae_backward_step")
        dist = torch.distributed
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1 and self.reduce_gradients:
            dist.all_reduce(tmp)
            tmp /= dist.get_world_size()
        fresh = self._grad_views_fresh()
        if self._gacc is None or self._gacc.device != dev:
            self._gacc = torch.zeros(self._n_flat, device=dev)
            fresh = True
        if fresh:
            self._gacc.copy_(tmp)
        else:
            self._gacc.add_(tmp)
        self._bind_grads()

    def _pack_preds(self, run: dict, preds) -> None:
        """Per-modality prediction tensors (``out_modalities`` order) into the packed buffers the loss kernels read."""
        plan, B = run["plan"], run["B"]
        G = self.img_size // self.patch_size
        pix = run["pred_pixel"].view(B, G, G, -1)
        for i, (m, t) in enumerate(zip(self.out_modalities, preds)):
            off = plan.col_offset(i)
            if modality_kind(m) in (nat.PIXEL_CONTINUOUS, nat.PIXEL_CATEGORICAL):
                n = self.patch_size ** 2 * self.out_chans[m]
                pix[..., off:off + n].copy_(t.detach().permute(0, 2, 3, 1))
            else:
                run["pred_image"][:, off:off + self.out_chans[m]].copy_(t.detach())

    def _random_crop(self, imgs_dict):
        """Same random window per sample for all pixel-wise modalities (``models/fcmae.py:419-434``)."""
        S = self.img_size
        x = imgs_dict["sentinel2"]
        B, _, H, W = x.shape
        if H < S or W < S:
            raise ValueError(f"input {H}x{W} smaller than img_size {S}")
        dev = x.device
        ys = torch.randint(0, H - S + 1, (B,), device=dev)
        xs = torch.randint(0, W - S + 1, (B,), device=dev)
        ar = torch.arange(S, device=dev)
        rows = (ys[:, None] + ar[None, :])[:, :, None]
        cols = (xs[:, None] + ar[None, :])[:, None, :]
        bidx = torch.arange(B, device=dev)[:, None, None]
        out = dict(imgs_dict)
        for m, t in imgs_dict.items():
            if torch.is_tensor(t) and t.dim() == 4 and t.shape[2] == H and t.shape[3] == W:
                out[m] = t[bidx, :, rows, cols].permute(0, 3, 1, 2).contiguous()
        return out

    # ------------------------------------------------------------------ reference surface
    def gen_random_mask(self, x: torch.Tensor, mask_ratio: float) -> torch.Tensor:
        """``models/fcmae.py:214-231`` with the same torch call sequence (mask indices bit-exact)."""
        N = x.shape[0]
        L = (x.shape[2] // self.patch_size) ** 2
        len_keep = int(L * (1 - mask_ratio))
        noise = torch.randn(N, L, device=x.device)
        ids_restore = torch.argsort(torch.argsort(noise, dim=1), dim=1)
        mask = torch.ones([N, L], device=x.device)
        mask[:, :len_keep] = 0
        return torch.gather(mask, dim=1, index=ids_restore)

    def patchify(self, imgs: torch.Tensor, modality: str) -> torch.Tensor:
        """``models/fcmae.py:180-197``."""
        p = self.patch_size
        channels = 1 if modality in PIXEL_CATEGORICAL else self.out_chans[modality]
        h = w = imgs.shape[2] // p
        x = imgs.reshape(imgs.shape[0], channels, h, p, w, p)
        return torch.einsum("nchpwq->nhwpqc", x).reshape(imgs.shape[0], h * w, p * p * channels)

    def forward_encoder(self, imgs: torch.Tensor, mask_ratio: float) -> Tuple[torch.Tensor, torch.Tensor]:
        """``models/fcmae.py:242-247``: (dense features [B, C3, G, G] with zeros at masked cells, mask).  Autograd-connected
        like the reference's when gradients are enabled: ``forward_encoder -> forward_decoder -> forward_loss`` composes into
        the same training step as ``forward`` (which stays the fast path: one native call each way)."""
        if abs(mask_ratio - self.mask_ratio) > 1e-7:
            self.mask_ratio = mask_ratio
        run = self._prepare({"sentinel2": imgs}, with_targets=False, with_preds=False)
        self.last_run = run
        if torch.is_grad_enabled():
            feats, mask = _EncoderFunction.apply(self, run, self._ddp_token, *self._param_list)
            return feats, mask
        self._native_forward_encoder(run)
        return self.encoder_features(run), run["mask"]

    def _stages(self, run: dict, stages: int) -> None:
        io = self._io(run)
        stream = torch.cuda.current_stream(run["dev"]).cuda_stream
        with torch.cuda.device(run["dev"]):
            nat.check(nat.lib.mpmae_forward_stages(run["plan"].handle, C.byref(io), stages, C.c_void_p(stream)),
                      "mpmae_forward_stages")

    def _check_mask(self, mask: torch.Tensor) -> None:
        V = int(self.num_patches * (1 - self.mask_ratio))
        kept = (mask == 0).sum(dim=1)
        if not bool(((mask == 0) | (mask == 1)).all()) or not bool((kept == V).all()):
            raise ValueError(f"mask must be {{0,1}} with exactly {V} visible patches per sample (mask_ratio "
                             f"{self.mask_ratio}); got between {int(kept.min())} and {int(kept.max())}")

    def _pred_dict(self, run: dict) -> Dict[str, torch.Tensor]:
        plan, B = run["plan"], run["B"]
        G = self.img_size // self.patch_size
        pred = {}
        for i, m in enumerate(self.out_modalities):
            off = plan.col_offset(i)
            if modality_kind(m) in (nat.PIXEL_CONTINUOUS, nat.PIXEL_CATEGORICAL):
                n = self.patch_size ** 2 * self.out_chans[m]
                pred[m] = run["pred_pixel"].view(B, G, G, -1)[..., off:off + n].permute(0, 3, 1, 2)
            else:
                pred[m] = run["pred_image"][:, off:off + self.out_chans[m]]
        return pred

    def forward_decoder(self, x: torch.Tensor, mask: torch.Tensor) -> Dict[str, torch.Tensor]:
        """``models/fcmae.py:249-265``: dense encoder features ``[B, C3, G, G]`` + mask -> predictions of every output
        modality.  Only the visible cells of ``x`` are read (the reference overwrites the masked ones with the mask
        token).  Autograd-connected (gradients flow to ``x`` and to the decoder / head parameters) when gradients are enabled."""
        self._device_check(x)
        self._check_mask(mask)
        run = self._prepare(None, with_targets=False, mask=mask)
        run["mask"] = mask.detach().to(run["dev"]).float().clone()         # the mask kernel reproduces it; needed before that
        self.last_run = run
        if torch.is_grad_enabled():
            pp, pi = _DecoderFunction.apply(self, run, x, self._ddp_token, *self._param_list)
            return self._pred_dict({"plan": run["plan"], "B": run["B"], "pred_pixel": pp, "pred_image": pi})
        B, C3 = x.shape[0], x.shape[1]
        rows = x.float().permute(0, 2, 3, 1).reshape(B * self.num_patches, C3)[mask.reshape(-1) == 0]
        self.tap(f"stage3.block{self.depths[3] - 1}.y", run).copy_(rows)      # [B*V, C3], ascending patch index
        self._stages(run, nat.STAGE_MASK | nat.STAGE_DECODER)
        return self._pred_dict(run)

    def forward_loss(self, imgs_dict: Dict[str, torch.Tensor], preds: Dict[str, torch.Tensor], mask: torch.Tensor):
        """``models/fcmae.py:267-412``: per-modality reconstruction losses of given predictions + their aggregate;
        returns ``(loss, loss_dict, log_vars, normalized_loss_list)``.  Autograd-connected (gradients flow to ``preds`` and to
        ``loss_fn.log_vars``) when gradients are enabled."""
        self._device_check(mask)
        self._check_mask(mask)
        for m in self.out_modalities:
            if modality_kind(m) in (nat.PIXEL_CONTINUOUS, nat.PIXEL_CATEGORICAL) and \
                    tuple(imgs_dict[m].shape[-2:]) != (self.img_size, self.img_size):
                raise ValueError(f"forward_loss: target {m!r} is {tuple(imgs_dict[m].shape)}; expected img_size "
                                 f"{self.img_size} (models/fcmae.py:419-434 crops in forward, before the loss)")
        run = self._prepare(None, with_targets=False, mask=mask)
        run["targets"] = self._targets(imgs_dict)
        T = run["T"]
        self.last_run = run
        plist = [preds[m] for m in self.out_modalities]
        if torch.is_grad_enabled():
            total, losses = _LossFunction.apply(self, run, self._ddp_token, *plist)
            loss_dict = {m: losses[i] for i, m in enumerate(self.out_modalities)}
            if self.loss_aggr == "uncertainty":
                return total, loss_dict, _LazyList(self.loss_fn.log_vars), losses[T:2 * T]
            return total, loss_dict, None, None
        self._pack_preds(run, plist)
        self._stages(run, nat.STAGE_MASK | nat.STAGE_LOSS)
        losses = run["losses"]
        loss_dict = {m: losses[i] for i, m in enumerate(self.out_modalities)}
        if self.loss_aggr == "uncertainty":
            return losses[2 * T], loss_dict, _LazyList(self.loss_fn.log_vars), losses[T:2 * T]
        return losses[2 * T], loss_dict, None, None

    def upsample_mask(self, mask: torch.Tensor, scale: int) -> torch.Tensor:
        """``models/fcmae.py:233-240``."""
        assert len(mask.shape) == 2
        p = int(mask.shape[1] ** 0.5)
        return mask.reshape(-1, p, p).repeat_interleave(scale, dim=1).repeat_interleave(scale, dim=2)

    def unpatchify(self, x: torch.Tensor) -> torch.Tensor:
        """``models/fcmae.py:199-212``: ``[N, L, p*p*in_chans]`` -> ``[N, in_chans, H, W]``."""
        p = self.patch_size
        h = w = self.img_size // p
        x = x.reshape(x.shape[0], h, w, p, p, self.in_chans)
        return torch.einsum("nhwpqc->nchpwq", x).reshape(x.shape[0], self.in_chans, h * p, h * p)

    def encoder_features(self, run: Optional[dict] = None) -> torch.Tensor:
        run = run or self.last_run
        G = self.img_size // self.patch_size
        out = torch.empty(run["B"], self.dims[3], G, G, device=run["dev"])
        io = self._io(run)
        stream = torch.cuda.current_stream(run["dev"]).cuda_stream
        with torch.cuda.device(run["dev"]):
            nat.check(nat.lib.mpmae_encoder_features(run["plan"].handle, C.byref(io), C.c_void_p(out.data_ptr()),
                                                     C.c_void_p(stream)), "mpmae_encoder_features")
        return out

    def tap(self, name: str, run: Optional[dict] = None) -> torch.Tensor:
        """Named intermediate tensor of the last forward (parity tests)."""
        run = run or self.last_run
        off, rows, cols = run["plan"].tap(name)
        return self._workspace[off // 4: off // 4 + rows * cols].view(rows, cols)

    def input_flags(self) -> List[int]:
        """[n all-zero visible input pixels, ...]: non-zero means the sample violates the fast-path
        precondition (``to_sparse`` would have dropped that pixel, MinkowskiOps.py:308-317).  Host sync."""
        return self._flags.tolist()

    def forward(self, imgs_dict: Dict[str, torch.Tensor], labels=None, mask_ratio: float = 0.6):
        """``models/fcmae.py:414-456``."""
        if abs(mask_ratio - self.mask_ratio) > 1e-7:
            self.mask_ratio = mask_ratio
        run = self._prepare(imgs_dict)
        total, losses = _StepFunction.apply(self, run, self._ddp_token, *self._param_list)
        self.last_run = run
        T = run["T"]
        pred = self._pred_dict(run)
        loss_dict = {m: losses[i] for i, m in enumerate(self.out_modalities)}
        if self.loss_aggr == "uncertainty":
            log_vars = _LazyList(self.loss_fn.log_vars)
            normalized = losses[T:2 * T]
        else:
            log_vars, normalized = None, None
        return total, pred, run["mask"], loss_dict, log_vars, normalized


def _drop_token(module, state_dict, prefix, local_metadata):
    state_dict.pop(prefix + "_ddp_token", None)
    return state_dict


def _add_token(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
    state_dict.setdefault(prefix + "_ddp_token", torch.zeros(1))


class _LazyList(list):
    """``log_vars`` is a Python list in the reference (``custom_loss.py:30`` -> host sync every step).
    This list fills itself from the parameter on first access, so the sync only happens if it is read."""

    def __init__(self, param):
        super().__init__()
        self._param = param
        self._filled = False

    def _fill(self):
        if not self._filled:
            self._filled = True
            super().extend(self._param.detach().tolist())

    def __iter__(self):
        self._fill()
        return super().__iter__()

    def __len__(self):
        return self._param.numel()

    def __getitem__(self, i):
        self._fill()
        return super().__getitem__(i)

    def __repr__(self):
        self._fill()
        return super().__repr__()


# ---------------------------------------------------------------------- factories, models/fcmae.py:459-496
def convnextv2_atto(**kwargs):
    return FCMAE(depths=[2, 2, 6, 2], dims=[40, 80, 160, 320], **kwargs)


def convnextv2_femto(**kwargs):
    return FCMAE(depths=[2, 2, 6, 2], dims=[48, 96, 192, 384], **kwargs)


def convnextv2_pico(**kwargs):
    return FCMAE(depths=[2, 2, 6, 2], dims=[64, 128, 256, 512], **kwargs)


def convnextv2_nano(**kwargs):
    return FCMAE(depths=[2, 2, 8, 2], dims=[80, 160, 320, 640], **kwargs)


def convnextv2_tiny(**kwargs):
    return FCMAE(depths=[3, 3, 9, 3], dims=[96, 192, 384, 768], **kwargs)


def convnextv2_base(**kwargs):
    return FCMAE(depths=[3, 3, 27, 3], dims=[128, 256, 512, 1024], **kwargs)


def convnextv2_large(**kwargs):
    return FCMAE(depths=[3, 3, 27, 3], dims=[192, 384, 768, 1536], **kwargs)


def convnextv2_huge(**kwargs):
    return FCMAE(depths=[3, 3, 27, 3], dims=[352, 704, 1408, 2816], **kwargs)
