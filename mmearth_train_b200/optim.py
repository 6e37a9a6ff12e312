"""Optimizer step on the flat buffers: fused AdamW + the reference's
per-iteration half-cosine learning-rate schedule.

Mirrors ``main_pretrain.py:312-320`` (``torch.optim.AdamW(param_groups, lr, betas=(0.9, 0.95))`` with timm's
``param_groups_weight_decay``: weight decay 0.05, none for ``ndim <= 1`` or ``*.bias``) and
``helpers.adjust_learning_rate`` (``helpers.py:647-665``).  One kernel launch per step instead of the
reference's per-tensor loop.
"""
from __future__ import annotations

import ctypes as C
import math

import torch

from . import _native as nat


def cosine_lr(epoch: float, lr: float, min_lr: float, warmup_epochs: float, epochs: float) -> float:
    """``helpers.adjust_learning_rate`` (``helpers.py:647-665``); ``epoch`` is fractional (per iteration)."""
    if epoch < warmup_epochs:
        return lr * epoch / warmup_epochs
    return min_lr + (lr - min_lr) * 0.5 * (1.0 + math.cos(math.pi * (epoch - warmup_epochs) / (epochs - warmup_epochs)))


class FlatGradScaler:
    """``helpers.NativeScalerWithGradNormCount`` (``helpers.py:470-506``) for the flat buffers.

    Same call ``scaler(loss, optimizer, clip_grad=None, parameters=None, create_graph=False, update_grad=True)`` and the
    same dynamic loss scaling as ``torch.cuda.amp.GradScaler`` (initial scale 2**16, x0.5 and a skipped step on a
    non-finite gradient, x2 after 2000 clean steps), which the reference keeps enabled on CUDA even though the step runs
    in fp32 (``main_pretrain.py`` passes ``device`` to the scaler).  Differences: the scale, the non-finite flag and the
    growth counter stay on the device -- no ``.item()`` per step (``GradScaler.step`` syncs) -- and un-scaling is folded
    into the AdamW kernel instead of a pass over every gradient.  Returns the un-scaled gradient norm as a 0-d device
    tensor (``helpers.get_grad_norm_``), or ``None`` when ``update_grad`` is false.
    """
    state_dict_key = "amp_scaler"

    def __init__(self, device="cuda", init_scale: float = 65536.0, growth_factor: float = 2.0, backoff_factor: float = 0.5,
                 growth_interval: int = 2000, enabled: bool = True):
        self.device = torch.device(device)
        self.enabled = enabled and self.device.type != "cpu"
        self.growth_factor, self.backoff_factor, self.growth_interval = growth_factor, backoff_factor, growth_interval
        self._scale = torch.full((), init_scale if self.enabled else 1.0, device=self.device)
        self._growth = torch.zeros((), dtype=torch.int32, device=self.device)
        self.last_found_inf = None

    def __call__(self, loss, optimizer, clip_grad=None, parameters=None, create_graph=False, update_grad=True):
        if create_graph:
            raise NotImplementedError("the native backward is hand-derived: no double backward")
        (loss * self._scale).backward()
        if not update_grad:
            return None
        g = optimizer.model.flat_grads
        inv = 1.0 / self._scale
        norm = torch.linalg.vector_norm(g) * inv                 # norm of the un-scaled gradient
        found_inf = (~torch.isfinite(norm)).to(torch.float32)
        factor = inv
        if clip_grad is not None:                                # torch.nn.utils.clip_grad_norm_
            factor = inv * torch.clamp(clip_grad / (norm + 1e-6), max=1.0)
        optimizer.step_dev(factor, found_inf)
        self.last_found_inf = found_inf
        if self.enabled:                                         # GradScaler.update (_amp_update_scale_)
            bad = found_inf != 0
            grown = self._growth + 1
            grow_now = (~bad) & (grown == self.growth_interval)
            self._scale = torch.where(bad, self._scale * self.backoff_factor,
                                      torch.where(grow_now, self._scale * self.growth_factor, self._scale))
            self._growth = torch.where(bad | grow_now, torch.zeros_like(grown), grown)
        return norm

    def get_scale(self) -> float:
        return float(self._scale)

    def state_dict(self):
        return {"scale": float(self._scale), "growth_factor": self.growth_factor, "backoff_factor": self.backoff_factor,
                "growth_interval": self.growth_interval, "_growth_tracker": int(self._growth)}

    def load_state_dict(self, sd):
        if not sd:                  # a disabled torch GradScaler (CPU run) saves an empty dict
            return
        self._scale = torch.full((), float(sd["scale"]), device=self.device)
        self._growth = torch.full((), int(sd["_growth_tracker"]), dtype=torch.int32, device=self.device)
        self.growth_factor, self.backoff_factor = sd["growth_factor"], sd["backoff_factor"]
        self.growth_interval = sd["growth_interval"]


class FlatAdamW:
    """AdamW over ``model.flat_params`` / ``model.flat_grads``; state is two flat fp32 buffers."""

    def __init__(self, model, lr: float = 1.5e-4, betas=(0.9, 0.95), eps: float = 1e-8, weight_decay: float = 0.05):
        self.model = model
        self.betas, self.eps, self.weight_decay = betas, eps, weight_decay
        flat = model.flat_params
        if not flat.is_cuda:
            raise RuntimeError("FlatAdamW runs on the CUDA flat buffers; move the model to the GPU first")
        self.exp_avg = torch.zeros_like(flat)
        self.exp_avg_sq = torch.zeros_like(flat)
        self.decay = model.decay_mask()
        self._t = 0
        self._dev_state = None        # [grad factor, found_inf, step number] on the device (step_dev / FlatGradScaler)
        # what helpers.adjust_learning_rate (helpers.py:647-665), engine_pretrain.py:100 and torch's GradScaler read
        self.param_groups = [{"params": list(model.parameters()), "lr": lr, "betas": betas, "eps": eps,
                              "weight_decay": weight_decay}]
        self.defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)

    # the learning rate lives in param_groups[0] so that the reference's scheduler drives it unchanged
    @property
    def lr(self) -> float:
        g = self.param_groups[0]
        return g["lr"]

    @lr.setter
    def lr(self, value: float) -> None:
        self.param_groups[0]["lr"] = value

    @property
    def t(self) -> int:
        """Completed optimizer steps.  After ``step_dev`` the count lives on the device (skipped steps do not count):
        reading it then is a host sync."""
        if self._dev_state is not None:
            self._t = int(self._dev_state[2].item()) - 1
        return self._t

    @t.setter
    def t(self, value: int) -> None:
        self._t = int(value)
        if self._dev_state is not None:
            self._dev_state[2] = float(value + 1)

    def zero_grad(self, set_to_none: bool = True) -> None:
        self.model.zero_grad(set_to_none=set_to_none)

    def step(self, grad_scale_inv: float = 1.0) -> None:
        g = self.model.flat_grads
        if g is None:
            raise RuntimeError("no gradients: call loss.backward() first")
        step = self.t + 1
        self.t = step
        p = self.model.flat_params
        stream = torch.cuda.current_stream(p.device).cuda_stream
        with torch.cuda.device(p.device):
            nat.check(nat.lib.mpmae_adamw_step(
                C.c_void_p(p.data_ptr()), C.c_void_p(g.data_ptr()), C.c_void_p(self.exp_avg.data_ptr()),
                C.c_void_p(self.exp_avg_sq.data_ptr()), C.c_void_p(self.decay.data_ptr()), p.numel(), self.lr,
                self.betas[0], self.betas[1], self.eps, self.weight_decay, step, grad_scale_inv,
                C.c_void_p(stream)), "mpmae_adamw_step")

    def _ensure_dev_state(self) -> torch.Tensor:
        if self._dev_state is None:
            self._dev_state = torch.tensor([1.0, 0.0, float(self._t + 1), float(self.lr)], device=self.model.flat_params.device)
        return self._dev_state

    def step_dev(self, grad_factor=None, found_inf=None, lr_on_device: bool = False) -> None:
        """One step whose gradient factor (1 / loss scale x clipping coefficient) and skip flag are 0-d DEVICE tensors:
        nothing is read back, a step with ``found_inf != 0`` changes nothing and is not counted (what
        ``GradScaler.step`` does after a host sync, ``helpers.py:498``).  ``None`` leaves the stored factor / flag as they are.
        ``lr_on_device``: the learning rate is read from ``dev_state[3]`` instead of ``param_groups`` (CUDA-graph replay)."""
        g = self.model.flat_grads
        if g is None:
            raise RuntimeError("no gradients: call loss.backward() first")
        p = self.model.flat_params
        st = self._ensure_dev_state()
        if grad_factor is not None:
            st[0] = grad_factor
        if found_inf is not None:
            st[1] = found_inf
        stream = torch.cuda.current_stream(p.device).cuda_stream
        with torch.cuda.device(p.device):
            nat.check(nat.lib.mpmae_adamw_step_dev(
                C.c_void_p(p.data_ptr()), C.c_void_p(g.data_ptr()), C.c_void_p(self.exp_avg.data_ptr()),
                C.c_void_p(self.exp_avg_sq.data_ptr()), C.c_void_p(self.decay.data_ptr()), p.numel(),
                -1.0 if lr_on_device else self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay,
                C.c_void_p(st.data_ptr()), C.c_void_p(stream)), "mpmae_adamw_step_dev")
        st[2] += (st[1] == 0).to(st.dtype)

    def grad_norm(self, clip: float = None) -> torch.Tensor:
        """2-norm of the whole gradient as a 0-d device tensor (no host sync): ``helpers.get_grad_norm_``
        (``helpers.py:509-526``) over ONE flat buffer instead of a stack of per-parameter norms.  With ``clip`` the gradients
        are rescaled in place like ``torch.nn.utils.clip_grad_norm_`` (``helpers.py:489-497``)."""
        g = self.model.flat_grads
        if g is None:
            raise RuntimeError("no gradients: call loss.backward() first")
        n = torch.linalg.vector_norm(g)
        if clip is not None:
            g.mul_(torch.clamp(clip / (n + 1e-6), max=1.0))
        return n

    # ------------------------------------------------------------------ checkpoint interchange
    def _reference_groups(self):
        """Parameters as the reference's optimizer enumerates them: ``timm.optim.param_groups_weight_decay`` over
        ``model.named_parameters()`` gives ``[no_decay, decay]`` (``main_pretrain.py:312-320``), and
        ``torch.optim.Optimizer.state_dict`` numbers the parameters through the groups in that order.  Returns two lists of
        ``(name, offset, numel, shape)`` in the REFERENCE's registration order (``reference_param_order``)."""
        model = self.model
        by_name = {name: (off, numel, shape) for (name, _s, off, _d), (_o, numel, shape)
                   in zip(model._layout, model._param_slices)}
        decay_of = {name: d for name, _s, _o, d in model._layout}
        public = {}                                    # native layout name -> the module's (reference) parameter name
        for (name, _s, _o, _d), p in zip(model._layout, model._param_list):
            public[id(p)] = name
        names = []
        for n, p in model.named_parameters():
            if n == "_ddp_token":
                continue
            names.append((n, public[id(p)]))
        order = reference_param_order([n for n, _ in names], model.out_modalities)
        lookup = dict(names)
        groups = ([], [])
        for n in order:
            native = lookup[n]
            off, numel, shape = by_name[native]
            groups[1 if decay_of[native] else 0].append((n, off, numel, tuple(shape)))
        return groups

    def state_dict(self):
        """The layout ``torch.optim.AdamW.state_dict()`` has for the reference's two parameter groups (``state`` /
        ``param_groups``), so that a checkpoint written here resumes in the reference and the other way round
        (``helpers.py:541-610``)."""
        no_decay, decay = self._reference_groups()
        t = self.t
        state, idx, groups = {}, 0, []
        for members, wd in ((no_decay, 0.0), (decay, self.weight_decay)):
            ids = []
            for _n, off, numel, shape in members:
                state[idx] = {"step": torch.tensor(float(t)),
                              "exp_avg": self.exp_avg[off:off + numel].view(shape).clone(),
                              "exp_avg_sq": self.exp_avg_sq[off:off + numel].view(shape).clone()}
                ids.append(idx)
                idx += 1
            groups.append({"lr": self.lr, "betas": tuple(self.betas), "eps": self.eps, "weight_decay": wd, "amsgrad": False,
                           "maximize": False, "foreach": None, "capturable": False, "differentiable": False, "fused": None,
                           "decoupled_weight_decay": True, "params": ids})
        return {"state": state, "param_groups": groups}

    def load_state_dict(self, sd):
        if "exp_avg" in sd and "state" not in sd:               # round-1 private format (flat buffers)
            self.exp_avg.copy_(sd["exp_avg"])
            self.exp_avg_sq.copy_(sd["exp_avg_sq"])
            self.t, self.lr = sd["t"], sd["lr"]
            return
        if "state" not in sd or "param_groups" not in sd:
            raise ValueError("optimizer state is neither the torch.optim.AdamW layout (state / param_groups) nor the "
                             "flat-buffer layout (exp_avg / exp_avg_sq / t / lr)")
        no_decay, decay = self._reference_groups()
        saved = sd["param_groups"]
        if len(saved) != 2 or len(saved[0]["params"]) != len(no_decay) or len(saved[1]["params"]) != len(decay):
            raise ValueError(f"optimizer state has groups of {[len(g['params']) for g in saved]} parameters; this model has "
                             f"[{len(no_decay)}, {len(decay)}] (no-decay, decay: timm's rule, main_pretrain.py:312-320)")
        self.exp_avg.zero_()
        self.exp_avg_sq.zero_()
        step = None
        for members, g in ((no_decay, saved[0]), (decay, saved[1])):
            for (name, off, numel, shape), idx in zip(members, g["params"]):
                st = sd["state"].get(idx)
                if st is None:                                   # a parameter that never received a gradient
                    continue
                if tuple(st["exp_avg"].shape) != shape:
                    raise ValueError(f"optimizer state {idx} has shape {tuple(st['exp_avg'].shape)}, parameter {name} "
                                     f"has {shape}: the checkpoint belongs to another model")
                self.exp_avg[off:off + numel].copy_(st["exp_avg"].reshape(-1))
                self.exp_avg_sq[off:off + numel].copy_(st["exp_avg_sq"].reshape(-1))
                step = int(float(st["step"])) if step is None else step
        self.t = step or 0
        self.lr = saved[0]["lr"]


_SPARSE_BLOCK_LEAVES = ("dwconv.kernel", "dwconv.bias", "norm.ln.weight", "norm.ln.bias", "pwconv1.linear.weight",
                        "pwconv1.linear.bias", "pwconv2.linear.weight", "pwconv2.linear.bias", "grn.gamma", "grn.beta")
_DENSE_BLOCK_LEAVES = ("dwconv.weight", "dwconv.bias", "norm.weight", "norm.bias", "pwconv1.weight", "pwconv1.bias",
                       "grn.gamma", "grn.beta", "pwconv2.weight", "pwconv2.bias")


def reference_param_order(names, out_modalities):
    """``names`` sorted the way the reference module's ``named_parameters()`` yields them: root parameters first
    (``mask_token``), then the children in the order ``FCMAE.__init__`` / ``SparseConvNeXtV2.__init__`` register them
    (``models/fcmae.py:93-155``, ``models/convnextv2_sparse.py:95-160``): ``loss_fn``, encoder (downsample layers, initial
    conv, stem, stages), ``proj``, ``decoder_dict`` (the shared block appears once, under the first modality), ``pred_dict`` in
    modality order, ``layer_norm_tmp``.  Pinned against the unmodified reference in ``tests/test_checkpoint.py``."""
    top = {"mask_token": 0, "loss_fn": 1, "encoder": 2, "proj": 3, "decoder_dict": 4, "pred_dict": 5, "layer_norm_tmp": 6}
    enc = {"downsample_layers": 0, "initial_conv": 1, "stem": 2, "stages": 3}
    mods = {m: i for i, m in enumerate(out_modalities)}
    wb = {"weight": 0, "kernel": 0, "bias": 1}

    def key(n):
        p = n.split(".")
        if p[0] == "encoder":
            if p[1] == "stages":
                return (2, 3, int(p[2]), int(p[3]), _SPARSE_BLOCK_LEAVES.index(".".join(p[4:])))
            return (2, enc[p[1]], int(p[2]), int(p[3]) if p[3].isdigit() else 0, wb[p[-1]], 0)
        if p[0] == "decoder_dict":
            return (4, mods[p[1]], int(p[2]), _DENSE_BLOCK_LEAVES.index(".".join(p[3:])), 0)
        if p[0] == "pred_dict":
            return (5, mods[p[1]], wb[p[-1]], 0, 0)
        return (top[p[0]], wb.get(p[-1], 0), 0, 0, 0)

    return sorted(names, key=key)
