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

    def step_dev(self, grad_factor: torch.Tensor, found_inf: torch.Tensor) -> None:
        """One step whose gradient factor (1 / loss scale x clipping coefficient) and skip flag are 0-d DEVICE tensors:
        nothing is read back, a step with ``found_inf != 0`` changes nothing and is not counted (what
        ``GradScaler.step`` does after a host sync, ``helpers.py:498``)."""
        g = self.model.flat_grads
        if g is None:
            raise RuntimeError("no gradients: call loss.backward() first")
        p = self.model.flat_params
        if self._dev_state is None:
            self._dev_state = torch.tensor([1.0, 0.0, float(self._t + 1)], device=p.device)
        st = self._dev_state
        st[0] = grad_factor
        st[1] = found_inf
        stream = torch.cuda.current_stream(p.device).cuda_stream
        with torch.cuda.device(p.device):
            nat.check(nat.lib.mpmae_adamw_step_dev(
                C.c_void_p(p.data_ptr()), C.c_void_p(g.data_ptr()), C.c_void_p(self.exp_avg.data_ptr()),
                C.c_void_p(self.exp_avg_sq.data_ptr()), C.c_void_p(self.decay.data_ptr()), p.numel(), self.lr,
                self.betas[0], self.betas[1], self.eps, self.weight_decay, C.c_void_p(st.data_ptr()),
                C.c_void_p(stream)), "mpmae_adamw_step_dev")
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

    def state_dict(self):
        return {"exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq, "t": self.t, "lr": self.lr}

    def load_state_dict(self, sd):
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        self.t, self.lr = sd["t"], sd["lr"]
