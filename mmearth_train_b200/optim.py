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


class FlatAdamW:
    """AdamW over ``model.flat_params`` / ``model.flat_grads``; state is two flat fp32 buffers."""

    def __init__(self, model, lr: float = 1.5e-4, betas=(0.9, 0.95), eps: float = 1e-8, weight_decay: float = 0.05):
        self.model = model
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        flat = model.flat_params
        if not flat.is_cuda:
            raise RuntimeError("FlatAdamW runs on the CUDA flat buffers; move the model to the GPU first")
        self.exp_avg = torch.zeros_like(flat)
        self.exp_avg_sq = torch.zeros_like(flat)
        self.decay = model.decay_mask()
        self.t = 0

    def zero_grad(self, set_to_none: bool = True) -> None:
        self.model.zero_grad(set_to_none=set_to_none)

    def step(self, grad_scale_inv: float = 1.0) -> None:
        g = self.model.flat_grads
        if g is None:
            raise RuntimeError("no gradients: call loss.backward() first")
        self.t += 1
        p = self.model.flat_params
        stream = torch.cuda.current_stream(p.device).cuda_stream
        with torch.cuda.device(p.device):
            nat.check(nat.lib.mpmae_adamw_step(
                C.c_void_p(p.data_ptr()), C.c_void_p(g.data_ptr()), C.c_void_p(self.exp_avg.data_ptr()),
                C.c_void_p(self.exp_avg_sq.data_ptr()), C.c_void_p(self.decay.data_ptr()), p.numel(), self.lr,
                self.betas[0], self.betas[1], self.eps, self.weight_decay, self.t, grad_scale_inv,
                C.c_void_p(stream)), "mpmae_adamw_step")

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
