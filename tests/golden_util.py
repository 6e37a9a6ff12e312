"""Shared helpers for the golden fixtures (tests/golden/*.npz, made by oracle/make_golden.py)."""
import json
import os

import numpy as np
import torch

from oracle import fcmae_oracle as fo
from oracle import make_golden as mg

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = list(mg.CASES.keys())


def load(case):
    z = np.load(os.path.join(GOLDEN_DIR, case + ".npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    return z, meta


def inputs(case):
    """Seeded oracle model, batch and noise of a fixture, with the RNG checksum verified."""
    z, meta = load(case)
    orc, batch, noise = mg.case_inputs(meta["cfg"])
    cs = meta["input_checksum"]
    assert abs(float(batch["sentinel2"].double().sum()) - cs["s2"]) < 1e-6 * max(1.0, abs(cs["s2"])), \
        "torch CPU generator changed: regenerate fixtures in the build container"
    assert abs(float(noise.double().sum()) - cs["noise"]) < 1e-9 * max(1.0, abs(cs["noise"])) + 1e-9
    return z, meta, orc, batch, noise


def rel_err(a, b):
    a = torch.as_tensor(np.asarray(a), dtype=torch.float64) if not torch.is_tensor(a) else a.double().cpu()
    b = torch.as_tensor(np.asarray(b), dtype=torch.float64) if not torch.is_tensor(b) else b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def max_rel(a, b):
    """max |a-b| / max |b| (the metric used for activations)."""
    a = torch.as_tensor(np.asarray(a), dtype=torch.float64) if not torch.is_tensor(a) else a.double().cpu()
    b = torch.as_tensor(np.asarray(b), dtype=torch.float64) if not torch.is_tensor(b) else b.double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def elementwise(a, b, rtol=1e-3, atol_frac=1e-3):
    """Element-wise criterion |a - b| <= rtol * |b| + atol with atol = atol_frac * rms(b) (the scale of the tensor: a pure
    relative bound is meaningless for the entries that happen to be near zero).  Returns the worst ratio
    |a - b| / (rtol * |b| + atol) -- the test passes when it is <= 1 -- and the largest absolute difference."""
    a = torch.as_tensor(np.asarray(a), dtype=torch.float64) if not torch.is_tensor(a) else a.double().cpu()
    b = torch.as_tensor(np.asarray(b), dtype=torch.float64) if not torch.is_tensor(b) else b.double().cpu()
    atol = atol_frac * float(b.pow(2).mean().sqrt()) + 1e-30
    d = (a - b).abs()
    return float((d / (rtol * b.abs() + atol)).max()), float(d.max())


def oracle_grads(orc, loss):
    orc.zero_grad(set_to_none=True)
    loss.backward()
    out, seen = {}, set()
    for n, p in orc.named_parameters():
        if id(p) in seen:
            continue
        seen.add(id(p))
        out[n] = None if p.grad is None else p.grad.detach().clone()
    return out
