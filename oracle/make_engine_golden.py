"""Golden trajectory of the UNMODIFIED reference training loop (TEST INFRASTRUCTURE ONLY; runs in the build container).

Runs ``engine_pretrain.train_one_epoch`` (``/root/reference/engine_pretrain.py:21-126``) around the unmodified reference
``FCMAE(sparse=True)`` (on ``oracle/me_shim.py``), with the optimizer ``main_pretrain.py:312-320`` builds (AdamW, betas
(0.9, 0.95), timm's no-decay rule restated here because timm is not installed) and ``helpers.NativeScalerWithGradNormCount``
(GradScaler disabled on CPU), on seeded synthetic batches, and writes ``tests/golden/engine_<case>.json``: the learning
rate and loss of every iteration (as the loop logs them), the returned statistics, the last ``loss_dict`` and the noise of
every iteration (so a replay draws the same masks).  Cases: BASELINE.json configs[0] (atto, S2 -> S2, 56/p8, mask 0.6, bs 8,
unweighted) and a 12-modality uncertainty-weighted run with gradient accumulation (update_freq 2).

    python -m oracle.make_engine_golden
"""
from __future__ import annotations

import contextlib
import io
import json
import os
from argparse import Namespace

import torch

from . import fcmae_oracle as fo
from . import ref_harness

OUT_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CASES = {
    "cfg1": dict(model="convnextv2_atto", img_size=56, patch_size=8, out_modalities=["sentinel2"], loss_aggr="unweighted",
                 B=8, n_iter=6, update_freq=1, epoch=0, nan_frac=0.0),
    "all_unc_accum": dict(model="convnextv2_atto", img_size=56, patch_size=8, out_modalities=None, loss_aggr="uncertainty",
                          B=2, n_iter=6, update_freq=2, epoch=1, nan_frac=0.05),
}
LOOP_ARGS = dict(lr=3e-4, min_lr=1e-6, warmup_epochs=1, epochs=4, mask_ratio=0.6, weight_decay=0.05, no_ffcv=True)
WEIGHT_SEED, DATA_SEED, NOISE_SEED = 3, 700, 900


def case_inputs(cfg):
    """Seeded oracle model (the weights both sides start from), the batches and the per-iteration noise."""
    orc = fo.build_oracle(model=cfg["model"], img_size=cfg["img_size"], patch_size=cfg["patch_size"],
                          out_modalities=cfg["out_modalities"], loss_aggr=cfg["loss_aggr"])
    fo.init_like_reference(orc, seed=WEIGHT_SEED)
    batches = [fo.synthetic_batch(cfg["B"], cfg["img_size"], cfg["out_modalities"], seed=DATA_SEED + i, nan_frac=cfg["nan_frac"])
               for i in range(cfg["n_iter"])]
    L = (cfg["img_size"] // cfg["patch_size"]) ** 2
    g = torch.Generator().manual_seed(NOISE_SEED)
    noises = [torch.randn(cfg["B"], L, generator=g) for _ in range(cfg["n_iter"])]
    return orc, batches, noises


def loop_args(cfg) -> Namespace:
    return Namespace(update_freq=cfg["update_freq"], **LOOP_ARGS)


def param_groups_weight_decay(model, weight_decay):
    """timm.optim.optim_factory.param_groups_weight_decay (timm 0.9.7): no decay for ndim <= 1 or names ending in .bias."""
    decay, no_decay, seen = [], [], set()
    for name, p in model.named_parameters():
        if not p.requires_grad or id(p) in seen:
            continue
        seen.add(id(p))
        (no_decay if (p.ndim <= 1 or name.endswith(".bias")) else decay).append(p)
    return [{"params": no_decay, "weight_decay": 0.0}, {"params": decay, "weight_decay": weight_decay}]


class _Writer:
    """log_writer stand-in: records what engine_pretrain.py:104-112 sends to tensorboard."""

    def __init__(self):
        self.rows = []

    def update(self, head="scalar", step=None, **kwargs):
        self.rows.append(dict(head=head, step=step, **{k: float(v) for k, v in kwargs.items()}))


def run_reference_engine(cfg):
    helpers = ref_harness.import_toplevel("helpers")             # engine_pretrain does `import helpers`
    engine_pretrain = ref_harness.import_toplevel("engine_pretrain")
    orc, batches, noises = case_inputs(cfg)
    model, _ = ref_harness.build_reference_model(model=cfg["model"], img_size=cfg["img_size"], patch_size=cfg["patch_size"],
                                                 out_modalities=cfg["out_modalities"], loss_aggr=cfg["loss_aggr"])
    model.load_state_dict(orc.state_dict())
    args = loop_args(cfg)
    optimizer = torch.optim.AdamW(param_groups_weight_decay(model, args.weight_decay), lr=args.lr, betas=(0.9, 0.95))
    scaler = helpers.NativeScalerWithGradNormCount("cpu")
    writer = _Writer()
    loader = [(i, {k: v.clone() for k, v in b.items()}) for i, b in enumerate(batches)]
    queue = [n.clone() for n in noises]
    real_randn = torch.randn
    torch.randn = lambda *a, **k: queue.pop(0)                   # the reference draws its noise with torch.randn (fcmae.py:220)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            stats, loss_dict, log_vars, normalized = engine_pretrain.train_one_epoch(
                model, None, loader, optimizer, torch.device("cpu"), cfg["epoch"], False, scaler, log_writer=writer, args=args)
    finally:
        torch.randn = real_randn
    assert not queue
    return dict(
        cfg=cfg, loop_args=LOOP_ARGS,
        stats={k: float(v) for k, v in stats.items()},
        loss_dict={k: float(v) for k, v in loss_dict.items()},
        log_vars=None if log_vars is None else [float(v) for v in log_vars],
        normalized=None if normalized is None else [float(v) for v in normalized],
        logged=writer.rows,
        final_param_norm=float(torch.sqrt(sum((p.detach().double() ** 2).sum() for p in
                                              {id(p): p for p in model.parameters()}.values()))),
        input_checksum=dict(s2=float(sum(b["sentinel2"].double().sum() for b in batches)),
                            noise=float(sum(n.double().sum() for n in noises))),
    )


def main():
    os.makedirs(OUT_DIR, exist_ok=True)
    for name, cfg in CASES.items():
        out = run_reference_engine(cfg)
        path = os.path.join(OUT_DIR, f"engine_{name}.json")
        with open(path, "w") as f:
            json.dump(out, f, indent=1)
        print(name, out["stats"], "->", path)


if __name__ == "__main__":
    main()
