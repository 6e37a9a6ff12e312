"""BASELINE.md section 5, leg (A): the UNMODIFIED reference on CPU (TEST / MEASUREMENT INFRASTRUCTURE; build container only).

``engine_pretrain.train_one_epoch`` (``/root/reference/engine_pretrain.py:21-126``) driving the reference's own dense
``FCMAE(sparse=False)`` at 112 / patch 16 -- the only geometry its CPU path supports -- for
``convnextv2_atto``, bs 8, S2 -> S2 and S2 -> all 12 modalities, AdamW as ``main_pretrain.py:312-320``, on the host cores of
THIS container.  ``/root/reference`` does not exist on the GPU box, so this leg cannot run next to the GPU numbers; its
result is committed as ``profiles/r2_reference_leg_A.json`` and quoted by ``bench.py``'s ``cpu_baseline`` note.

    python -m oracle.time_reference_leg_a
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import time
from argparse import Namespace

import torch

from . import fcmae_oracle as fo
from . import make_engine_golden as meg
from . import ref_harness

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r2_reference_leg_A.json")


def run(out_modalities, loss_aggr, n_iter=12, warm=2, B=8):
    helpers = ref_harness.import_toplevel("helpers")
    engine_pretrain = ref_harness.import_toplevel("engine_pretrain")
    ref = ref_harness.load_reference()
    args_m = ref_harness.make_args(out_modalities, loss_aggr)
    lf = ref.custom_loss.UncertaintyWeightingStrategy(len(args_m.out_modalities)) if loss_aggr == "uncertainty" else None
    torch.manual_seed(0)
    model = ref.fcmae.convnextv2_atto(mask_ratio=0.6, decoder_depth=1, decoder_embed_dim=512, norm_pix_loss=True, patch_size=16,
                                      img_size=112, args=args_m, loss_fn=lf, sparse=False)
    args = Namespace(update_freq=1, lr=1.5e-4, min_lr=1e-6, warmup_epochs=1, epochs=4, mask_ratio=0.6, weight_decay=0.05, no_ffcv=True)
    optimizer = torch.optim.AdamW(meg.param_groups_weight_decay(model, args.weight_decay), lr=args.lr, betas=(0.9, 0.95))
    scaler = helpers.NativeScalerWithGradNormCount("cpu")
    batches = [fo.synthetic_batch(B, 112, out_modalities, seed=1234 + i) for i in range(4)]

    def epoch(n):
        loader = [(i, {k: v.clone() for k, v in batches[i % 4].items()}) for i in range(n)]
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            stats = engine_pretrain.train_one_epoch(model, None, loader, optimizer, torch.device("cpu"), 0, False, scaler,
                                                    log_writer=None, args=args)[0]
        return time.perf_counter() - t0, stats

    epoch(warm)
    dt, stats = epoch(n_iter)
    return {"samples_per_s": B * n_iter / dt, "ms_per_iteration": dt / n_iter * 1e3, "iterations": n_iter, "batch": B,
            "loss": float(stats["loss"])}


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    out = {"what": "unmodified reference engine_pretrain.train_one_epoch + FCMAE(sparse=False), convnextv2_atto, 112/p16, bs 8, CPU",
           "cores": torch.get_num_threads(), "torch": torch.__version__,
           "S2_to_S2": run(["sentinel2"], "unweighted"), "S2_to_all": run(None, "uncertainty")}
    with open(OUT, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
