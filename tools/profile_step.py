"""Runs a few MP-MAE steps of a bench config for ncu (no timing printed: numbers under a profiler are not bench values).

    ncu --metrics gpu__time_duration.sum --clock-control none -s <skip> -c <n> --csv --log-file gpurun_out/launches.csv \
        python tools/profile_step.py --config cfg2 --steps 2
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import mmearth_train_b200 as mp  # noqa: E402
from bench import CONFIGS  # noqa: E402
from mmearth_train_b200.optim import FlatAdamW  # noqa: E402
from mmearth_train_b200 import synthetic as fo  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="cfg2")
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--batch", type=int, default=None)
ap.add_argument("--backend", type=int, default=None)
a = ap.parse_args()
cfg = CONFIGS[a.config]
B = a.batch or cfg["batch"]
args = fo.make_args(cfg["out_modalities"], cfg["loss_aggr"])
lf = mp.UncertaintyWeightingStrategy(len(args.out_modalities)) if cfg["loss_aggr"] == "uncertainty" else None
torch.manual_seed(0)
model = getattr(mp, cfg["model"])(mask_ratio=0.6, decoder_depth=1, decoder_embed_dim=512, norm_pix_loss=True,
                                  patch_size=cfg["patch_size"], img_size=cfg["img_size"], args=args, loss_fn=lf,
                                  gemm_backend=a.backend).cuda()
opt = FlatAdamW(model)
batch = {k: v.cuda() for k, v in fo.synthetic_batch(B, cfg["img_size"], cfg["out_modalities"], seed=1).items()}
for i in range(a.steps):
    loss = model(batch, mask_ratio=0.6)[0]
    loss.backward()
    opt.step()
    opt.zero_grad(set_to_none=True)
torch.cuda.synchronize()
plan = model.last_run["plan"]
print("launches per step:", plan.launches(False), "+", plan.launches(True), "+ 1")
