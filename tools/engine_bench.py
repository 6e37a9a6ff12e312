#!/usr/bin/env python
"""Loop-level throughput at cfg2 (bs 256, pinned host batches): mmearth_train_b200.engine.train_one_epoch against a loop
written the way the reference's engine_pretrain.train_one_epoch is (engine_pretrain.py:44-113): .to(device) on the compute
stream, loss.item() + one .item() per modality + normalized_loss.cpu() every step, torch GradScaler (host sync in step),
empty_cache() after every update.  Both loops drive the SAME native model; the difference is host syncs only.

    python tools/engine_bench.py [--iters 40] [--batch 256]
Prints one JSON line.  Not bench.py: this measures the loop around the step.
"""
import argparse
import json
import os
import sys
import time
from argparse import Namespace

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402


def reference_style_epoch(model, loader, optimizer, device, args, scaler):
    from mmearth_train_b200.optim import cosine_lr
    model.train()
    optimizer.zero_grad()
    n_iter = len(loader)
    total = 0.0
    for it, (_, samples) in enumerate(loader):
        lr = cosine_lr(it / n_iter, args.lr, args.min_lr, args.warmup_epochs, args.epochs)
        for g in optimizer.param_groups:
            g["lr"] = lr
        samples = {k: v.to(device, non_blocking=True) for k, v in samples.items()}
        loss, pred, mask, loss_dict_, log_vars, normalized = model(samples, mask_ratio=args.mask_ratio)
        _ = normalized.detach().cpu().numpy() if normalized is not None else None
        loss_value = loss.item()
        _ = {k: v.item() for k, v in loss_dict_.items()}
        total += loss_value
        scaler.scale(loss).backward()
        scaler.unscale_(optimizer)
        _ = torch.linalg.vector_norm(model.flat_grads)          # get_grad_norm_ (one flat norm instead of 184)
        scaler.step(optimizer)
        scaler.update()
        optimizer.zero_grad()
        torch.cuda.empty_cache()
    torch.cuda.synchronize()
    return total / n_iter


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=40)
    ap.add_argument("--batch", type=int, default=256)
    a = ap.parse_args()
    import mmearth_train_b200 as mp
    from mmearth_train_b200 import engine
    from mmearth_train_b200.optim import FlatAdamW, FlatGradScaler
    from mmearth_train_b200 import synthetic as fo
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    margs = fo.make_args(None, "uncertainty")
    args = Namespace(update_freq=1, lr=1.5e-4, min_lr=1e-6, warmup_epochs=1, epochs=10, mask_ratio=0.6, no_ffcv=True)
    host = []
    for i in range(4):
        d = fo.synthetic_batch(a.batch, 56, seed=1234 + i)
        host.append({k: v.pin_memory() for k, v in d.items()})
    loader = [(i, host[i % 4]) for i in range(a.iters)]
    out = {}
    for name in ("reference_style", "engine"):
        torch.manual_seed(0)
        model = mp.convnextv2_atto(mask_ratio=0.6, decoder_depth=1, decoder_embed_dim=512, norm_pix_loss=True, patch_size=8,
                                   img_size=56, args=margs, loss_fn=mp.UncertaintyWeightingStrategy(12)).to(dev)
        opt = FlatAdamW(model, lr=args.lr)
        warm = loader[:6]
        if name == "engine":
            run = lambda ld: engine.train_one_epoch(model, None, ld, opt, dev, 0, False, scaler, args=args, quiet=True)[0]["loss"]
            scaler = FlatGradScaler(dev)
        else:
            scaler = torch.amp.GradScaler("cuda")
            run = lambda ld: reference_style_epoch(model, ld, opt, dev, args, scaler)
        run(warm)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        loss = run(loader)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        out[name] = {"samples_per_s": a.batch * a.iters / dt, "ms_per_iter": dt / a.iters * 1e3, "mean_loss": loss}
        del model, opt
    out["speedup"] = out["engine"]["samples_per_s"] / out["reference_style"]["samples_per_s"]
    out["config"] = f"cfg2 atto 56/p8 all modalities, bs {a.batch}, {a.iters} iterations, pinned host batches, fwd+bwd+AdamW"
    print(json.dumps(out))


if __name__ == "__main__":
    main()
