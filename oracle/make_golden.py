"""Generates ``tests/golden/*.npz`` by running the UNMODIFIED reference (``/root/reference``).

TEST INFRASTRUCTURE ONLY; runs in the build container only (the reference does not exist on the
GPU box).  The reference's own ``models/fcmae.py:FCMAE(sparse=True)``, ``models/convnextv2_sparse.py``
and ``custom_loss.py`` execute unmodified on top of ``oracle/me_shim.py`` (CPU restatement of the six
MinkowskiEngine ops, itself pinned by the reference's depthwise known-answer vectors); forward
outputs, losses, the mask and sampled parameter gradients are written as small fixtures.

    python -m oracle.make_golden            # rewrites every fixture

Inputs are NOT stored: they are regenerated from seeds by ``oracle.fcmae_oracle.synthetic_batch`` /
``init_like_reference`` (CPU torch generators), and the fixture keeps checksums of them so a torch
RNG change is detected instead of silently comparing different problems.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import fcmae_oracle as fo  # noqa: E402
from oracle import ref_harness  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

PIX = ["sentinel2", "sentinel1", "aster", "dynamic_world", "canopy_height_eth", "esa_worldcover"]
IMG = ["era5", "lat", "lon", "biome", "eco_region", "month"]

# name -> config (BASELINE.json configs at fixture-sized batches)
CASES = {
    "atto_p8_all_unc": dict(model="convnextv2_atto", img_size=56, patch_size=8, out_modalities=None,
                            loss_aggr="uncertainty", B=2, nan_frac=0.05),
    "atto_p8_s2_unw": dict(model="convnextv2_atto", img_size=56, patch_size=8, out_modalities=["sentinel2"],
                           loss_aggr="unweighted", B=2, nan_frac=0.0),
    "atto_p16_all_unc": dict(model="convnextv2_atto", img_size=112, patch_size=16, out_modalities=None,
                             loss_aggr="uncertainty", B=1, nan_frac=0.05),
    "tiny_p8_all_unc": dict(model="convnextv2_tiny", img_size=56, patch_size=8, out_modalities=None,
                            loss_aggr="uncertainty", B=1, nan_frac=0.0),
    "atto_p8_pix_unw": dict(model="convnextv2_atto", img_size=56, patch_size=8, out_modalities=PIX,
                            loss_aggr="unweighted", B=1, nan_frac=0.05),
    "atto_p8_img_unw": dict(model="convnextv2_atto", img_size=56, patch_size=8, out_modalities=IMG,
                            loss_aggr="unweighted", B=2, nan_frac=0.05),
}
WEIGHT_SEED, DATA_SEED, NOISE_SEED = 3, 5, 11
GRAD_SAMPLES = 24


def sample_index(numel: int) -> np.ndarray:
    return np.unique(np.linspace(0, numel - 1, GRAD_SAMPLES).astype(np.int64))


def case_inputs(cfg):
    """(oracle model with seeded weights, batch, noise) -- shared by generator and tests."""
    orc = fo.build_oracle(model=cfg["model"], img_size=cfg["img_size"], patch_size=cfg["patch_size"],
                          out_modalities=cfg["out_modalities"], loss_aggr=cfg["loss_aggr"])
    fo.init_like_reference(orc, seed=WEIGHT_SEED)
    batch = fo.synthetic_batch(cfg["B"], cfg["img_size"], cfg["out_modalities"], seed=DATA_SEED,
                               nan_frac=cfg["nan_frac"])
    L = (cfg["img_size"] // cfg["patch_size"]) ** 2
    g = torch.Generator().manual_seed(NOISE_SEED)
    noise = torch.randn(cfg["B"], L, generator=g)
    return orc, batch, noise


def run_reference(cfg):
    orc, batch, noise = case_inputs(cfg)
    ref, args = ref_harness.build_reference_model(model=cfg["model"], img_size=cfg["img_size"],
                                                  patch_size=cfg["patch_size"], out_modalities=cfg["out_modalities"],
                                                  loss_aggr=cfg["loss_aggr"])
    ref.load_state_dict(orc.state_dict())
    ref.train()
    # the reference draws its noise with torch.randn(N, L) from the global generator (fcmae.py:220)
    real_randn = torch.randn
    torch.randn = lambda *a, **k: noise.clone()
    try:
        loss, pred, mask, loss_dict, log_vars, weighted = ref({k: v.clone() for k, v in batch.items()}, mask_ratio=0.6)
        # encoder features of the same forward: run the reference encoder again on the same mask
        with torch.no_grad():
            feats = ref.encoder(batch["sentinel2"].clone(), mask)
    finally:
        torch.randn = real_randn
    loss.backward()
    return orc, ref, batch, noise, dict(loss=loss, pred=pred, mask=mask, loss_dict=loss_dict, weighted=weighted,
                                        feats=feats)


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    for name, cfg in CASES.items():
        orc, ref, batch, noise, out = run_reference(cfg)
        arrays = {}
        arrays["loss"] = out["loss"].detach().numpy()
        arrays["mask"] = out["mask"].numpy().astype(np.uint8)
        arrays["encoder_features"] = out["feats"].numpy().astype(np.float32)
        mods = list(out["pred"].keys())
        for m in mods:
            arrays[f"pred.{m}"] = out["pred"][m].detach().numpy().astype(np.float32)
            arrays[f"loss.{m}"] = out["loss_dict"][m].detach().numpy()
        if out["weighted"] is not None:
            arrays["weighted"] = out["weighted"].detach().numpy()
        seen = set()
        gnames = []
        for pname, p in ref.named_parameters():
            if id(p) in seen or p.grad is None:
                continue
            seen.add(id(p))
            g = p.grad.reshape(-1)
            idx = sample_index(g.numel())
            arrays[f"grad.{pname}.sample"] = g[idx].numpy().astype(np.float32)
            arrays[f"grad.{pname}.norm"] = np.float64(g.double().norm().item())
            gnames.append(pname)
        meta = dict(case=name, cfg=cfg, seeds=dict(weight=WEIGHT_SEED, data=DATA_SEED, noise=NOISE_SEED),
                    modalities=mods, grad_params=gnames,
                    state_keys={k: list(v.shape) for k, v in ref.state_dict().items()},
                    input_checksum=dict(s2=float(batch["sentinel2"].double().sum()), noise=float(noise.double().sum()),
                                        w=float(sum(p.double().sum() for p in orc.parameters()))),
                    torch=torch.__version__, reference="vishalned/MMEarth-train@1f368b3 on oracle/me_shim.py")
        arrays["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
        path = os.path.join(GOLDEN_DIR, name + ".npz")
        np.savez_compressed(path, **arrays)
        print(f"{name}: loss {float(out['loss']):.6f}  {os.path.getsize(path) / 1e6:.2f} MB  ({len(gnames)} grads)")


if __name__ == "__main__":
    main()
