"""Golden outputs of the UNMODIFIED reference dense ConvNeXtV2 (TEST INFRASTRUCTURE ONLY; runs in the build container).

``models/convnextv2.py`` is imported as it is (timm stubbed by ``oracle.ref_harness``), loaded with the seeded weights of
``oracle.dense_oracle.seeded_state_dict`` and run on seeded inputs; logits and pooled features go to
``tests/golden/dense_<case>.npz``.

    python -m oracle.make_dense_golden
"""
from __future__ import annotations

import importlib
import json
import os

import numpy as np
import torch

from . import dense_oracle as do
from . import ref_harness

OUT_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
CASES = {"dense_atto_p8": dict(model="convnextv2_atto", img_size=56, patch_size=8, B=3, num_classes=10),
         "dense_atto_p16": dict(model="convnextv2_atto", img_size=112, patch_size=16, B=2, num_classes=19),
         "dense_tiny_p8": dict(model="convnextv2_tiny", img_size=56, patch_size=8, B=2, num_classes=10)}


def case_inputs(cfg):
    depths, dims = do.ZOO[cfg["model"]]
    sd = do.seeded_state_dict(do.state_dict_shapes(depths, dims, 12, cfg["patch_size"], cfg["num_classes"]), seed=5)
    x = torch.randn(cfg["B"], 12, cfg["img_size"], cfg["img_size"], generator=torch.Generator().manual_seed(77))
    return sd, x, depths, dims


def main():
    ref = ref_harness.load_reference()
    cn = importlib.import_module(f"{ref.pkg}.models.convnextv2")
    for name, cfg in CASES.items():
        sd, x, depths, dims = case_inputs(cfg)
        model = cn.__dict__[cfg["model"]](patch_size=cfg["patch_size"], img_size=cfg["img_size"], in_chans=12,
                                          num_classes=cfg["num_classes"])
        missing = model.load_state_dict(sd, strict=True)
        model.eval()
        with torch.no_grad():
            feats = model.forward_features(x.clone())
            logits = model(x.clone())
        meta = dict(cfg=cfg, checksum=dict(x=float(x.double().sum()), w=float(sum(v.double().sum() for v in sd.values()))),
                    keys=sorted(sd))
        np.savez_compressed(os.path.join(OUT_DIR, name + ".npz"), features=feats.numpy(), logits=logits.numpy(),
                            meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8))
        print(name, tuple(feats.shape), tuple(logits.shape), str(missing))


if __name__ == "__main__":
    main()
