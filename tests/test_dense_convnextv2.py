"""The reference's dense ConvNeXtV2 (models/convnextv2.py:59-207; finetuning / linear probing, hubconf.py:77-93): the CPU
restatement against outputs of the UNMODIFIED reference module (tests/golden/dense_*.npz, oracle/make_dense_golden.py), the
native module's state-dict surface, and (gpu) the native inference forward against both."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import dense_oracle as do
from oracle import make_dense_golden as mdg

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    sd, x, depths, dims = mdg.case_inputs(meta["cfg"])
    assert abs(float(x.double().sum()) - meta["checksum"]["x"]) < 1e-6 * max(1.0, abs(meta["checksum"]["x"])), "torch CPU generator changed"
    assert abs(float(sum(v.double().sum() for v in sd.values())) - meta["checksum"]["w"]) < 1e-6 * abs(meta["checksum"]["w"])
    return z, meta["cfg"], sd, x, depths, dims


@pytest.mark.parametrize("name", list(mdg.CASES))
def test_dense_oracle_matches_unmodified_reference(name):
    z, cfg, sd, x, depths, dims = _load(name)
    with torch.no_grad():
        feats = do.forward_features(sd, x, depths, cfg["patch_size"])
        logits = do.forward(sd, x, depths, cfg["patch_size"])
    assert np.allclose(feats.numpy(), z["features"], rtol=1e-5, atol=1e-5)
    assert np.allclose(logits.numpy(), z["logits"], rtol=1e-5, atol=1e-5)


def test_native_module_has_the_reference_state_dict_and_loads_a_remapped_pretraining_checkpoint(native_lib):
    import mmearth_train_b200 as mp
    from mmearth_train_b200 import convnextv2 as cn
    from mmearth_train_b200.checkpoint import to_dense_state_dict
    from oracle import fcmae_oracle as fo
    for (name, ps, S, nc) in (("convnextv2_atto", 8, 56, 10), ("convnextv2_tiny", 16, 112, 19)):
        depths, dims = do.ZOO[name]
        m = getattr(cn, name)(patch_size=ps, img_size=S, num_classes=nc)
        assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == do.state_dict_shapes(depths, dims, 12, ps, nc)
    # a pretraining checkpoint of the native FCMAE, filtered like hubconf.py:31-36 and remapped like helpers.remap_checkpoint_keys
    pre = mp.convnextv2_atto(mask_ratio=0.6, decoder_depth=1, decoder_embed_dim=512, norm_pix_loss=True, patch_size=8, img_size=56,
                             args=fo.make_args(None, "uncertainty"), loss_fn=mp.UncertaintyWeightingStrategy(12))
    from mmearth_train_b200.checkpoint import load_pretrained_encoder
    dense = cn.convnextv2_atto(patch_size=8, img_size=56, num_classes=10)
    missing = load_pretrained_encoder(dense, {"model": pre.state_dict()}, linear_probe=True)
    assert sorted(missing) == ["head.bias", "head.weight", "norm.bias", "norm.weight"]
    assert to_dense_state_dict({"encoder.stages.0.0.norm.ln.weight": torch.ones(40)}).keys() == {"stages.0.0.norm.weight"}
    k = pre.state_dict()["encoder.stages.0.0.dwconv.kernel"]                       # ME [49, C] -> torch [C, 1, 7, 7]
    assert torch.equal(dense.stages[0][0].dwconv.weight[:, 0, 2, 5], k[2 + 7 * 5])
    assert dense.stages[0][0].grn.gamma.shape == (1, 1, 1, 160)
    with pytest.raises(RuntimeError):
        dense(torch.zeros(1, 12, 56, 56))                                           # no CPU path


@pytest.mark.gpu
@pytest.mark.parametrize("backend", [3, 0])
@pytest.mark.parametrize("name", list(mdg.CASES))
def test_native_dense_forward_matches_reference(name, backend):
    from mmearth_train_b200 import convnextv2 as cn
    from tests import golden_util as gu
    z, cfg, sd, x, depths, dims = _load(name)
    m = getattr(cn, cfg["model"])(patch_size=cfg["patch_size"], img_size=cfg["img_size"], num_classes=cfg["num_classes"],
                                  gemm_backend=backend)
    m.load_state_dict(sd)
    m = m.cuda()
    feats = m.forward_features(x.cuda())
    logits = m(x.cuda())
    for got, key in ((feats, "features"), (logits, "logits")):
        r, d = gu.elementwise(got, z[key])
        assert r <= 1.0, (key, r, d)
    with torch.no_grad():
        fmap = do.forward_features(sd, x, depths, cfg["patch_size"], pooled=False)
    r, d = gu.elementwise(m.feature_map(x.cuda()), fmap)
    assert r <= 1.0, ("feature_map", r, d)
    # parameters changed in place are picked up (the folded weights are rebuilt)
    with torch.no_grad():
        m.stages[0][0].norm.weight.mul_(1.5)
    sd2 = {k: v.clone() for k, v in sd.items()}
    sd2["stages.0.0.norm.weight"] = sd2["stages.0.0.norm.weight"] * 1.5
    with torch.no_grad():
        want = do.forward(sd2, x, depths, cfg["patch_size"])
    r, d = gu.elementwise(m(x.cuda()), want)
    assert r <= 1.0, ("after update", r, d)
