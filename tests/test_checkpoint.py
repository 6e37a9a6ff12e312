"""Sparse -> dense checkpoint conversion: against an independent index formula, and against the
reference's own ``helpers.remap_checkpoint_keys`` when the reference tree is present (it is not on the GPU box)."""
import os

import pytest
import torch


def _fake_ckpt():
    g = torch.Generator().manual_seed(0)
    r = lambda *s: torch.randn(*s, generator=g)
    return {
        "encoder.initial_conv.0.kernel": r(9, 12, 40), "encoder.initial_conv.0.bias": r(1, 40),
        "encoder.initial_conv.1.ln.weight": r(40), "encoder.initial_conv.1.ln.bias": r(40),
        "encoder.stem.0.kernel": r(4, 40), "encoder.stem.0.bias": r(1, 40),
        "encoder.downsample_layers.0.0.ln.weight": r(40), "encoder.downsample_layers.0.1.kernel": r(4, 40, 80),
        "encoder.downsample_layers.0.1.bias": r(1, 80),
        "encoder.stages.0.0.dwconv.kernel": r(49, 40), "encoder.stages.0.0.dwconv.bias": r(1, 40),
        "encoder.stages.0.0.norm.ln.weight": r(40), "encoder.stages.0.0.pwconv1.linear.weight": r(160, 40),
        "encoder.stages.0.0.pwconv1.linear.bias": r(160), "encoder.stages.0.0.grn.gamma": r(1, 160),
        "encoder.stages.0.0.grn.beta": r(1, 160), "encoder.stages.0.0.pwconv2.linear.weight": r(40, 160),
        "proj.weight": r(512, 320, 1, 1), "mask_token": r(1, 512, 1, 1),
    }


def test_dense_layout_formula(native_lib):
    from mmearth_train_b200.checkpoint import to_dense_state_dict
    ck = _fake_ckpt()
    d = to_dense_state_dict(ck)
    w = d["initial_conv.0.weight"]
    assert w.shape == (40, 12, 3, 3)
    k = ck["encoder.initial_conv.0.kernel"]
    for kh in range(3):
        for kw in range(3):
            assert torch.equal(w[:, :, kh, kw], k[kh + 3 * kw].t())           # dense[o, i, kh, kw] = kernel[kh + ks*kw, i, o]
    dw = d["stages.0.0.dwconv.weight"]
    assert dw.shape == (40, 1, 7, 7)
    kd = ck["encoder.stages.0.0.dwconv.kernel"]
    assert torch.equal(dw[:, 0, 2, 5], kd[2 + 7 * 5]) and torch.equal(dw[:, 0, 6, 0], kd[6])
    assert d["stages.0.0.dwconv.bias"].shape == (40,) and d["stages.0.0.grn.gamma"].shape == (1, 1, 1, 160)
    assert "stages.0.0.norm.weight" in d and "stages.0.0.pwconv1.weight" in d and "downsample_layers.0.1.weight" in d
    assert d["proj.weight"].shape == (512, 320, 1, 1)


@pytest.mark.skipif(not os.path.isfile("/root/reference/helpers.py"), reason="reference tree not present")
def test_matches_reference_remap(native_lib):
    from mmearth_train_b200.checkpoint import to_dense_state_dict
    src = open("/root/reference/helpers.py").read()
    start = src.index("def remap_checkpoint_keys")
    end = src.index("\ndef ", start + 10)
    ns = {"OrderedDict": __import__("collections").OrderedDict, "math": __import__("math")}
    exec(compile(src[start:end], "reference_remap", "exec"), ns)          # the reference function, unmodified, in isolation
    ck = _fake_ckpt()
    ref, got = ns["remap_checkpoint_keys"](dict(ck)), to_dense_state_dict(ck)
    assert list(ref.keys()) == list(got.keys())
    for k in ref:
        assert ref[k].shape == got[k].shape and torch.equal(ref[k].contiguous(), got[k].contiguous()), k


def test_save_load_roundtrip(native_lib, tmp_path):
    from mmearth_train_b200.checkpoint import load_checkpoint, save_checkpoint

    class Tiny(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.arange(6.0).reshape(2, 3))

    a, b = Tiny(), Tiny()
    with torch.no_grad():
        b.w.zero_()
    p = str(tmp_path / "ck.pth")
    save_checkpoint(p, a, epoch=7, extra={"args": {"model": "convnextv2_atto"}})
    rest = load_checkpoint(p, b)
    assert torch.equal(a.w, b.w) and rest["epoch"] == 7 and rest["args"]["model"] == "convnextv2_atto"
