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


def test_save_model_rotation_and_auto_resume(tmp_path):
    """helpers.save_model / auto_load_model semantics (helpers.py:529-610) with stand-in objects: file naming, the rotation
    window, newest-checkpoint pick, start_epoch, optimizer and scaler state."""
    from argparse import Namespace
    import os
    from mmearth_train_b200 import checkpoint as ck

    class Scaler:
        def __init__(self):
            self.scale = 65536.0

        def state_dict(self):
            return {"scale": self.scale}

        def load_state_dict(self, sd):
            self.scale = sd["scale"]

    torch.manual_seed(0)
    net = torch.nn.Linear(4, 3)
    opt = torch.optim.AdamW(net.parameters(), lr=1e-3)
    scaler = Scaler()
    args = Namespace(output_dir=str(tmp_path / "run"), save_ckpt_num=2, save_ckpt_freq=1, auto_resume=True, resume="", start_epoch=0)
    for epoch in range(4):
        net(torch.randn(2, 4)).sum().backward()
        opt.step()
        scaler.scale = 65536.0 * (epoch + 1)
        ck.save_model(args, epoch, net, net, opt, scaler)
    assert sorted(os.listdir(args.output_dir)) == ["checkpoint-2.pth", "checkpoint-3.pth"]      # window of 2
    ck.save_model(args, "best", net, net, opt, scaler)                                          # non-integer tags are kept
    want = {k: v.clone() for k, v in net.state_dict().items()}
    step_before = opt.state_dict()["state"][0]["step"]

    net2 = torch.nn.Linear(4, 3)
    opt2 = torch.optim.AdamW(net2.parameters(), lr=1e-3)
    scaler2 = Scaler()
    args2 = Namespace(output_dir=args.output_dir, auto_resume=True, resume="", start_epoch=0)
    ck.auto_load_model(args2, net2, net2, opt2, scaler2)
    assert args2.resume.endswith("checkpoint-3.pth") and args2.start_epoch == 4
    assert all(torch.equal(net2.state_dict()[k], v) for k, v in want.items())
    assert opt2.state_dict()["state"][0]["step"] == step_before and scaler2.scale == 65536.0 * 4
    blob = torch.load(args2.resume, map_location="cpu", weights_only=False)
    assert set(blob) == {"model", "optimizer", "epoch", "scaler", "args"}                      # helpers.py:545-551

    args3 = Namespace(output_dir=str(tmp_path / "empty"), auto_resume=True, resume="", start_epoch=0)
    ck.auto_load_model(args3, net2, net2, opt2, scaler2)                                        # nothing to resume: a no-op
    assert args3.start_epoch == 0 and args3.resume == ""


def test_checkpoint_written_by_the_reference_resumes_here(tmp_path):
    """A checkpoint written by the unmodified helpers.save_model is picked up by auto_load_model (build container only)."""
    from argparse import Namespace
    from oracle import ref_harness
    if not ref_harness.reference_available():
        pytest.skip("needs /root/reference")
    helpers = ref_harness.import_toplevel("helpers")
    from mmearth_train_b200 import checkpoint as ck
    torch.manual_seed(1)
    net = torch.nn.Linear(5, 2)
    opt = torch.optim.AdamW(net.parameters(), lr=1e-3)
    net(torch.randn(3, 5)).sum().backward()
    opt.step()
    scaler = helpers.NativeScalerWithGradNormCount("cpu")
    args = Namespace(output_dir=str(tmp_path), save_ckpt_num=3, save_ckpt_freq=1)
    helpers.save_model(args, 7, net, net, opt, scaler)
    ours = ck.__dict__                                                       # and ours writes the same entries
    args_b = Namespace(output_dir=str(tmp_path / "b"), save_ckpt_num=3, save_ckpt_freq=1)
    ours["save_model"](args_b, 7, net, net, opt, scaler)
    a = torch.load(str(tmp_path / "checkpoint-7.pth"), map_location="cpu", weights_only=False)
    b = torch.load(str(tmp_path / "b" / "checkpoint-7.pth"), map_location="cpu", weights_only=False)
    assert set(a) == set(b) and a["epoch"] == b["epoch"] == 7
    assert all(torch.equal(a["model"][k], b["model"][k]) for k in a["model"])

    net2 = torch.nn.Linear(5, 2)
    opt2 = torch.optim.AdamW(net2.parameters(), lr=1e-3)
    args2 = Namespace(output_dir=str(tmp_path), auto_resume=True, resume="", start_epoch=0)
    ck.auto_load_model(args2, net2, net2, opt2, helpers.NativeScalerWithGradNormCount("cpu"))
    assert args2.start_epoch == 8 and torch.equal(net2.weight, net.weight)


def test_reference_param_order_matches_the_reference_module(native_lib):
    """``optim.reference_param_order`` against ``named_parameters()`` of the unmodified reference ``FCMAE`` (fixture
    ``tests/golden/param_order.json``, made by importing the reference in the build container; the optimizer state of a
    reference checkpoint is indexed by that order, helpers.py:541-547)."""
    import json
    from mmearth_train_b200.optim import reference_param_order
    fx = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "param_order.json")))
    for tag, row in fx.items():
        assert reference_param_order(sorted(row["names"]), row["out_modalities"]) == row["names"], tag
    if os.path.isfile("/root/reference/models/fcmae.py"):      # and live, where the reference exists
        from oracle import ref_harness as rh
        ref, args = rh.build_reference_model(model="convnextv2_femto")
        names = [n for n, _ in ref.named_parameters()]
        assert reference_param_order(sorted(names), list(args.out_modalities.keys())) == names
