"""engine.train_one_epoch against golden trajectories of the UNMODIFIED reference loop (engine_pretrain.train_one_epoch
around the unmodified reference FCMAE, AdamW as main_pretrain.py builds it, NativeScalerWithGradNormCount), made by
oracle/make_engine_golden.py: BASELINE.json configs[0] (atto, S2 -> S2, 56/p8, bs 8, the reference's CPU-runnable case) and
a 12-modality uncertainty-weighted run with gradient accumulation.  The loop under test is the product's host code; the model
inside it here is the CPU oracle (the native module needs a GPU: tests/test_engine_gpu.py runs the same loop around it)."""
import json
import os

import pytest
import torch

from mmearth_train_b200 import engine
from oracle import make_engine_golden as meg

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 2e-5


class _Replay(torch.nn.Module):
    """The oracle behind the reference's call signature, drawing the fixture's noise instead of torch.randn."""

    def __init__(self, orc, noises):
        super().__init__()
        self.orc, self.args, self.queue = orc, orc.args, list(noises)

    def forward(self, samples, labels=None, mask_ratio=0.6):
        return self.orc(samples, mask_ratio=mask_ratio, noise=self.queue.pop(0))


class _CpuScaler:
    """helpers.NativeScalerWithGradNormCount on CPU (GradScaler disabled, helpers.py:473): backward, step on update."""

    def __call__(self, loss, optimizer, clip_grad=None, parameters=None, create_graph=False, update_grad=True):
        loss.backward()
        if update_grad:
            optimizer.step()


def _close(a, b, tol=TOL):
    return abs(a - b) <= tol * abs(b) + 1e-9


@pytest.mark.parametrize("case", list(meg.CASES))
@pytest.mark.parametrize("lag", [1, 3])
def test_loop_reproduces_reference_engine_trajectory(case, lag):
    gold = json.load(open(os.path.join(GOLDEN_DIR, f"engine_{case}.json")))
    cfg = gold["cfg"]
    assert cfg == meg.CASES[case] and gold["loop_args"] == meg.LOOP_ARGS, "fixture is stale: python -m oracle.make_engine_golden"
    orc, batches, noises = meg.case_inputs(cfg)
    cs = gold["input_checksum"]
    assert _close(float(sum(b["sentinel2"].double().sum() for b in batches)), cs["s2"], 1e-9), "torch CPU generator changed"
    assert _close(float(sum(n.double().sum() for n in noises)), cs["noise"], 1e-9)
    args = meg.loop_args(cfg)
    model = _Replay(orc, noises)
    optimizer = torch.optim.AdamW(meg.param_groups_weight_decay(orc, args.weight_decay), lr=args.lr, betas=(0.9, 0.95))
    writer = meg._Writer()
    loader = [(i, b) for i, b in enumerate(batches)]
    stats, loss_dict, log_vars, normalized = engine.train_one_epoch(
        model, None, loader, optimizer, torch.device("cpu"), cfg["epoch"], False, _CpuScaler(), log_writer=writer, args=args,
        lag=lag, quiet=True)

    assert _close(stats["loss"], gold["stats"]["loss"]) and _close(stats["lr"], gold["stats"]["lr"], 1e-12)
    assert set(loss_dict) == set(gold["loss_dict"])
    for m, v in gold["loss_dict"].items():
        assert _close(loss_dict[m], v), m
    if gold["log_vars"] is None:
        assert log_vars is None and normalized is None
    else:
        assert all(_close(a, b) for a, b in zip(log_vars, gold["log_vars"]))
        assert all(_close(float(a), b) for a, b in zip(normalized, gold["normalized"]))
    # what the loop sends to the log writer: same steps, same learning rates, same per-iteration losses (engine_pretrain.py:104-112)
    for head, key in (("loss", "train_loss"), ("opt", "lr")):
        got = [r for r in writer.rows if r["head"] == head]
        want = [r for r in gold["logged"] if r["head"] == head]
        assert [r["step"] for r in got] == [r["step"] for r in want], head
        for g, w in zip(got, want):
            assert _close(g[key], w[key], TOL if key == "train_loss" else 1e-12), (head, g, w)
    norm = float(torch.sqrt(sum((p.detach().double() ** 2).sum() for p in {id(p): p for p in orc.parameters()}.values())))
    assert _close(norm, gold["final_param_norm"], 1e-6)
    assert optimizer.param_groups[0]["lr"] == optimizer.param_groups[1]["lr"]


@pytest.mark.skipif(not meg.ref_harness.reference_available(), reason="needs /root/reference (build container only)")
def test_fixture_generator_still_matches_the_live_reference():
    """Re-runs the unmodified reference loop for the small case and compares with the committed fixture."""
    gold = json.load(open(os.path.join(GOLDEN_DIR, "engine_cfg1.json")))
    live = meg.run_reference_engine(meg.CASES["cfg1"])
    assert _close(live["stats"]["loss"], gold["stats"]["loss"], 1e-6)
    assert [r["step"] for r in live["logged"]] == [r["step"] for r in gold["logged"]]
    assert _close(live["final_param_norm"], gold["final_param_norm"], 1e-9)


def test_loop_device_branch_with_stand_in_transport(monkeypatch):
    """The branch the loop takes for a CUDA device (DevicePrefetcher + LossReader + the module's own loss vector), exercised
    without a GPU: the two transport classes are replaced by host stand-ins with the same interface and the model publishes
    ``last_run["losses"]`` like the native module.  Same golden trajectory, so the branch's bookkeeping is the checked one."""
    gold = json.load(open(os.path.join(GOLDEN_DIR, "engine_all_unc_accum.json")))
    cfg = gold["cfg"]
    orc, batches, noises = meg.case_inputs(cfg)
    args = meg.loop_args(cfg)
    made = {}

    class Prefetch:
        def __init__(self, src, device, depth=2):
            made["prefetch"] = (device.type, depth)
            self.src = src

        def __iter__(self):
            return iter(self.src)

    class Reader(engine._LaggedHostReader):
        def __init__(self, device, depth=2, width=1):
            made["reader"] = (device.type, depth, width)
            super().__init__(depth)

        def push(self, value):
            assert value.shape == (made["reader"][2],)
            return super().push(value)

    class Native(_Replay):
        def forward(self, samples, labels=None, mask_ratio=0.6):
            out = super().forward(samples, labels, mask_ratio)
            per = torch.stack([v.detach() for v in out[3].values()])
            self.last_run = {"losses": torch.cat([per, out[5].detach(), out[0].detach().reshape(1)])}
            return out

    monkeypatch.setattr(engine, "DevicePrefetcher", Prefetch)
    monkeypatch.setattr(engine, "LossReader", Reader)
    model = Native(orc, noises)
    model.out_modalities, model.loss_aggr = list(orc.args.out_modalities), "uncertainty"
    optimizer = torch.optim.AdamW(meg.param_groups_weight_decay(orc, args.weight_decay), lr=args.lr, betas=(0.9, 0.95))
    writer = meg._Writer()
    stats, loss_dict, log_vars, normalized = engine.train_one_epoch(
        model, None, [(i, b) for i, b in enumerate(batches)], optimizer, torch.device("cuda"), cfg["epoch"], False, _CpuScaler(),
        log_writer=writer, args=args, lag=2, quiet=True)
    assert made == {"prefetch": ("cuda", 2), "reader": ("cuda", 2, 25)}
    assert _close(stats["loss"], gold["stats"]["loss"]) and _close(stats["lr"], gold["stats"]["lr"], 1e-12)
    assert all(_close(loss_dict[m], v) for m, v in gold["loss_dict"].items())
    assert all(_close(float(a), b) for a, b in zip(normalized, gold["normalized"]))
    got = [r for r in writer.rows if r["head"] == "loss"]
    want = [r for r in gold["logged"] if r["head"] == "loss"]
    assert [r["step"] for r in got] == [r["step"] for r in want]
    assert all(_close(g["train_loss"], w["train_loss"]) for g, w in zip(got, want))


def test_fit_interrupted_and_resumed_equals_uninterrupted(tmp_path):
    """Checkpoint / resume through the epoch loop (main_pretrain.py:322-366 + helpers.save_model / auto_load_model): a run that
    dies in the middle of its third epoch and is started again with auto_resume ends with the same per-epoch statistics and
    the same weights, bit for bit, as a run that was never interrupted."""
    from argparse import Namespace
    from oracle import fcmae_oracle as fo

    class Scaler(_CpuScaler):
        def state_dict(self):
            return {}

        def load_state_dict(self, sd):
            pass

    class Crash(Exception):
        pass

    def launch(out_dir, crash_at_step=None):
        """One process lifetime: fresh objects, weights / optimizer state come from the newest checkpoint if there is one."""
        orc = fo.build_oracle(out_modalities=["sentinel2"], loss_aggr="unweighted")
        fo.init_like_reference(orc, seed=3)
        batches = [fo.synthetic_batch(2, 56, ["sentinel2"], seed=60 + i) for i in range(2)]
        opt = torch.optim.AdamW(meg.param_groups_weight_decay(orc, 0.05), lr=3e-4, betas=(0.9, 0.95))
        args = Namespace(update_freq=1, lr=3e-4, min_lr=1e-6, warmup_epochs=1, epochs=4, mask_ratio=0.6, no_ffcv=True,
                         output_dir=str(out_dir), auto_resume=True, resume="", save_ckpt=True, save_ckpt_freq=1, save_ckpt_num=2)

        class Model(_Replay):
            def forward(self, samples, labels=None, mask_ratio=0.6):
                step = int(next(iter(opt.state.values()))["step"]) if len(opt.state) else 0     # restored with the optimizer
                if crash_at_step is not None and step == crash_at_step:
                    raise Crash()
                g = torch.Generator().manual_seed(1000 + step)                                   # masks resumable by step
                return self.orc(samples, mask_ratio=mask_ratio, noise=torch.randn(2, 49, generator=g))

        hist = engine.fit(Model(orc, []), orc, [(i, b) for i, b in enumerate(batches)], opt, Scaler(), torch.device("cpu"), args)
        norm = float(torch.sqrt(sum((p.detach().double() ** 2).sum() for p in {id(p): p for p in orc.parameters()}.values())))
        return hist, norm

    whole, norm_a = launch(tmp_path / "a")
    with pytest.raises(Crash):
        launch(tmp_path / "b", crash_at_step=5)                     # 2 iterations per epoch: dies inside epoch 2
    assert sorted(os.listdir(tmp_path / "b")) == ["checkpoint-0.pth", "checkpoint-1.pth"]
    rest, norm_b = launch(tmp_path / "b")
    assert [h["epoch"] for h in whole] == [0, 1, 2, 3] and [h["epoch"] for h in rest] == [2, 3]
    for a, b in zip(whole[2:], rest):
        assert a["train_loss"] == b["train_loss"] and a["train_lr"] == b["train_lr"], (a["epoch"], a["train_loss"], b["train_loss"])
    assert norm_a == norm_b
    assert sorted(os.listdir(tmp_path / "b")) == ["checkpoint-2.pth", "checkpoint-3.pth"]
    assert whole[-1]["train_loss"] < whole[0]["train_loss"]
