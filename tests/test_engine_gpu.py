"""The loop around the step: FlatGradScaler (helpers.NativeScalerWithGradNormCount semantics
without host syncs), FlatAdamW.step_dev, and engine.train_one_epoch against a plain hand-written loop."""
import copy
from argparse import Namespace

import pytest
import torch

from oracle import fcmae_oracle as fo
from tests import golden_util as gu

pytestmark = pytest.mark.gpu


def _model(backend=3):
    from tests.test_parity_gpu import build_native
    z, meta, orc, batch, noise = gu.inputs("atto_p8_all_unc")
    return build_native(meta["cfg"], orc, backend), batch, noise


def test_scaler_step_equals_unscaled_step_and_skips_non_finite():
    from mmearth_train_b200.optim import FlatAdamW, FlatGradScaler
    a, batch, noise = _model()
    b, _, _ = _model()
    dev = {k: v.cuda() for k, v in batch.items()}
    oa, ob = FlatAdamW(a, lr=3e-4), FlatAdamW(b, lr=3e-4)
    init = a.flat_params.clone()
    scaler = FlatGradScaler("cuda")
    for step in range(5):
        nz = torch.randn(noise.shape, generator=torch.Generator().manual_seed(100 + step))
        a.noise_override = b.noise_override = nz
        la = a(dev)[0]
        norm = scaler(la, oa, parameters=a.parameters())
        oa.zero_grad()
        lb = b(dev)[0]
        lb.backward()
        ref_norm = ob.grad_norm()
        ob.step()
        ob.zero_grad()
        assert abs(float(la) - float(lb)) <= 1e-4 * abs(float(lb)), step
        assert abs(float(norm) - float(ref_norm)) <= 1e-3 * float(ref_norm), step     # un-scaled norm (helpers.get_grad_norm_)
    # 2**16 scaling is exact in fp32: the two trajectories differ by atomics order only (compare the UPDATES: Adam moves
    # every element by ~lr per step, which is invisible next to the std-1 weights themselves)
    assert gu.rel_err(a.flat_params - init, b.flat_params - init) < 0.05
    assert oa.t == 5 and scaler.get_scale() == 65536.0

    # overflow: the scaled gradient is not finite -> nothing changes, the step is not counted, the scale backs off
    big = FlatGradScaler("cuda", init_scale=3e38)
    before, m_before = a.flat_params.clone(), oa.exp_avg.clone()
    norm = big(a(dev)[0], oa, parameters=a.parameters())
    oa.zero_grad()
    assert not torch.isfinite(norm)
    assert torch.equal(a.flat_params, before) and torch.equal(oa.exp_avg, m_before)
    assert oa.t == 5 and big.get_scale() == torch.tensor(3e38).item() * 0.5
    # growth after `growth_interval` clean steps, clipping folded into the step
    grow = FlatGradScaler("cuda", init_scale=1024.0, growth_interval=2)
    for _ in range(2):
        n = grow(a(dev)[0], oa, clip_grad=0.5, parameters=a.parameters())
        oa.zero_grad()
        assert torch.isfinite(n) and float(n) > 0                 # the norm BEFORE clipping, like clip_grad_norm_
    assert grow.get_scale() == 2048.0 and oa.t == 7
    sd = grow.state_dict()
    other = FlatGradScaler("cuda")
    other.load_state_dict(sd)
    assert other.get_scale() == 2048.0


def test_reference_scaler_and_scheduler_drive_flat_adamw():
    """The reference's own NativeScalerWithGradNormCount body (torch GradScaler: scale -> backward -> unscale_ -> step ->
    update, helpers.py:485-500) and adjust_learning_rate (param_groups, helpers.py:660-664) work on FlatAdamW as is."""
    from mmearth_train_b200.optim import FlatAdamW
    a, batch, noise = _model()
    b, _, _ = _model()
    dev = {k: v.cuda() for k, v in batch.items()}
    oa, ob = FlatAdamW(a, lr=1.0), FlatAdamW(b, lr=3e-4)
    init = a.flat_params.clone()
    a.noise_override = b.noise_override = noise
    for group in oa.param_groups:                    # what helpers.adjust_learning_rate does
        group["lr"] = 3e-4
    assert oa.lr == 3e-4
    scaler = torch.amp.GradScaler("cuda")
    for _ in range(3):
        scaler.scale(a(dev)[0]).backward()
        scaler.unscale_(oa)
        scaler.step(oa)
        scaler.update()
        oa.zero_grad()
        b(dev)[0].backward()
        ob.step()
        ob.zero_grad()
    assert oa.t == 3
    assert gu.rel_err(a.flat_params - init, b.flat_params - init) < 0.05


@pytest.mark.parametrize("update_freq", [1, 2])
def test_train_one_epoch_matches_hand_written_loop(update_freq):
    from mmearth_train_b200 import engine
    from mmearth_train_b200.optim import FlatAdamW, FlatGradScaler, cosine_lr
    a, batch, _ = _model()
    b, _, _ = _model()
    n_iter = 6
    batches = []
    for i in range(n_iter):
        d = fo.synthetic_batch(2, 56, seed=500 + i, nan_frac=0.05)
        batches.append({k: v.pin_memory() for k, v in d.items()})
    args = Namespace(update_freq=update_freq, lr=3e-4, min_lr=1e-6, warmup_epochs=1, epochs=4, mask_ratio=0.6, no_ffcv=True)
    oa, ob = FlatAdamW(a, lr=args.lr), FlatAdamW(b, lr=args.lr)
    init = a.flat_params.clone()

    torch.manual_seed(7)
    loader = [(i, d) for i, d in enumerate(batches)]             # no_ffcv loaders yield (index, dict): engine_pretrain.py:50
    stats, loss_dict, log_vars, normalized = engine.train_one_epoch(
        a, None, loader, oa, torch.device("cuda"), epoch=1, use_mixed=False, loss_scaler=FlatGradScaler("cuda"), args=args,
        quiet=True)

    torch.manual_seed(7)
    losses, lrs = [], []
    for i, d in enumerate(batches):
        if i % update_freq == 0:
            ob.lr = cosine_lr(i / n_iter + 1, args.lr, args.min_lr, args.warmup_epochs, args.epochs)
        lrs.append(ob.lr)
        out = b({k: v.cuda() for k, v in d.items()}, mask_ratio=0.6)
        (out[0] / update_freq).backward()
        losses.append(float(out[0]))
        if (i + 1) % update_freq == 0:
            ob.step()
            ob.zero_grad()
    assert abs(stats["loss"] - sum(losses) / n_iter) <= 5e-4 * abs(sum(losses) / n_iter)
    assert abs(stats["lr"] - sum(lrs) / n_iter) < 1e-12
    assert oa.t == n_iter // update_freq == ob.t
    assert gu.rel_err(a.flat_params - init, b.flat_params - init) < 0.05
    assert set(loss_dict) == set(a.out_modalities)
    for m in a.out_modalities:
        assert abs(loss_dict[m] - float(out[3][m])) <= 1e-3 * abs(float(out[3][m])) + 1e-6, m
    assert len(log_vars) == 12 and normalized.shape == (12,)


def test_optimizer_state_interchanges_with_torch_adamw():
    """a checkpoint written by the reference's ``save_model`` holds ``torch.optim.AdamW.state_dict()`` over
    timm's two parameter groups (main_pretrain.py:312-320).  FlatAdamW reads and writes exactly that layout: a torch AdamW
    run is resumed by FlatAdamW and the other way round, with identical parameters afterwards."""
    from mmearth_train_b200.optim import FlatAdamW, reference_param_order
    a, batch, noise = _model()
    b, _, _ = _model()
    dev = {k: v.cuda() for k, v in batch.items()}
    a.noise_override = b.noise_override = noise
    named = {n: p for n, p in b.named_parameters() if n != "_ddp_token"}
    order = reference_param_order(list(named), b.out_modalities)
    no_decay = [named[n] for n in order if named[n].ndim <= 1 or n.endswith(".bias")]
    decay = [named[n] for n in order if not (named[n].ndim <= 1 or n.endswith(".bias"))]
    ot = torch.optim.AdamW([{"params": no_decay, "weight_decay": 0.0}, {"params": decay, "weight_decay": 0.05}], lr=3e-4,
                           betas=(0.9, 0.95))
    oa = FlatAdamW(a, lr=3e-4, betas=(0.9, 0.95), weight_decay=0.05)
    for _ in range(2):                                  # torch AdamW on b, FlatAdamW on a: same trajectory
        b(dev)[0].backward(); ot.step(); ot.zero_grad(set_to_none=False)
        a(dev)[0].backward(); oa.step(); oa.zero_grad()
    sd_t, sd_f = ot.state_dict(), oa.state_dict()
    assert [len(g["params"]) for g in sd_t["param_groups"]] == [len(g["params"]) for g in sd_f["param_groups"]]
    for i in sd_t["state"]:
        assert sd_f["state"][i]["exp_avg"].shape == sd_t["state"][i]["exp_avg"].shape, i
        assert gu.rel_err(sd_f["state"][i]["exp_avg"], sd_t["state"][i]["exp_avg"]) < 2e-2, i
        assert float(sd_f["state"][i]["step"]) == float(sd_t["state"][i]["step"]) == 2.0
    # resume the torch run with FlatAdamW and the flat run with torch AdamW
    c, _, _ = _model()
    c.noise_override = noise
    c.load_state_dict(b.state_dict())
    oc = FlatAdamW(c, lr=1.0, betas=(0.9, 0.95), weight_decay=0.05)
    oc.load_state_dict(sd_t)
    assert oc.t == 2 and oc.lr == 3e-4
    ot.load_state_dict(sd_f)                            # b continues from a's moments (the same up to atomics order)
    b(dev)[0].backward(); ot.step()
    c(dev)[0].backward(); oc.step()
    assert gu.rel_err(c.flat_params, b.flat_params) < 1e-5
    with pytest.raises(ValueError):
        oc.load_state_dict({"state": {}, "param_groups": [{"params": [0]}, {"params": [1]}]})
