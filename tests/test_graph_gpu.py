"""GraphedStep: one training iteration (forward + backward + AdamW) captured in a CUDA graph and replayed, against the
eager step on the same weights, batches, masks and learning rates."""
import pytest
import torch

from oracle import fcmae_oracle as fo
from tests import golden_util as gu

pytestmark = pytest.mark.gpu


def _pair():
    from tests.test_parity_gpu import build_native
    z, meta, orc, batch, noise = gu.inputs("atto_p8_all_unc")
    return build_native(meta["cfg"], orc, 3), build_native(meta["cfg"], orc, 3), batch, noise


def test_graphed_step_follows_the_eager_trajectory():
    import mmearth_train_b200 as mp
    from mmearth_train_b200.optim import FlatAdamW
    a, b, batch, noise = _pair()
    oa, ob = FlatAdamW(a, lr=3e-4), FlatAdamW(b, lr=3e-4)
    init = a.flat_params.clone()
    batches = [{k: v.cuda() for k, v in fo.synthetic_batch(2, 56, seed=600 + i, nan_frac=0.05).items()} for i in range(3)]
    a.noise_override = noise.cuda().clone()                    # static device tensor: refreshed in place between replays
    step = mp.GraphedStep(a, oa, batches[0], mask_ratio=0.6)
    assert torch.equal(a.flat_params, init) and oa.t == 0      # the warm-up iterations were rolled back
    gen = torch.Generator().manual_seed(99)
    for i in range(6):
        nz = torch.randn(noise.shape, generator=gen)
        lr = 3e-4 * (1.0 - 0.1 * i)
        a.noise_override.copy_(nz)
        la = step(batches[i % 3], lr=lr)
        b.noise_override = nz
        lb = b(batches[i % 3], mask_ratio=0.6)[0]
        lb.backward()
        ob.lr = lr
        ob.step()
        ob.zero_grad()
        assert abs(float(la) - float(lb)) <= 1e-4 * abs(float(lb)), (i, float(la), float(lb))
        assert torch.equal(step.mask, b.last_run["mask"])
        for j, m in enumerate(a.out_modalities):
            assert abs(float(step.losses[j]) - float(b.last_run["losses"][j])) <= 1e-4 * abs(float(b.last_run["losses"][j])) + 1e-7, m
    assert oa.t == 6 == ob.t
    assert gu.rel_err(a.flat_params - init, b.flat_params - init) < 0.05
    with pytest.raises(ValueError):
        step({k: v[:1] for k, v in batches[0].items()})


def test_graphed_step_draws_a_new_mask_every_replay():
    import mmearth_train_b200 as mp
    from mmearth_train_b200.optim import FlatAdamW
    a, _, batch, _ = _pair()
    dev = {k: v.cuda() for k, v in batch.items()}
    step = mp.GraphedStep(a, FlatAdamW(a, lr=1e-4), dev)
    masks = []
    for _ in range(4):
        loss = step(dev)
        assert torch.isfinite(loss)
        masks.append(step.mask.clone())
        assert torch.all(step.mask.sum(1) == 30)               # int(49 * 0.4) = 19 visible patches per sample
    assert any(not torch.equal(masks[0], m) for m in masks[1:])
