"""Host-side pieces of the loop that need no GPU: meters and the per-iteration schedule."""
import math
from argparse import Namespace

from mmearth_train_b200 import engine
from mmearth_train_b200.optim import cosine_lr


def test_meters_average_like_the_reference():
    m = engine.SmoothedValue(window_size=3)
    for v in (1.0, 2.0, 3.0, 4.0):
        m.update(v)
    assert m.value == 4.0 and m.avg == 3.0 and m.global_avg == 2.5        # window of 3 vs the whole series
    log = engine.MetricLogger()
    log.update(loss=2.0, skipped=None)
    log.update(loss=4.0)
    assert log.meters["loss"].global_avg == 3.0 and "skipped" not in log.meters
    assert "loss: 3.0000 (3.0000)" in str(log)
    log.synchronize_between_processes()                                    # single process: a no-op


def test_cosine_schedule_matches_reference_formula():
    a = Namespace(lr=1.5e-4, min_lr=1e-6, warmup_epochs=40, epochs=800)
    assert cosine_lr(0, a.lr, a.min_lr, a.warmup_epochs, a.epochs) == 0.0
    assert cosine_lr(20, a.lr, a.min_lr, a.warmup_epochs, a.epochs) == a.lr * 0.5
    assert abs(cosine_lr(40, a.lr, a.min_lr, a.warmup_epochs, a.epochs) - a.lr) < 1e-18
    mid = a.min_lr + (a.lr - a.min_lr) * 0.5 * (1.0 + math.cos(math.pi * 0.5))
    assert abs(cosine_lr(420, a.lr, a.min_lr, a.warmup_epochs, a.epochs) - mid) < 1e-18
    assert abs(cosine_lr(800, a.lr, a.min_lr, a.warmup_epochs, a.epochs) - a.min_lr) < 1e-18
