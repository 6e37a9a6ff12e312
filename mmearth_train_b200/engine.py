"""The training loop on the far side of the step: ``engine_pretrain.train_one_epoch`` without its per-step host syncs.

The reference loop (``engine_pretrain.py:21-126``) works unchanged with the native ``FCMAE`` -- the module keeps the
reference's surface -- but it drains the GPU several times per iteration: ``loss.item()`` plus one ``.item()`` per modality
(``:72-75``), ``normalized_loss_list.cpu()`` (``:66-70``), ``GradScaler.step``'s non-finite check (``helpers.py:498``) and
``torch.cuda.empty_cache()`` (``:96``), on top of H2D copies issued on the compute stream (``:59-61``).  With a 7.8 ms step
those stalls are a visible fraction of the iteration.  ``train_one_epoch`` below has the same signature, bookkeeping and return
value, and differs only in how the host learns what happened:

* batches go through ``DevicePrefetcher`` (copy stream, persistent buffers);
* every step's ``2T + 1`` loss values (per modality, weighted, total -- one contiguous device vector written by the loss
  kernel) are copied to pinned memory without blocking and read ``lag`` steps later (``LossReader``), so the meters see every
  step's values, late by ``lag`` iterations; the non-finite check (``:77-79``) fires with the same delay;
* ``FlatGradScaler`` + ``FlatAdamW.step_dev`` keep the scale / skip decision on the device;
* no ``empty_cache``: nothing is allocated per step.

SURVEY.md section 8f ranks 1, 2 and 4.

The loop itself computes nothing: it moves batches and reads losses.  With a CUDA device it uses the two transport classes above;
with ``device.type == "cpu"`` it iterates host tensors directly -- that branch exists so the tests can replay golden trajectories
of the unmodified reference loop through THIS loop around the CPU oracle (``tests/test_engine_golden.py``).  It is not a CPU
fallback of the step: the native ``FCMAE`` refuses non-CUDA tensors whatever loop drives it.
"""
from __future__ import annotations

import math
import sys
import time
from collections import defaultdict, deque
from typing import Dict, Iterable, List, Optional

import torch

from .data import DevicePrefetcher, LossReader
from .optim import cosine_lr


class SmoothedValue:
    """Windowed + global average of a series (the part of ``helpers.SmoothedValue`` the loop uses, ``helpers.py:33-96``)."""

    def __init__(self, window_size: int = 20):
        self.deque = deque(maxlen=window_size)
        self.total, self.count = 0.0, 0

    def update(self, value: float, n: int = 1) -> None:
        self.deque.append(value)
        self.count += n
        self.total += value * n

    @property
    def value(self) -> float:
        return self.deque[-1]

    @property
    def avg(self) -> float:
        return sum(self.deque) / max(len(self.deque), 1)

    @property
    def global_avg(self) -> float:
        return self.total / max(self.count, 1)

    def synchronize_between_processes(self) -> None:
        """Sum count / total over ranks (``helpers.py:51-63``)."""
        dist = torch.distributed
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return
        dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
        t = torch.tensor([self.count, self.total], dtype=torch.float64, device=dev)
        dist.all_reduce(t)
        self.count, self.total = int(t[0].item()), float(t[1].item())


class MetricLogger:
    """Named meters (``helpers.MetricLogger``, ``helpers.py:99-186``), without the reference's per-iteration CUDA memory query."""

    def __init__(self, delimiter: str = "  "):
        self.meters: Dict[str, SmoothedValue] = defaultdict(SmoothedValue)
        self.delimiter = delimiter

    def add_meter(self, name: str, meter: SmoothedValue) -> None:
        self.meters[name] = meter

    def update(self, **kwargs) -> None:
        for k, v in kwargs.items():
            if v is None:
                continue
            self.meters[k].update(float(v))

    def synchronize_between_processes(self) -> None:
        for m in self.meters.values():
            m.synchronize_between_processes()

    def __str__(self) -> str:
        return self.delimiter.join(f"{k}: {m.avg:.4f} ({m.global_avg:.4f})" for k, m in self.meters.items())


def _len_or_none(it) -> Optional[int]:
    try:
        return len(it)
    except TypeError:
        return None


class _LaggedHostReader:
    """``LossReader``'s interface for tensors that already live on the host (``device.type == "cpu"``: the loop run around
    a CPU model, which is how the tests replay the reference engine's golden trajectories): same ``lag`` semantics."""

    def __init__(self, depth: int):
        self._q = deque()
        self.depth = depth

    def push(self, value: torch.Tensor):
        self._q.append(value.detach().reshape(-1).tolist())
        return self._q.popleft() if len(self._q) > self.depth else None

    def flush(self) -> list:
        out = list(self._q)
        self._q.clear()
        return out


def _loss_vector(core, out) -> torch.Tensor:
    """``[per-modality losses (T), weighted losses (T), total]``: the native module's loss kernels write exactly this vector
    (``FCMAE.last_run["losses"]``); for any other module with the reference's return tuple it is assembled here."""
    run = getattr(core, "last_run", None)
    if isinstance(run, dict) and "losses" in run:
        return run["losses"]
    loss, _pred, _mask, loss_dict, _log_vars, weighted = out
    per = torch.stack([v.detach().reshape(()) for v in loss_dict.values()])
    w = weighted.detach().reshape(-1) if weighted is not None else torch.zeros_like(per)
    return torch.cat([per, w, loss.detach().reshape(1)])


def train_one_epoch(model: torch.nn.Module, modalities, data_loader: Iterable, optimizer, device: torch.device, epoch: int,
                    use_mixed: bool, loss_scaler, log_writer=None, args=None, lag: int = 2, print_freq: int = 20,
                    quiet: bool = False):
    """``engine_pretrain.train_one_epoch`` (``engine_pretrain.py:21-126``): same arguments (``lag`` / ``print_freq`` / ``quiet``
    are extra), same return ``(averaged stats, loss_dict, log_var_list, normalized_loss_list)``.

    ``optimizer`` = ``FlatAdamW`` and ``loss_scaler`` = ``FlatGradScaler`` (or any pair with the reference's call
    signatures: ``helpers.NativeScalerWithGradNormCount`` + ``torch.optim.AdamW`` also work, with their syncs).
    """
    if use_mixed:
        raise NotImplementedError("the native step computes in fp32 (the reference runs this path with use_mixed False: "
                                  "MinkowskiEngine has no half kernels)")
    model.train()
    core = model.module if hasattr(model, "module") else model
    device = torch.device(device)
    on_gpu = device.type == "cuda"
    metric_logger = MetricLogger()
    metric_logger.add_meter("lr", SmoothedValue(window_size=1))
    header = "Epoch: [{}]".format(epoch)
    update_freq = args.update_freq
    n_iter = _len_or_none(data_loader)
    if n_iter is None:
        raise TypeError("data_loader needs __len__ (the per-iteration schedule divides by it, engine_pretrain.py:53-56)")
    names = list(getattr(core, "out_modalities", None) or core.args.out_modalities.keys())
    T = len(names)
    uncertainty = getattr(core, "loss_aggr", None) or core.args.loss_aggr
    dist = torch.distributed
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    # the loss vector carries one more entry when world > 1: the mean over ranks of this step's total loss
    # (helpers.all_reduce_mean, engine_pretrain.py:103), reduced on the device by EVERY rank EVERY step
    width = 2 * T + (2 if world > 1 else 1)
    reader = LossReader(device, depth=max(lag, 1), width=width) if on_gpu else _LaggedHostReader(max(lag, 1))
    pending = deque()                      # (epoch_1000x, is an update step) of the iterations whose losses are in flight
    last = {"loss_dict": None, "weighted": None}

    def consume(vals: Optional[List[float]]) -> None:
        if vals is None:
            return
        step, was_update = pending.popleft()
        loss_value = vals[2 * T]
        if not math.isfinite(loss_value):                                       # engine_pretrain.py:77-79
            print("Loss is {}, stopping training".format(loss_value))
            sys.exit(1)
        metric_logger.update(loss=loss_value)
        last["loss_dict"] = {m: vals[i] for i, m in enumerate(names)}
        last["weighted"] = vals[T:2 * T]
        loss_value_reduce = vals[2 * T + 1] if world > 1 else loss_value        # :103, already reduced on the device
        if log_writer is not None and was_update:                               # :104-112 (rank 0 alone has a writer)
            log_writer.update(train_loss=loss_value_reduce, head="loss", step=step)

    def as_dict(data):
        if isinstance(data, dict):
            return data
        if getattr(args, "no_ffcv", True):
            return data[1]                                                       # engine_pretrain.py:50
        return {m: data[i] for i, m in enumerate(modalities)}                    # helpers.make_modality_dict

    if on_gpu:
        batches = DevicePrefetcher((as_dict(d) for d in data_loader), device)
    else:
        batches = ({k: v.to(device) for k, v in as_dict(d).items()} for d in data_loader)
    optimizer.zero_grad()
    log_var_list = None
    t0 = time.time()
    for data_iter_step, samples in enumerate(batches):
        if data_iter_step % update_freq == 0:                                   # per-iteration schedule, :53-56
            lr = cosine_lr(data_iter_step / n_iter + epoch, args.lr, args.min_lr, args.warmup_epochs, args.epochs)
            for group in optimizer.param_groups:
                group["lr"] = lr * group["lr_scale"] if "lr_scale" in group else lr
        out = model(samples, mask_ratio=args.mask_ratio)
        loss, log_var_list = out[0], out[4]
        update = (data_iter_step + 1) % update_freq == 0
        step_1000x = int((data_iter_step / n_iter + epoch) * 1000)              # epoch_1000x, :108
        pending.append((step_1000x, update))
        consume(reader.push(_with_rank_mean(_loss_vector(core, out), T, world)))
        loss_scaler(loss / update_freq if update_freq != 1 else loss, optimizer, parameters=model.parameters(),
                    update_grad=update)
        if update:
            optimizer.zero_grad()
        metric_logger.update(lr=optimizer.param_groups[0]["lr"])
        if log_writer is not None and update:
            log_writer.update(lr=optimizer.param_groups[0]["lr"], head="opt", step=step_1000x)
        if not quiet and (data_iter_step % print_freq == 0 or data_iter_step == n_iter - 1):
            print(f"{header} [{data_iter_step}/{n_iter}]  {metric_logger}  time: {(time.time() - t0) / (data_iter_step + 1):.4f}")
    for vals in reader.flush():
        consume(vals)
    metric_logger.synchronize_between_processes()
    if not quiet:
        print("Averaged stats:", metric_logger)
    normalized = None
    if last["weighted"] is not None and uncertainty == "uncertainty":
        import numpy as np
        normalized = np.asarray(last["weighted"], dtype=np.float32)
    return ({k: m.global_avg for k, m in metric_logger.meters.items()}, last["loss_dict"],
            list(log_var_list) if log_var_list is not None else None, normalized)


def _with_rank_mean(vec: torch.Tensor, T: int, world: int) -> torch.Tensor:
    """``helpers.all_reduce_mean`` (``helpers.py:393-401``) without its host sync: the total loss is all-reduced on the
    device and appended to the loss vector, so it reaches the host through the same lagged read.  The reference calls it
    unconditionally on every rank every step (``engine_pretrain.py:103``); so does this -- a collective that only the rank
    holding the log writer joined would pair with the other ranks' next gradient all-reduce."""
    if world == 1:
        return vec
    red = vec[2 * T:2 * T + 1].detach().clone()
    torch.distributed.all_reduce(red)
    return torch.cat([vec.detach().reshape(-1), red / world])


def fit(model, model_without_ddp, data_loader, optimizer, loss_scaler, device, args, log_writer=None, quiet: bool = True,
        lag: int = 2) -> List[dict]:
    """The epoch loop of ``main_pretrain.main`` (``main_pretrain.py:322-366``): resume from the newest checkpoint when
    ``args.auto_resume`` says so, then for every epoch from ``args.start_epoch`` run ``train_one_epoch`` and write a checkpoint
    every ``args.save_ckpt_freq`` epochs and after the last one.  Returns the per-epoch statistics (``train_<meter>`` keys plus
    ``epoch``, as the reference logs them)."""
    from . import checkpoint

    if not hasattr(args, "start_epoch"):
        args.start_epoch = 0
    if getattr(args, "output_dir", ""):
        checkpoint.auto_load_model(args, model, model_without_ddp, optimizer, loss_scaler)
    history = []
    for epoch in range(args.start_epoch, args.epochs):
        sampler = getattr(data_loader, "sampler", None)
        if hasattr(sampler, "set_epoch"):                                      # main_pretrain.py:337-338
            sampler.set_epoch(epoch)
        stats, loss_dict, log_vars, normalized = train_one_epoch(model, getattr(args, "modalities", None), data_loader, optimizer,
                                                                 device, epoch, False, loss_scaler, log_writer=log_writer,
                                                                 args=args, lag=lag, quiet=quiet)
        if getattr(args, "output_dir", "") and getattr(args, "save_ckpt", True):
            if (epoch + 1) % args.save_ckpt_freq == 0 or epoch + 1 == args.epochs:
                checkpoint.save_model(args, epoch, model, model_without_ddp, optimizer, loss_scaler)
        row = {f"train_{k}": v for k, v in stats.items()}
        row.update(epoch=epoch, loss_dict=loss_dict, log_vars=log_vars,
                   normalized=None if normalized is None else [float(v) for v in normalized])
        history.append(row)
    return history
