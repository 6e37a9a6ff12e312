"""Input side of the pretraining step: pinned-host -> device prefetch on a copy stream.

The reference feeds ``train_one_epoch`` from a ``DataLoader(pin_memory=True)`` and copies every modality with
``.to(device, non_blocking=True)`` inside the loop (``engine_pretrain.py:50-61``), i.e. on the compute stream, so the
91.7 MB of a 256-sample batch (12 modalities, int64 label maps included) serialises with the step.  ``DevicePrefetcher``
is the drop-in wrapper around any iterable of sample dicts: the copy of batch i+1 into one of two persistent device
buffer sets runs on its own stream while batch i is being computed; events order buffer reuse, nothing is allocated per
step.  (the step before the hot path.)
"""
from __future__ import annotations

from typing import Dict, Iterable, Iterator, Optional

import torch


class DevicePrefetcher:
    """Iterates device-resident batches; host tensors should be pinned for the copies to be asynchronous.

    ``depth`` buffer sets rotate; a batch handed out stays valid until ``depth - 1`` further batches have been requested
    (the training loop consumes a batch within its own iteration).
    """

    def __init__(self, batches: Iterable[Dict[str, torch.Tensor]], device: torch.device, depth: int = 2):
        if depth < 2:
            raise ValueError("depth must be >= 2")
        self.src = batches
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("DevicePrefetcher copies to a CUDA device")
        self.depth = depth
        self.copy_stream = torch.cuda.Stream(self.device)
        self._bufs: list = [None] * depth
        self._ready = [torch.cuda.Event() for _ in range(depth)]
        self._free = [torch.cuda.Event() for _ in range(depth)]
        self.bytes_copied = 0

    def _stage(self, slot: int, host: Dict[str, torch.Tensor]) -> None:
        bufs = self._bufs[slot]
        if bufs is None or any(k not in bufs or bufs[k].shape != v.shape or bufs[k].dtype != v.dtype for k, v in host.items()):
            bufs = {k: torch.empty(v.shape, dtype=v.dtype, device=self.device) for k, v in host.items()}
            self._bufs[slot] = bufs
        else:
            self.copy_stream.wait_event(self._free[slot])        # the step that used this slot has been queued behind us
        with torch.cuda.stream(self.copy_stream):
            for k, v in host.items():
                bufs[k].copy_(v, non_blocking=True)
                self.bytes_copied += v.numel() * v.element_size()
            self._ready[slot].record(self.copy_stream)

    def __iter__(self) -> Iterator[Dict[str, torch.Tensor]]:
        it = iter(self.src)
        slot = 0
        nxt: Optional[Dict[str, torch.Tensor]] = next(it, None)
        if nxt is None:
            return
        self._stage(slot, nxt)
        while True:
            cur_slot = slot
            nxt = next(it, None)
            if nxt is not None:
                slot = (slot + 1) % self.depth
                self._stage(slot, nxt)
            compute = torch.cuda.current_stream(self.device)
            compute.wait_event(self._ready[cur_slot])
            yield self._bufs[cur_slot]
            self._free[cur_slot].record(torch.cuda.current_stream(self.device))   # everything that read the batch is queued
            if nxt is None:
                return


class LossReader:
    """Per-step device -> host read of a scalar without stalling the launch queue.

    The reference calls ``loss.item()`` right after the forward pass of every iteration (``engine_pretrain.py:72``), which
    drains the GPU before the next step can even be queued.  ``push(loss)`` copies the 0-d tensor into one of ``depth`` pinned
    host slots (non-blocking, event-tracked) and returns the value pushed ``depth`` calls earlier, whose copy has had at
    least a whole step to land; ``flush()`` returns the values still in flight.  Every step's result is still read on the
    host, ``depth`` steps late (engine-loop hygiene).
    """

    def __init__(self, device: torch.device, depth: int = 2, width: int = 1):
        if depth < 1 or width < 1:
            raise ValueError("depth and width must be >= 1")
        self.device = torch.device(device)
        self.depth, self.width = depth, width
        self._host = [torch.zeros(width, dtype=torch.float32).pin_memory() for _ in range(depth)]
        self._ev = [torch.cuda.Event() for _ in range(depth)]
        self._busy = [False] * depth
        self._n = 0
        self.bytes_read = 0

    def _take(self, slot: int):
        self._ev[slot].synchronize()
        self._busy[slot] = False
        if self.width == 1:
            return float(self._host[slot][0])
        return self._host[slot].tolist()

    def push(self, value: torch.Tensor):
        """``value``: 0-d tensor (``width`` 1, returns floats) or a ``[width]`` vector (returns lists of floats)."""
        slot = self._n % self.depth
        out = self._take(slot) if self._busy[slot] else None
        self._host[slot].copy_(value.detach().reshape(self.width), non_blocking=True)
        self._ev[slot].record(torch.cuda.current_stream(self.device))
        self._busy[slot] = True
        self._n += 1
        self.bytes_read += 4 * self.width
        return out

    def flush(self) -> list:
        out = []
        for k in range(self.depth):
            slot = (self._n + k) % self.depth
            if self._busy[slot]:
                out.append(self._take(slot))
        return out
