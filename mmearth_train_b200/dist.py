"""Gradient exchange of the data-parallel step.

The reference wraps the model in ``DistributedDataParallel`` (``main_pretrain.py:306-310``), whose reducer all-reduces
per-parameter buckets from autograd hooks.  Here every gradient lives in ONE flat fp32 buffer, the backward runs in
three parts in reverse layer order (``mpmae_backward_part``), and the slice of the buffer that a part has completed is
all-reduced (sum; the 1/world factor is folded into the backward seed) while the next part is still computing.
``torch.distributed`` runs each collective on the backend's own stream after the work already queued on the current
stream, so the overlap needs no extra stream management; NCCL over NVLink 5 / NVSwitch on the GPUs, gloo in the CPU
tests.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def world_size(group=None) -> int:
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


class FlatGradReducer:
    """Asynchronous all-reduce of contiguous slices of a flat gradient buffer."""

    def __init__(self, ranges: Sequence[Tuple[int, int]], total: int, group=None):
        ranges = [(int(lo), int(hi)) for lo, hi in ranges]
        covered = sorted(ranges)
        pos = 0
        for lo, hi in covered:                       # the parts must tile [0, total) exactly once
            if lo != pos or hi <= lo:
                raise ValueError(f"gradient ranges {ranges} do not partition [0, {total})")
            pos = hi
        if pos != total:
            raise ValueError(f"gradient ranges {ranges} do not partition [0, {total})")
        self.ranges, self.total, self.group = ranges, total, group
        self._pending: List = []
        self.bytes_reduced = 0

    def reduce_part(self, flat: torch.Tensor, part: int) -> None:
        lo, hi = self.ranges[part]
        if world_size(self.group) == 1:
            return
        self._pending.append(dist.all_reduce(flat[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        self.bytes_reduced += (hi - lo) * flat.element_size()

    def wait(self) -> None:
        for h in self._pending:
            h.wait()
        self._pending.clear()


def seed_scale(group=None) -> float:
    """Factor folded into the backward seed so that a SUM all-reduce yields DDP's mean gradient."""
    return 1.0 / world_size(group)
