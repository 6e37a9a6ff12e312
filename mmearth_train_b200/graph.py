"""One training iteration as a CUDA graph (``engine_pretrain.py:44-113`` is the loop it serves).

The native step has fixed shapes -- every sample keeps exactly ``V`` visible patches, so the launch sequence, the grids and
every pointer are the same on every iteration -- which makes the whole iteration (mask, encoder, decoder, losses, hand-derived
backward, fused AdamW: ~185 kernel launches) capturable once and replayable with a single ``cudaGraphLaunch``.  At small
per-GPU batches the iteration is bound by launch latency, not by the kernels: BASELINE.json ``configs[0]`` (atto, bs 8) drops
from ~3 ms to well under 1 ms per iteration.

What is captured: ``model(static_batch, mask_ratio)`` (the noise is drawn inside the graph by torch's graph-safe generator),
``loss.backward()``, ``optimizer.step_dev`` with the learning rate and the step number in device memory, and the clearing of
the gradient buffer; under ``torch.distributed`` also the NCCL all-reduces of the flat gradient buffer, overlapped with the
backward parts exactly as in the eager step.  What stays outside: copying the next batch into the static input buffers and
writing the learning rate of the iteration (one scalar).  Gradient accumulation (``update_freq`` > 1) is not captured: use
the eager path (``engine.train_one_epoch``) for it.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch


class GraphedStep:
    """``loss = step(batch, lr)``: replays forward + backward + AdamW for one batch of the captured shapes.

    ``example_batch`` fixes shapes and dtypes (device tensors).  The returned loss, ``step.losses`` (the ``2T + 1`` vector
    ``[per-modality, weighted, total]``), ``step.mask`` and ``step.pred`` are STATIC tensors, overwritten by the next call.
    """

    def __init__(self, model, optimizer, example_batch: Dict[str, torch.Tensor], mask_ratio: float = 0.6, warmup: int = 3):
        # Under torch.distributed the part-wise NCCL all-reduces of the flat gradient buffer are captured with the kernels
        # (NCCL >= 2.9.6 supports stream capture; torch joins its communication stream back into the capturing stream when
        # the work handles are waited on), so every rank must build and replay its GraphedStep in lockstep.
        if not hasattr(optimizer, "step_dev"):
            raise TypeError("GraphedStep needs FlatAdamW (the step number and learning rate live in device memory)")
        self.model, self.optimizer, self.mask_ratio = model, optimizer, mask_ratio
        dev = model.flat_params.device
        if dev.type != "cuda":
            raise RuntimeError("move the model to the GPU first")
        self.static = {k: torch.empty_like(v, device=dev) for k, v in example_batch.items()}
        for k, v in example_batch.items():
            self.static[k].copy_(v)
        st = optimizer._ensure_dev_state()
        # ---- warm-up on a side stream (lazy allocations, kernel attributes, job-table uploads), then put everything back
        snap = (model.flat_params.clone(), optimizer.exp_avg.clone(), optimizer.exp_avg_sq.clone(), st.clone())
        rng = torch.cuda.get_rng_state(dev)
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):
                self._iteration()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        model.flat_params.copy_(snap[0]); optimizer.exp_avg.copy_(snap[1]); optimizer.exp_avg_sq.copy_(snap[2]); st.copy_(snap[3])
        torch.cuda.set_rng_state(rng, dev)
        # ---- capture
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            out = self._iteration()
        self.loss, self.pred, self.mask = out[0].detach(), out[1], out[2]
        self.losses = model.last_run["losses"]
        self.loss_dict = {m: self.losses[i] for i, m in enumerate(model.out_modalities)}
        self.replays = 0

    def close(self) -> None:
        """Destroys the CUDA graph.  Under ``torch.distributed`` call this (or drop every reference to the object) BEFORE
        ``destroy_process_group()``: NCCL's communicator teardown waits for every graph that captured its collectives."""
        graph, self.graph = getattr(self, "graph", None), None
        if graph is not None:
            torch.cuda.synchronize()
            graph.reset()
            del graph

    def _iteration(self):
        model, opt = self.model, self.optimizer
        out = model(self.static, mask_ratio=self.mask_ratio)
        out[0].backward()
        opt.step_dev(lr_on_device=True)
        opt.zero_grad(set_to_none=True)
        return out

    def __call__(self, batch: Optional[Dict[str, torch.Tensor]] = None, lr: Optional[float] = None) -> torch.Tensor:
        if batch is not None:
            for k, dst in self.static.items():
                src = batch[k]
                if src.shape != dst.shape or src.dtype != dst.dtype:
                    raise ValueError(f"{k}: batch {tuple(src.shape)} {src.dtype} does not match the captured "
                                     f"{tuple(dst.shape)} {dst.dtype}; capture another GraphedStep for another shape")
                if src.data_ptr() != dst.data_ptr():
                    dst.copy_(src, non_blocking=True)
        if lr is not None:
            self.optimizer.lr = lr
        if self.graph is None:
            raise RuntimeError("this GraphedStep was closed")
        self.optimizer._dev_state[3] = float(self.optimizer.lr)
        self.graph.replay()
        self.replays += 1
        return self.loss
