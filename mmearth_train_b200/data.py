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
        compute = torch.cuda.current_stream(self.device)
        if bufs is None or any(k not in bufs or bufs[k].shape != v.shape or bufs[k].dtype != v.dtype for k, v in host.items()):
            # (re)allocation: the caching allocator may hand back a block whose last use is still queued on the compute
            # stream (the buffers of the previous epoch / of a differently shaped batch), so the copy stream first waits
            # for everything queued there; allocating under the copy stream makes it the blocks' owning stream, and
            # record_stream tells the allocator that the compute stream reads them too
            self.copy_stream.wait_stream(compute)
            with torch.cuda.stream(self.copy_stream):
                bufs = {k: torch.empty(v.shape, dtype=v.dtype, device=self.device) for k, v in host.items()}
            for t in bufs.values():
                t.record_stream(compute)
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


# ---------------------------------------------------------------------------------------------------------------------------
# Stored-dtype batches -> model-ready batches, on the device.
#: value that marks "no data" in the stored arrays (MODALITIES.py:37-53)
NO_DATA_VAL = {"sentinel2": 0, "sentinel1": float("-inf"), "aster": float("-inf"), "canopy_height_eth": 255, "dynamic_world": 0,
               "esa_worldcover": 0, "lat": float("-inf"), "lon": float("-inf"), "month": float("-inf"), "era5": float("nan"),
               "biome": 255, "eco_region": 65535}
#: modalities whose targets are class indices (MODALITIES.py:163-180: "segmentation" / "classification")
CLASS_TARGETS = ("esa_worldcover", "dynamic_world", "biome", "eco_region")
_NOT_NORMALISED = ("biome", "eco_region", "dynamic_world", "esa_worldcover")


def _label_lut(modality: str) -> torch.Tensor:
    """256-entry table stored label -> class index (NaN = ignore), built by replaying the loader's rule on every byte value:
    dynamic_world 0 -> no data, 1..9 -> 0..8, anything else -> ignore (``mmearth_dataset.py:88-97``); esa_worldcover 0 -> no
    data, 10, 20, .., 90, 95, 100 -> 0..10 applied ONE AFTER THE OTHER, then anything above 10 -> ignore (``:99-108``; stored
    values 1..9 therefore pass through as classes, exactly as in the reference).  See the note on class 0 below."""
    v = torch.arange(256, dtype=torch.float64)
    nan = torch.full_like(v, float("nan"))
    v = torch.where(v == NO_DATA_VAL[modality], nan, v)
    if modality == "dynamic_world":
        olds, news, top = range(1, 10), range(0, 9), 8
    else:
        olds, news, top = (10, 20, 30, 40, 50, 60, 70, 80, 90, 95, 100), range(0, 11), 10
    for old, new in zip(olds, news):
        v = torch.where(v == old, torch.full_like(v, float(new)), v)
    v = torch.where(v > top, nan, v)
    if modality != "dynamic_world":
        # the loader's general no-data pass runs AFTER the remap (``:110-115``) and esa_worldcover's no-data value is 0: the
        # freshly remapped class 0 (stored 10, tree cover) becomes "ignore" as well.  Kept, because it is what the reference trains on.
        v = torch.where(v == NO_DATA_VAL[modality], nan, v)
    return v


class RawBatchTransform:
    """``MMEarthDataset.__getitem__`` (``mmearth_dataset.py:58-153``) for a whole batch, on whatever device the tensors are on.

    The reference widens every sample to float32 / int64 in the loader workers (band selection, label remap, no-data -> NaN,
    per-band z-scoring, NaN -> -1 for class targets), so a 256-sample batch crosses PCIe as 91.7 MB.  Here the loader hands
    over the arrays as they are STORED (uint16 Sentinel-2, uint8 label maps and canopy height, float32 for the rest), they are
    copied in that form (``DevicePrefetcher``) and this transform runs after the copy with a handful of elementwise torch
    kernels.  Same arithmetic as the reference (float64 subtraction / division, then float32) unless ``exact=False``.

    ``modalities`` / ``modalities_full`` / ``band_stats`` are the ``args`` fields of the same names
    (``mmearth_dataset.py:36-50``); ``l2a`` says per sample which Sentinel-2 statistics apply (``tile_info[..]["S2_type"]``).
    """

    def __init__(self, modalities: Dict[str, object], modalities_full: Dict[str, list], band_stats: Dict[str, dict],
                 exact: bool = True):
        self.modalities, self.exact = dict(modalities), exact
        self._idx, self._stats, self._lut, self._stats_host = {}, {}, {}, {}
        for m, bands in self.modalities.items():
            if m not in NO_DATA_VAL:
                raise ValueError(f"unknown modality {m!r}")
            full = list(modalities_full[m])
            idx = list(range(len(full))) if bands == "all" else [full.index(b) for b in bands]
            self._idx[m] = torch.tensor(idx, dtype=torch.long)
            if m not in _NOT_NORMALISED:
                keys = ("sentinel2_l1c", "sentinel2_l2a") if m == "sentinel2" else (m, m)
                self._stats[m] = torch.stack([torch.tensor([band_stats[k]["mean"], band_stats[k]["std"]],
                                                           dtype=torch.float64)[:, idx] for k in keys])      # [2 (l1c/l2a), 2, nb]
                self._stats_host[m] = self._stats[m].clone()
            if m in ("dynamic_world", "esa_worldcover"):
                self._lut[m] = _label_lut(m)

    def _on(self, cache: dict, m: str, device) -> torch.Tensor:
        t = cache[m]
        if t.device != device:
            t = cache[m] = t.to(device)
        return t

    # ---- device path: one fused kernel per modality (csrc/raw_transform.cuh) through the C ABI
    _SRC_TYPE = {torch.uint8: 0, torch.uint16: 1, torch.float32: 2}

    def _native(self, raw: Dict[str, torch.Tensor], l2a: torch.Tensor, into: Optional[Dict[str, torch.Tensor]] = None) -> Dict[str, torch.Tensor]:
        import ctypes as C

        from . import _native as nat
        out = {}
        dev = next(iter(raw.values())).device
        flags = l2a.to(device=dev, dtype=torch.uint8).contiguous()
        stream = torch.cuda.current_stream(dev).cuda_stream
        with torch.cuda.device(dev):
            for m in self.modalities:
                x = raw[m].contiguous()
                if x.dtype not in self._SRC_TYPE:
                    raise TypeError(f"{m}: stored dtype {x.dtype} (expected uint8, uint16 or float32)")
                B, src_bands = x.shape[0], x.shape[1]
                inner = 1
                for d in x.shape[2:]:
                    inner *= d
                idx = self._idx[m].tolist()
                whole = m in ("biome", "eco_region")                   # one-hot rows are taken whole
                n_bands = src_bands if whole else len(idx)
                is_class = m in CLASS_TARGETS
                shape, dtype = (B, n_bands) + tuple(x.shape[2:]), torch.int64 if is_class else torch.float32
                o = into[m] if into is not None else torch.empty(shape, dtype=dtype, device=dev)
                if tuple(o.shape) != shape or o.dtype != dtype or not o.is_contiguous() or o.device != dev:
                    raise ValueError(f"{m}: output buffer {tuple(o.shape)} {o.dtype}, expected contiguous {shape} {dtype} on {dev}")
                d = nat.RawDesc()
                d.src, d.out, d.l2a = x.data_ptr(), o.data_ptr(), flags.data_ptr()
                d.lut = self._on(self._lut, m, dev).data_ptr() if m in self._lut else None
                d.inner, d.B, d.src_bands, d.n_bands = inner, B, src_bands, n_bands
                d.src_type, d.out_int64 = self._SRC_TYPE[x.dtype], 1 if is_class else 0
                nd = NO_DATA_VAL[m]
                d.has_nodata, d.nodata = (1, float(nd)) if nd == nd else (0, 0.0)
                d.normalize = 1 if m in self._stats else 0
                if n_bands <= nat.RAW_MAX_BANDS:
                    for b in range(n_bands):
                        d.band[b] = b if whole else idx[b]
                    if m in self._stats:
                        st = self._stats_host[m]                           # [2 (l1c / l2a), 2 (mean / std), nb] float64, host
                        for s_ in range(2):
                            for b in range(n_bands):
                                d.mean[s_][b], d.std[s_][b] = float(st[s_, 0, b]), float(st[s_, 1, b])
                nat.check(nat.lib.mpmae_raw_transform(C.byref(d), C.c_void_p(stream)), "mpmae_raw_transform")
                out[m] = o
        return out

    def __call__(self, raw: Dict[str, torch.Tensor], l2a: torch.Tensor,
                 into: Optional[Dict[str, torch.Tensor]] = None) -> Dict[str, torch.Tensor]:
        """``into`` (device path only): write the model-ready tensors into these buffers, e.g. the static inputs of a
        ``GraphedStep``, instead of allocating new ones."""
        if self.exact and all(v.is_cuda for v in raw.values()):
            return self._native(raw, l2a, into)
        if into is not None:
            raise ValueError("`into` needs CUDA inputs and exact=True (the fused device kernel)")
        out = {}
        for m in self.modalities:
            x = raw[m]
            dev = x.device
            if m in self._lut:                                   # label maps: one gather through the 256-entry table
                x = x.index_select(1, self._on(self._idx, m, dev))
                v = self._on(self._lut, m, dev)[x.long()]
            else:
                v = x.double() if self.exact else x.float()      # widen first: torch.uint16 supports little besides conversion
                if m not in ("biome", "eco_region"):             # one-hot rows are taken whole (mmearth_dataset.py:76-79)
                    v = v.index_select(1, self._on(self._idx, m, dev))
                nd = NO_DATA_VAL[m]
                if nd == nd:                                     # NaN marks nothing: `data == nan` is never true
                    v = torch.where(v == nd, torch.full_like(v, float("nan")), v)
            if m in self._stats:
                st = self._on(self._stats, m, dev)[l2a.to(dev).long()]                  # [B, 2, nb]
                shape = st.shape[:1] + st.shape[2:] + (1,) * (v.dim() - 2)
                mean, std = st[:, 0].reshape(shape), st[:, 1].reshape(shape)
                if not self.exact:
                    mean, std = mean.float(), std.float()
                v = (v - mean) / std
            if m in CLASS_TARGETS:
                out[m] = torch.where(torch.isnan(v), torch.full_like(v, -1.0), v).long()
            else:
                out[m] = v.float()
        return out
