"""Synthetic MMEarth batches and the ``args`` namespace of a pretraining run.

What ``bench.py`` and the profiling tools feed the step with: tensors with the dtypes, shapes and value ranges of
``mmearth_dataset.py:58-153`` (z-scored float bands, int64 label maps with -1 = ignore, one-hot int64 image labels), and the
fields ``main_pretrain.py:175-180`` derives from ``MODALITIES.py:75-161`` and puts on ``args`` for ``FCMAE`` to read.  The test
oracle has its own generator; ``tests/test_host_logic.py`` holds the two to bit-identical output for the same seed.
"""
from __future__ import annotations

from argparse import Namespace
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

#: sentinel2 input bands used for pretraining (MODALITIES.py: INP_MODALITIES) -- 12 of the 13 stored bands
S2_BANDS = ["B1", "B2", "B3", "B4", "B5", "B6", "B7", "B8A", "B8", "B9", "B11", "B12"]
#: bands stored per modality (MODALITIES.py: MODALITIES_FULL)
FULL_BANDS = {"sentinel2": 13, "sentinel1": 8, "aster": 2, "era5": 12, "dynamic_world": 1, "canopy_height_eth": 2, "lat": 2,
              "lon": 2, "biome": 1, "eco_region": 1, "month": 2, "esa_worldcover": 1}
#: output modalities in the reference's dict order (MODALITIES.py:75-101): the order of loss_dict and of log_vars
ALL_OUT = ["sentinel2", "sentinel1", "aster", "era5", "dynamic_world", "canopy_height_eth", "lat", "lon", "biome",
           "eco_region", "month", "esa_worldcover"]


def make_args(out_modalities: Optional[List[str]] = None, loss_aggr: str = "uncertainty") -> Namespace:
    """``args`` as ``FCMAE.__init__`` reads it (``models/fcmae.py:44-91``): S2 in, the listed modalities (default all 12) out."""
    outs = ALL_OUT if out_modalities is None else list(out_modalities)
    out = {m: (S2_BANDS if m == "sentinel2" else "all") for m in outs}
    mods = {"sentinel2": S2_BANDS}
    mods.update(out)
    full = {m: [f"{m}_{i}" for i in range(n)] for m, n in FULL_BANDS.items()}
    return Namespace(inp_modalities={"sentinel2": S2_BANDS}, out_modalities=out, modalities=mods, modalities_full=full,
                     use_orig_stem=False, loss_aggr=loss_aggr)


def synthetic_batch(B: int, img_size: int, out_modalities: Optional[List[str]] = None, seed: int = 1234,
                    nan_frac: float = 0.0) -> Dict[str, torch.Tensor]:
    """One host batch at model size (the random crop is then the identity).  ``nan_frac`` > 0 puts that fraction of NaNs
    into the continuous pixel targets (and half of it into era5), the no-data the loss has to skip (``fcmae.py:384-402``)."""
    g = torch.Generator().manual_seed(seed)
    outs = ALL_OUT if out_modalities is None else list(out_modalities)
    S = img_size
    d = {"sentinel2": torch.randn(B, 12, S, S, generator=g)}
    for m, c in (("sentinel1", 8), ("aster", 2), ("canopy_height_eth", 2)):
        if m in outs:
            t = torch.randn(B, c, S, S, generator=g)
            if nan_frac > 0:
                t[torch.rand(t.shape, generator=g) < nan_frac] = float("nan")
            d[m] = t
    if "dynamic_world" in outs:
        d["dynamic_world"] = torch.randint(-1, 9, (B, 1, S, S), generator=g)
    if "esa_worldcover" in outs:
        d["esa_worldcover"] = torch.randint(-1, 11, (B, 1, S, S), generator=g)
    if "biome" in outs:
        d["biome"] = F.one_hot(torch.randint(0, 14, (B,), generator=g), 14)
    if "eco_region" in outs:
        d["eco_region"] = F.one_hot(torch.randint(0, 846, (B,), generator=g), 846)
    for m, c in (("lat", 2), ("lon", 2), ("month", 2), ("era5", 12)):
        if m in outs:
            t = torch.randn(B, c, generator=g)
            if m == "era5" and nan_frac > 0:
                t[torch.rand(t.shape, generator=g) < nan_frac / 2] = float("nan")
            d[m] = t
    return d


# ------------------------------------------------------------------------------------------------ stored-dtype batches
#: Sentinel-2 band names as stored (MODALITIES.py: MODALITIES_FULL): 13 bands, of which pretraining reads the 12 in S2_BANDS
S2_FULL = ["B1", "B2", "B3", "B4", "B5", "B6", "B7", "B8A", "B8", "B9", "B10", "B11", "B12"]


def raw_modalities_full() -> Dict[str, List[str]]:
    full = {m: [f"{m}_{i}" for i in range(n)] for m, n in FULL_BANDS.items()}
    full["sentinel2"] = list(S2_FULL)
    return full


def synthetic_band_stats() -> Dict[str, dict]:
    """Per-band mean / std in the layout of the reference's ``band_stats`` json (``mmearth_dataset.py:36-50``), matched to
    the distributions of :func:`synthetic_raw_batch` so that the z-scored bands come out ~N(0, 1) like real data."""
    spec = {"sentinel2_l1c": (13, 6000.0, 3464.0), "sentinel2_l2a": (13, 5900.0, 3400.0), "sentinel1": (8, -3.0, 7.0),
            "aster": (2, -3.0, 7.0), "canopy_height_eth": (2, 29.5, 17.3), "lat": (2, 0.0, 0.577), "lon": (2, 0.0, 0.577),
            "month": (2, 0.0, 0.577), "era5": (12, 280.0, 30.0)}
    return {k: {"mean": [m + 0.01 * i for i in range(n)], "std": [sd * (1.0 + 0.01 * i) for i in range(n)]}
            for k, (n, m, sd) in spec.items()}


def synthetic_raw_batch(B: int, img_size: int, seed: int = 1234) -> Dict[str, torch.Tensor]:
    """One host batch AS STORED in the MMEarth HDF5 files (``mmearth_dataset.py:58-153`` reads these and widens them):
    16-bit Sentinel-2 digital numbers (13 bands, 0 = no data), float32 Sentinel-1 / ASTER / ERA5 / lat / lon / month, one byte per
    pixel for canopy height and the two label maps, one-hot uint8 / uint16 rows for biome / eco-region.  56.7 MB for 256 samples
    of 56 x 56 against 91.7 MB for the widened float32 / int64 tensors the reference loader hands to the training loop."""
    g = torch.Generator().manual_seed(seed)
    S = img_size
    d = {"sentinel2": torch.randint(1, 12000, (B, 13, S, S), generator=g, dtype=torch.int32).to(torch.uint16),
         "sentinel1": torch.randn(B, 8, S, S, generator=g) * 7 - 3,
         "aster": torch.randn(B, 2, S, S, generator=g) * 7 - 3,
         "canopy_height_eth": torch.randint(0, 60, (B, 2, S, S), generator=g, dtype=torch.int32).to(torch.uint8),
         "dynamic_world": torch.randint(0, 10, (B, 1, S, S), generator=g, dtype=torch.int32).to(torch.uint8),
         "esa_worldcover": (torch.randint(0, 10, (B, 1, S, S), generator=g, dtype=torch.int32) * 10).to(torch.uint8),
         "era5": torch.randn(B, 12, generator=g) * 30 + 280}
    for m in ("lat", "lon", "month"):
        d[m] = torch.rand(B, 2, generator=g) * 2 - 1
    d["biome"] = F.one_hot(torch.randint(0, 14, (B,), generator=g), 14).to(torch.uint8)
    d["eco_region"] = F.one_hot(torch.randint(0, 846, (B,), generator=g), 846).to(torch.int32).to(torch.uint16)
    return d
