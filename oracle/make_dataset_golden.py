"""Golden samples of the UNMODIFIED reference dataset transform (TEST INFRASTRUCTURE ONLY; runs in the build container).

``MMEarthDataset.__getitem__`` (``/root/reference/mmearth_dataset.py:58-153``) is run on synthetic RAW arrays (the dtypes and
value ranges the HDF5 files store: 16-bit Sentinel-2 digital numbers with 0 = no data, float bands with -inf = no data, one
byte per pixel for the label maps, one-hot rows with 255 / 65535 = no data) through a stand-in for the ``h5py.File`` handle
(h5py is not installed; the class only indexes the handle like a dict of arrays).  Raw inputs, band statistics and the
transformed samples go to ``tests/golden/dataset_transform.npz`` for ``tests/test_input_transform.py``.

    python -m oracle.make_dataset_golden
"""
from __future__ import annotations

import importlib
import json
import os
import sys
import types
from argparse import Namespace

import numpy as np

from . import ref_harness

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "dataset_transform.npz")
N, S = 6, 16            # samples, raster side


def raw_arrays(seed: int = 2024):
    r = np.random.default_rng(seed)
    raw = {}
    s2 = r.integers(1, 12000, size=(N, 13, S, S)).astype(np.uint16)
    s2[r.random(s2.shape) < 0.03] = 0                                                   # no data
    raw["sentinel2"] = s2
    for m, c in (("sentinel1", 8), ("aster", 2)):
        a = r.normal(size=(N, c, S, S)).astype(np.float32) * 7 - 3
        a[r.random(a.shape) < 0.04] = -np.inf
        raw[m] = a
    ch = r.integers(0, 60, size=(N, 2, S, S)).astype(np.uint8)
    ch[r.random(ch.shape) < 0.05] = 255
    raw["canopy_height_eth"] = ch
    raw["dynamic_world"] = r.integers(0, 12, size=(N, 1, S, S)).astype(np.uint8)        # 0 = no data, 10 / 11 invalid
    raw["esa_worldcover"] = r.choice(np.array([0, 10, 20, 30, 40, 50, 60, 70, 80, 90, 95, 100, 255, 7], dtype=np.uint8),
                                     size=(N, 1, S, S))
    for m, c in (("lat", 2), ("lon", 2), ("month", 2)):
        a = r.uniform(-1, 1, size=(N, c)).astype(np.float32)
        raw[m] = a
    raw["lat"][1, :] = -np.inf
    e = r.normal(size=(N, 12)).astype(np.float32) * 30 + 280
    e[2, 3] = np.nan
    raw["era5"] = e
    bi = np.eye(14, dtype=np.uint8)[r.integers(0, 14, size=N)]
    bi[3] = 255                                                                          # a sample without a biome
    raw["biome"] = bi
    eco = np.eye(846, dtype=np.uint16)[r.integers(0, 846, size=N)]
    eco[4] = 65535
    raw["eco_region"] = eco
    return raw


def band_stats(seed: int = 7):
    r = np.random.default_rng(seed)
    full = {"sentinel2_l1c": 13, "sentinel2_l2a": 13, "sentinel1": 8, "aster": 2, "canopy_height_eth": 2, "lat": 2, "lon": 2,
            "month": 2, "era5": 12}
    return {k: {"mean": r.uniform(-5, 2000, size=n).tolist(), "std": r.uniform(0.5, 900, size=n).tolist()} for k, n in full.items()}


def run_reference(mods=None):
    """(raw arrays, band statistics, l2a flags, modality dict, {modality: stacked reference outputs}) for the given
    ``modalities`` dict (default: the reference's INP + OUT modalities)."""
    _stub = types.ModuleType("h5py")
    _stub.File = object
    sys.modules.setdefault("h5py", _stub)
    ref = ref_harness.load_reference()
    ds_mod = importlib.import_module(f"{ref.pkg}.mmearth_dataset")
    M = ref.MODALITIES
    raw, stats = raw_arrays(), band_stats()
    names = [f"tile_{i}" for i in range(N)]
    tile_info = {n: {"S2_type": "l2a" if i % 2 else "l1c"} for i, n in enumerate(names)}
    splits = os.path.join("/tmp", "mpmae_ds_splits.json")
    json.dump({"train": list(range(N))}, open(splits, "w"))
    if mods is None:
        mods = dict(M.INP_MODALITIES)
        mods.update(M.OUT_MODALITIES)
    args = Namespace(data_path=None, data_name="synthetic", splits_path=splits, tile_info=tile_info, modalities=mods,
                     modalities_full=dict(M.MODALITIES_FULL), band_stats=stats)
    ds = ds_mod.MMEarthDataset(args, split="train")
    handle = dict(raw)
    handle["metadata"] = np.array([(n.encode(),) for n in names], dtype=[("name", "S16")])
    ds.data_full = handle                                   # what _open_hdf5 would have produced
    out = {}
    for i in range(N):
        sample = ds[i]
        assert sample["id"] == names[i]
        for m, v in sample.items():
            if m != "id":
                out.setdefault(m, []).append(np.asarray(v))
    l2a = np.array([tile_info[n]["S2_type"] == "l2a" for n in names])
    return raw, stats, l2a, mods, {k: np.stack(v) for k, v in out.items()}, dict(M.MODALITIES_FULL)


def main():
    raw, stats, l2a, mods, out, _full = run_reference()
    arrays = {f"raw.{k}": v for k, v in raw.items()}
    arrays.update({f"out.{k}": v for k, v in out.items()})
    arrays["l2a"] = l2a
    arrays["meta"] = np.frombuffer(json.dumps({"band_stats": stats, "modalities": {k: v for k, v in mods.items()},
                                               "order": list(out)}).encode(), dtype=np.uint8)
    np.savez_compressed(OUT, **arrays)
    print({k: (v.dtype, v.shape) for k, v in arrays.items() if k.startswith("out.")})


if __name__ == "__main__":
    main()
