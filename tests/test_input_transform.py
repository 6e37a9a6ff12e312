"""data.RawBatchTransform (stored dtypes in, model-ready batch out, on the device) against samples produced by the unmodified
reference loader transform MMEarthDataset.__getitem__ (fixture: oracle/make_dataset_golden.py)."""
import json
import os

import numpy as np
import pytest
import torch

from mmearth_train_b200.data import RawBatchTransform, _label_lut
from mmearth_train_b200 import synthetic as syn

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dataset_transform.npz")


def _load():
    z = np.load(GOLD)
    meta = json.loads(bytes(z["meta"]).decode())
    raw = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("raw.")}      # stored dtypes, torch.uint16 included
    want = {k[4:]: z[k] for k in z.files if k.startswith("out.")}
    return z, meta, raw, want


def _full_bands():
    full = {m: [f"{m}_{i}" for i in range(n)] for m, n in syn.FULL_BANDS.items()}
    # sentinel2 band NAMES matter (the 12 input bands are picked by name out of the 13 stored ones, MODALITIES.py)
    full["sentinel2"] = ["B1", "B2", "B3", "B4", "B5", "B6", "B7", "B8A", "B8", "B9", "B10", "B11", "B12"]
    return full


def test_batch_transform_equals_reference_loader_bit_for_bit():
    z, meta, raw, want = _load()
    full = _full_bands()
    from oracle import ref_harness
    if ref_harness.reference_available():                      # the stored band order is the reference's
        assert list(ref_harness.load_reference().MODALITIES.MODALITIES_FULL["sentinel2"]) == full["sentinel2"]
    tf = RawBatchTransform(meta["modalities"], full, meta["band_stats"])
    got = tf(raw, torch.from_numpy(z["l2a"]))
    assert list(got) == meta["order"]
    for m, w in want.items():
        g = got[m].numpy()
        assert g.dtype == w.dtype and g.shape == w.shape, m
        if w.dtype == np.int64:
            assert np.array_equal(g, w), m
        else:
            assert np.array_equal(np.isnan(g), np.isnan(w)), m
            assert np.array_equal(np.nan_to_num(g, nan=0.0), np.nan_to_num(w, nan=0.0)), m     # same float64 arithmetic
    # the no-data conventions the step relies on: NaN in continuous targets, -1 in class targets
    assert np.isnan(want["sentinel2"]).any() and (want["dynamic_world"] == -1).any() and (want["biome"] == -1).any()
    fast = RawBatchTransform(meta["modalities"], full, meta["band_stats"], exact=False)(raw, torch.from_numpy(z["l2a"]))
    for m, w in want.items():
        if w.dtype != np.int64:
            assert np.allclose(np.nan_to_num(fast[m].numpy()), np.nan_to_num(w), rtol=1e-5, atol=1e-5), m


def test_label_tables():
    dw, esa = _label_lut("dynamic_world"), _label_lut("esa_worldcover")
    assert torch.isnan(dw[0]) and dw[1:10].tolist() == list(range(9)) and torch.isnan(dw[10:]).all()
    assert torch.isnan(esa[0]) and esa[[20, 90, 95, 100]].tolist() == [1, 8, 9, 10]
    assert torch.isnan(esa[10])          # stored 10 -> class 0 -> caught by the no-data pass that follows (value 0): reference behaviour
    assert esa[7] == 7 and torch.isnan(esa[11]) and torch.isnan(esa[255])       # stored 1..9 pass through, like the reference


def test_unknown_modality_is_refused():
    with pytest.raises(ValueError):
        RawBatchTransform({"sentinel3": "all"}, {"sentinel3": ["a"]}, {})


def test_band_subsets_against_the_live_reference_loader():
    """Explicit band lists for several modalities (the fixture only has Sentinel-2's 12-of-13): the unmodified loader is run
    here on the same raw arrays (build container only)."""
    from oracle import ref_harness
    if not ref_harness.reference_available():
        pytest.skip("needs /root/reference")
    from oracle import make_dataset_golden as mdg
    mods = {"sentinel2": ["B4", "B3", "B2", "B8"], "sentinel1": ["desc_VV", "asc_VH"], "era5": ["year_avg_temp", "curr_month_total_precip"],
            "aster": ["slope"], "canopy_height_eth": "all", "esa_worldcover": "all", "dynamic_world": "all", "biome": "all",
            "lat": ["cos"], "month": "all"}
    raw, stats, l2a, mods, want, full = mdg.run_reference(mods)
    got = RawBatchTransform(mods, full, stats)({k: torch.from_numpy(v) for k, v in raw.items()}, torch.from_numpy(l2a))
    assert list(got) == list(want)
    for m, w in want.items():
        g = got[m].numpy()
        assert g.dtype == w.dtype and g.shape == w.shape, m
        assert np.array_equal(np.isnan(g.astype(np.float64)), np.isnan(w.astype(np.float64))), m
        assert np.array_equal(np.nan_to_num(g.astype(np.float64)), np.nan_to_num(w.astype(np.float64))), m
