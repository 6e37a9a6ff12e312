"""DevicePrefetcher: batches arrive on the device in order, bit-identical, from rotating persistent buffers."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_prefetcher_order_and_values():
    from mmearth_train_b200.data import DevicePrefetcher
    g = torch.Generator().manual_seed(0)
    host = [{"a": torch.randn(64, 12, 8, 8, generator=g).pin_memory(),
             "b": torch.randint(-1, 9, (64, 1, 8, 8), generator=g).pin_memory()} for _ in range(5)]
    dev = torch.device("cuda", 0)
    pf = DevicePrefetcher(iter(host), dev, depth=2)
    seen = 0
    ptrs = set()
    for i, b in enumerate(pf):
        # consume on the compute stream before the slot can be reused
        assert torch.equal(b["a"].cpu(), host[i]["a"]) and torch.equal(b["b"].cpu(), host[i]["b"])
        assert b["a"].device == dev and b["b"].dtype == torch.int64
        ptrs.add(b["a"].data_ptr())
        seen += 1
    assert seen == 5 and len(ptrs) == 2                 # two persistent buffer sets, no per-step allocation
    assert pf.bytes_copied == sum(v.numel() * v.element_size() for h in host for v in h.values())


def test_prefetcher_rejects_cpu_target():
    from mmearth_train_b200.data import DevicePrefetcher
    with pytest.raises(ValueError):
        DevicePrefetcher(iter([]), torch.device("cpu"))


def test_loss_reader_returns_every_value_one_step_late():
    from mmearth_train_b200.data import LossReader
    dev = torch.device("cuda", 0)
    r = LossReader(dev, depth=2)
    got = []
    for i in range(5):
        v = r.push(torch.tensor(float(i) + 0.5, device=dev) * 2.0)
        if v is not None:
            got.append(v)
    got += r.flush()
    assert got == [1.0, 3.0, 5.0, 7.0, 9.0] and r.bytes_read == 20


def test_raw_batch_transform_on_cuda_is_bit_exact_against_reference_loader():
    """data.RawBatchTransform ON THE DEVICE (stored uint16 / uint8 / float32 arrays copied as they are,
    transformed after the copy) against samples of the unmodified ``MMEarthDataset.__getitem__``
    (tests/golden/dataset_transform.npz, oracle/make_dataset_golden.py) -- bit for bit, as on the CPU."""
    import numpy as np
    from mmearth_train_b200.data import DevicePrefetcher, RawBatchTransform
    from tests.test_input_transform import _full_bands, _load
    z, meta, raw, want = _load()
    dev = torch.device("cuda", 0)
    tf = RawBatchTransform(meta["modalities"], _full_bands(), meta["band_stats"])
    l2a = torch.from_numpy(z["l2a"])
    host = {k: v.pin_memory() for k, v in raw.items()}
    for b in DevicePrefetcher(iter([host]), dev):            # stored dtypes cross PCIe, the transform runs behind the copy
        assert all(b[k].dtype == raw[k].dtype and b[k].is_cuda for k in raw)
        got = tf(b, l2a.to(dev))
        got = {k: v.cpu() for k, v in got.items()}
    assert list(got) == meta["order"]
    for m, w in want.items():
        g = got[m].numpy()
        assert g.dtype == w.dtype and g.shape == w.shape, m
        if w.dtype == np.int64:
            assert np.array_equal(g, w), m
        else:
            assert np.array_equal(np.isnan(g), np.isnan(w)), m
            assert np.array_equal(np.nan_to_num(g, nan=0.0), np.nan_to_num(w, nan=0.0)), m
