"""CPU restatement of the six MinkowskiEngine ops the FCMAE pretraining step uses.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is on the product path; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline leg may
import it (see DESIGN.md "Oracle").

The reference's sparse encoder is built on MinkowskiEngine 0.5.4 (vendored
submodule ``MinkowskiEngine/`` of the reference, fork with a depthwise op).  That
library cannot be compiled in this image (numpy.distutils, cblas.h, CUDA 12.9
CCCL clashes) and its depthwise op is GPU-only, so the algorithm is restated
here in pure torch (any dtype, CPU), operator by operator, with autograd:

* ``SparseTensor`` / coordinate manager: coordinates ``(b, y, x)`` int32 plus a
  tensor stride; strided coordinate sets are ``floor(c / s) * s``
  (``MinkowskiEngine/src/coordinate_map.hpp:59-66``,
  ``src/coordinate_map_gpu.cu:388-391``).
* kernel maps: every ``(k, in_row, out_row)`` with ``in = out + offset(k)``
  present in the input set (``src/coordinate_map_gpu.cu:1479-1547``); offset
  enumeration with axis 0 fastest, centred for odd kernel sizes and ``0..K-1``
  for even ones (``src/kernel_region.hpp:199-221``).
* ``MinkowskiConvolution``: ``y[o] = sum_k x[in(o,k)] @ W[k] (+ bias)``, kernel
  ``[K, Cin, Cout]``, bias ``[1, Cout]``
  (``MinkowskiEngine/MinkowskiConvolution.py:268-330``,
  ``src/convolution_kernel.cu:114-180``).
* ``MinkowskiDepthwiseConvolution``: ``y[o, c] = sum_k x[in(o,k), c] * W[k, c]``,
  kernel ``[K, C]`` (``MinkowskiEngine/MinkowskiDepthwiseConvolution.py:138-198``,
  ``src/depthwise_convolution_kernel.cu:27-52``).
* ``MinkowskiLinear`` / ``MinkowskiGELU``: ``nn.Linear`` / exact-erf GELU on the
  feature matrix (``MinkowskiOps.py:40-67``, ``MinkowskiNonlinearity.py:113-114``).
* ``to_sparse``: active set = pixels with ``sum_c |x| != 0``
  (``MinkowskiOps.py:279-317``); ``SparseTensor.dense()``
  (``MinkowskiSparseTensor.py:460-556``).

Pinning: the depthwise forward/backward known-answer vectors printed in
``MinkowskiDepthwiseConvolution.py:200-263`` and the sparse<->dense kernel
layout identity in ``helpers.py:676-690`` (tests/test_oracle_me_shim.py).

The class names mirror the ME names on purpose: ``install_as_minkowski()``
registers this module as ``MinkowskiEngine`` / ``MinkowskiOps`` so that the
reference's own ``models/convnextv2_sparse.py`` and ``models/fcmae.py`` run
unmodified on top of it (used only by ``oracle/make_golden.py`` in the build
container, where ``/root/reference`` exists).
"""
from __future__ import annotations

import math
import sys
import types
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn as nn


# --------------------------------------------------------------------------- coordinates
def kernel_offsets(kernel_size: Sequence[int], tensor_stride: Sequence[int],
                   dilation: Optional[Sequence[int]] = None) -> np.ndarray:
    """Offsets ``[K_volume, D]`` in the order ME enumerates kernel indices.

    ``src/kernel_region.hpp:199-221``: axis 0 is the fastest-varying digit of the
    kernel index; odd sizes are centred, even sizes start at the origin.
    """
    D = len(kernel_size)
    dilation = [1] * D if dilation is None else list(dilation)
    vol = int(np.prod(kernel_size))
    offs = np.zeros((vol, D), dtype=np.int64)
    for k in range(vol):
        r = k
        for i in range(D):
            ks = kernel_size[i]
            idx = r % ks
            if ks % 2 == 0:
                offs[k, i] = dilation[i] * tensor_stride[i] * idx
            else:
                offs[k, i] = (idx - ks // 2) * dilation[i] * tensor_stride[i]
            r //= ks
    return offs


def _encode(coords: np.ndarray) -> np.ndarray:
    """Injective int64 key for coordinate rows (batch, x0, x1, ...); offsets may be negative."""
    c = coords.astype(np.int64) + 4096
    key = np.zeros(len(c), dtype=np.int64)
    for i in range(c.shape[1]):
        key = key * 16384 + c[:, i]
    return key


class CoordinateMapKey:
    def __init__(self, uid: int, tensor_stride: Tuple[int, ...]):
        self.uid = uid
        self.tensor_stride = tuple(tensor_stride)

    def get_tensor_stride(self):
        return list(self.tensor_stride)


class CoordinateManager:
    """Holds coordinate sets and caches strided sets / kernel maps (ME coordinate_map_manager)."""

    def __init__(self, D: int):
        self.D = D
        self._coords: Dict[int, np.ndarray] = {}
        self._keys: Dict[int, CoordinateMapKey] = {}
        self._stride_cache: Dict[Tuple[int, Tuple[int, ...]], CoordinateMapKey] = {}
        self._next = 0

    def insert(self, coords: np.ndarray, tensor_stride: Tuple[int, ...]) -> CoordinateMapKey:
        key = CoordinateMapKey(self._next, tensor_stride)
        self._coords[key.uid] = np.ascontiguousarray(coords, dtype=np.int64)
        self._keys[key.uid] = key
        self._next += 1
        return key

    def coords(self, key: CoordinateMapKey) -> np.ndarray:
        return self._coords[key.uid]

    def stride(self, key: CoordinateMapKey, stride: Sequence[int]) -> CoordinateMapKey:
        """Strided coordinate set: floor(c / new_stride) * new_stride, de-duplicated."""
        stride = tuple(int(s) for s in stride)
        if all(s == 1 for s in stride):
            return key
        ck = (key.uid, stride)
        if ck in self._stride_cache:
            return self._stride_cache[ck]
        new_ts = tuple(t * s for t, s in zip(key.tensor_stride, stride))
        c = self._coords[key.uid].copy()
        ts = np.asarray(new_ts, dtype=np.int64)
        c[:, 1:] = np.floor_divide(c[:, 1:], ts) * ts
        _, first = np.unique(_encode(c), return_index=True)
        out = c[np.sort(first)]
        nk = self.insert(out, new_ts)
        self._stride_cache[ck] = nk
        return nk

    def kernel_map(self, in_key: CoordinateMapKey, out_key: CoordinateMapKey,
                   kernel_size: Sequence[int]) -> List[Tuple[np.ndarray, np.ndarray]]:
        """Per kernel index k: (in_rows, out_rows) with in = out + offset(k) present."""
        cin, cout = self._coords[in_key.uid], self._coords[out_key.uid]
        offs = kernel_offsets(kernel_size, in_key.tensor_stride)
        kin = _encode(cin)
        order = np.argsort(kin)
        kin_sorted = kin[order]
        maps = []
        for k in range(len(offs)):
            q = cout.copy()
            q[:, 1:] += offs[k]
            kq = _encode(q)
            pos = np.searchsorted(kin_sorted, kq)
            pos = np.clip(pos, 0, len(kin_sorted) - 1)
            hit = kin_sorted[pos] == kq
            out_rows = np.nonzero(hit)[0]
            in_rows = order[pos[hit]]
            maps.append((in_rows, out_rows))
        return maps


class SparseTensor:
    def __init__(self, features: torch.Tensor, coordinates: Optional[torch.Tensor] = None,
                 coordinate_map_key: Optional[CoordinateMapKey] = None,
                 coordinate_manager: Optional[CoordinateManager] = None, device=None,
                 tensor_stride=1):
        self._F = features
        if coordinate_manager is None:
            assert coordinates is not None
            D = coordinates.shape[1] - 1
            coordinate_manager = CoordinateManager(D)
            ts = (tensor_stride,) * D if isinstance(tensor_stride, int) else tuple(tensor_stride)
            coordinate_map_key = coordinate_manager.insert(coordinates.cpu().numpy(), ts)
        self._manager = coordinate_manager
        self.coordinate_map_key = coordinate_map_key

    # ME surface used by the reference model code
    @property
    def F(self):
        return self._F

    @property
    def C(self):
        return torch.from_numpy(self._manager.coords(self.coordinate_map_key)).int()

    @property
    def coordinate_manager(self):
        return self._manager

    @property
    def D(self):
        return self._manager.D

    @property
    def tensor_stride(self):
        return list(self.coordinate_map_key.tensor_stride)

    @property
    def shape(self):
        return self._F.shape

    def __len__(self):
        return self._F.shape[0]

    def __add__(self, other):
        """Same-coordinate-map addition (``MinkowskiTensor.py`` binary ops): features add row-wise."""
        assert isinstance(other, SparseTensor) and other.coordinate_map_key.uid == self.coordinate_map_key.uid
        return SparseTensor(self._F + other._F, coordinate_map_key=self.coordinate_map_key,
                            coordinate_manager=self._manager)

    def dense(self, shape=None, min_coordinate=None, contract_stride=True):
        """``MinkowskiSparseTensor.py:460-556`` with ``min_coordinate=None``."""
        c = self._manager.coords(self.coordinate_map_key)
        ts = np.asarray(self.coordinate_map_key.tensor_stride, dtype=np.int64)
        sp = c[:, 1:] // ts if contract_stride else c[:, 1:]
        b = c[:, 0]
        if shape is None:
            size = sp.max(0) + 1
            shape = (int(b.max()) + 1, self._F.shape[1], *[int(s) for s in size])
        out = torch.zeros(shape, dtype=self._F.dtype)
        idx = (torch.from_numpy(b), slice(None)) + tuple(torch.from_numpy(sp[:, i]) for i in range(sp.shape[1]))
        out[idx] = self._F
        return out, torch.zeros(len(ts), dtype=torch.int32), torch.IntTensor(list(ts))


def to_sparse(x: torch.Tensor, format: str = None, coordinates=None, device=None) -> SparseTensor:
    """``MinkowskiOps.py:279-317``: active pixels are those whose channel abs-sum is non-zero."""
    assert x.ndim > 2
    reduced = torch.abs(x).sum(1)
    bcoords = torch.where(reduced != 0)
    stacked = torch.stack(bcoords, dim=1).int()
    idx = (bcoords[0], slice(None)) + tuple(bcoords[1:])
    feats = x[idx]
    return SparseTensor(features=feats, coordinates=stacked)


# --------------------------------------------------------------------------- operators
def _tuple(v, D):
    return tuple(v) if isinstance(v, (list, tuple)) else (int(v),) * D


class MinkowskiConvolution(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=-1, stride=1, dilation=1, bias=False,
                 kernel_generator=None, expand_coordinates=False, convolution_mode=None, dimension=None):
        super().__init__()
        D = dimension
        self.dimension = D
        self.kernel_size = _tuple(kernel_size, D)
        self.stride = _tuple(stride, D)
        self.in_channels, self.out_channels = in_channels, out_channels
        vol = int(np.prod(self.kernel_size))
        self.kernel = nn.Parameter(torch.empty(vol, in_channels, out_channels))
        self.bias = nn.Parameter(torch.empty(1, out_channels)) if bias else None
        with torch.no_grad():  # ME reset_parameters (overwritten by the reference's _init_weights)
            stdv = 1.0 / math.sqrt(in_channels * vol)
            self.kernel.uniform_(-stdv, stdv)
            if self.bias is not None:
                self.bias.uniform_(-stdv, stdv)

    def forward(self, input: SparseTensor, coordinates=None):
        cm = input.coordinate_manager
        out_key = cm.stride(input.coordinate_map_key, self.stride)
        maps = cm.kernel_map(input.coordinate_map_key, out_key, self.kernel_size)
        n_out = len(cm.coords(out_key))
        out = input.F.new_zeros(n_out, self.out_channels)
        for k, (ir, orow) in enumerate(maps):
            if len(ir) == 0:
                continue
            out = out.index_add(0, torch.from_numpy(orow), input.F[torch.from_numpy(ir)] @ self.kernel[k])
        if self.bias is not None:
            out = out + self.bias
        return SparseTensor(out, coordinate_map_key=out_key, coordinate_manager=cm)


class MinkowskiDepthwiseConvolution(nn.Module):
    def __init__(self, in_channels, kernel_size=-1, stride=1, dilation=1, bias=False,
                 kernel_generator=None, convolution_mode=None, dimension=-1, use_cuda_kernel=True):
        super().__init__()
        D = dimension
        self.dimension = D
        self.kernel_size = _tuple(kernel_size, D)
        self.stride = _tuple(stride, D)
        self.in_channels = in_channels
        vol = int(np.prod(self.kernel_size))
        self.kernel = nn.Parameter(torch.empty(vol, in_channels))
        self.bias = nn.Parameter(torch.empty(1, in_channels)) if bias else None
        with torch.no_grad():
            stdv = 1.0 / math.sqrt(in_channels * vol)
            self.kernel.uniform_(-stdv, stdv)
            if self.bias is not None:
                self.bias.uniform_(-stdv, stdv)

    def forward(self, input: SparseTensor, coordinates=None):
        cm = input.coordinate_manager
        out_key = cm.stride(input.coordinate_map_key, self.stride)
        maps = cm.kernel_map(input.coordinate_map_key, out_key, self.kernel_size)
        n_out = len(cm.coords(out_key))
        out = input.F.new_zeros(n_out, self.in_channels)
        for k, (ir, orow) in enumerate(maps):
            if len(ir) == 0:
                continue
            out = out.index_add(0, torch.from_numpy(orow), input.F[torch.from_numpy(ir)] * self.kernel[k])
        if self.bias is not None:
            out = out + self.bias
        return SparseTensor(out, coordinate_map_key=out_key, coordinate_manager=cm)


class MinkowskiLinear(nn.Module):
    def __init__(self, in_features, out_features, bias=True):
        super().__init__()
        self.linear = nn.Linear(in_features, out_features, bias=bias)

    def forward(self, input: SparseTensor):
        return SparseTensor(self.linear(input.F), coordinate_map_key=input.coordinate_map_key,
                            coordinate_manager=input.coordinate_manager)


class MinkowskiGELU(nn.Module):
    def __init__(self):
        super().__init__()
        self.module = nn.GELU()

    def forward(self, input: SparseTensor):
        return SparseTensor(self.module(input.F), coordinate_map_key=input.coordinate_map_key,
                            coordinate_manager=input.coordinate_manager)


def install_as_minkowski() -> None:
    """Register this shim as ``MinkowskiEngine`` and ``MinkowskiOps`` in ``sys.modules``."""
    me = types.ModuleType("MinkowskiEngine")
    for name in ("SparseTensor", "MinkowskiConvolution", "MinkowskiDepthwiseConvolution",
                 "MinkowskiLinear", "MinkowskiGELU", "CoordinateManager", "CoordinateMapKey"):
        setattr(me, name, globals()[name])
    ops = types.ModuleType("MinkowskiOps")
    ops.to_sparse = to_sparse
    me.MinkowskiOps = ops
    sys.modules["MinkowskiEngine"] = me
    sys.modules["MinkowskiOps"] = ops
