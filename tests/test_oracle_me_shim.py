"""Pins oracle/me_shim.py (CPU restatement of the MinkowskiEngine ops) against the ONLY golden vectors the
reference holds for this path: the depthwise known-answer values printed in
MinkowskiEngine/MinkowskiEngine/MinkowskiDepthwiseConvolution.py:200-263, and the sparse<->dense kernel
layout identity of helpers.py:676-690."""
import numpy as np
import torch
import torch.nn.functional as F

from oracle import me_shim as me
from oracle.fcmae_oracle import me_to_torch_conv_weight


def test_depthwise_known_answer_forward_backward():
    # 1-D, K=2, stride 1: in [[0,1],[1,0],[1,1]], W [[1,2],[2,1]] -> out [[2,2],[3,1],[1,2]]
    coords = torch.IntTensor([[0, 0], [0, 1], [0, 2]])
    feats = torch.tensor([[0.0, 1.0], [1.0, 0.0], [1.0, 1.0]], requires_grad=True)
    x = me.SparseTensor(features=feats, coordinates=coords)
    conv = me.MinkowskiDepthwiseConvolution(2, kernel_size=2, stride=1, bias=False, dimension=1)
    with torch.no_grad():
        conv.kernel.copy_(torch.tensor([[1.0, 2.0], [2.0, 1.0]]))
    y = conv(x)
    assert torch.equal(y.F.detach(), torch.tensor([[2.0, 2.0], [3.0, 1.0], [1.0, 2.0]]))
    y.F.sum().backward()
    assert torch.equal(feats.grad, torch.tensor([[1.0, 2.0], [3.0, 3.0], [3.0, 3.0]]))
    assert torch.equal(conv.kernel.grad, torch.tensor([[2.0, 2.0], [2.0, 1.0]]))


def test_kernel_offsets_axis0_fastest_centred_and_even():
    o3 = me.kernel_offsets((3, 3), (1, 1))
    assert o3[0].tolist() == [-1, -1] and o3[1].tolist() == [0, -1] and o3[3].tolist() == [-1, 0]
    o2 = me.kernel_offsets((2, 2), (4, 4))
    assert o2.tolist() == [[0, 0], [4, 0], [0, 4], [4, 4]]


def _dense_to_sparse(x):
    return me.to_sparse(x)


def test_sparse_conv_equals_dense_conv_all_active():
    """helpers.remap_checkpoint_keys (helpers.py:676-690) maps ME kernels to torch conv weights by
    permute -> reshape(ks, ks) -> transpose(3, 2); with every pixel active the sparse conv must equal the
    zero-padded dense conv with the remapped weight."""
    torch.manual_seed(0)
    x = torch.randn(2, 5, 6, 6)
    conv = me.MinkowskiConvolution(5, 7, kernel_size=3, stride=1, bias=True, dimension=2)
    y = conv(_dense_to_sparse(x)).dense()[0]
    k = conv.kernel.detach()
    w = k.permute(2, 1, 0).reshape(7, 5, 3, 3).transpose(3, 2)          # the reference's remap
    assert torch.allclose(w, me_to_torch_conv_weight(k, 3))
    ref = F.conv2d(x, w, conv.bias.detach().reshape(-1), padding=1)
    assert torch.allclose(y, ref, atol=1e-5)
    dw = me.MinkowskiDepthwiseConvolution(5, kernel_size=7, bias=True, dimension=2)
    yd = dw(_dense_to_sparse(x)).dense()[0]
    kd = dw.kernel.detach()
    wd = kd.permute(1, 0).reshape(5, 1, 7, 7).transpose(3, 2)
    refd = F.conv2d(x, wd, dw.bias.detach().reshape(-1), padding=3, groups=5)
    assert torch.allclose(yd, refd, atol=1e-5)


def test_strided_conv_floor_coordinates_and_partial_cells():
    """2x2 stride-2: out coord = floor(c / 2) * 2, an output cell exists iff any child is active."""
    x = torch.zeros(1, 1, 4, 4)
    x[0, 0, 1, 1] = 3.0     # only child (dy=1, dx=1) of cell (0,0)
    x[0, 0, 2, 3] = 5.0     # child (dy=0, dx=1) of cell (1,1)
    conv = me.MinkowskiConvolution(1, 1, kernel_size=2, stride=2, bias=False, dimension=2)
    with torch.no_grad():
        conv.kernel.copy_(torch.tensor([1.0, 10.0, 100.0, 1000.0]).reshape(4, 1, 1))   # k = dy + 2*dx
    y = conv(me.to_sparse(x))
    d = y.dense()[0]
    assert d.shape == (1, 1, 2, 2)
    assert d[0, 0, 0, 0].item() == 3000.0 and d[0, 0, 1, 1].item() == 500.0
    assert len(y) == 2
