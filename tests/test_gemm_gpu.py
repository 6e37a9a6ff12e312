"""Stand-alone GEMM entry of the C ABI (mpmae_gemm_rows): fp32 SIMT tiles and the tcgen05/TMA/TMEM path against
torch fp64, over the ragged shapes the model produces (K = 40, N = 40, M not a multiple of 128, ...)."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [(1000, 160, 40), (4864, 40, 160), (300, 48, 64), (12544, 2816, 512), (777, 320, 80), (128, 16, 8),
          (2500, 640, 160), (4864, 512, 320), (19456, 80, 320), (65, 1280, 320), (4097, 96, 384)]


def run(nat, backend, a, b, bias):
    M, K = a.shape
    N = b.shape[0]
    out = torch.full((M, N), float("nan"), device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    scratch = torch.empty(2 * N * K, device="cuda")
    nat.check(nat.lib.mpmae_gemm_rows(backend, C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()),
                                      C.c_void_p(bias.data_ptr()) if bias is not None else None, C.c_void_p(out.data_ptr()),
                                      M, N, K, C.c_void_p(scratch.data_ptr()), C.c_void_p(st)), "gemm_rows")
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("backend,tol", [(0, 2e-6), (2, 1.5e-3), (1, 2e-5), (3, 6e-5)])
@pytest.mark.parametrize("shape", SHAPES)
def test_gemm_rows(native_lib, shape, backend, tol):
    M, N, K = shape
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g).cuda()
    b = torch.randn(N, K, generator=g).cuda()
    bias = torch.randn(N, generator=g).cuda()
    ref = (a.double() @ b.double().t() + bias.double())
    out = run(native_lib, backend, a, b, bias)
    assert torch.isfinite(out).all()
    err = float((out.double() - ref).abs().max() / ref.abs().max())
    assert err < tol, (shape, backend, err)


WG_SHAPES = [(2432, 40, 160), (2432, 160, 40), (608, 80, 320), (4096, 128, 128), (1000, 512, 320), (12544, 2048, 512),
             (333, 96, 384), (5000, 2816, 512)]


@pytest.mark.parametrize("backend,tol", [(0, 2e-6), (2, 2e-3), (1, 2e-5)])
@pytest.mark.parametrize("shape", WG_SHAPES)
def test_gemm_wgrad(native_lib, shape, backend, tol):
    """dW[N, K] += X[R, N]^T . Y[R, K]: contraction over rows (MN-major tcgen05 operands), accumulating."""
    nat = native_lib
    R, N, K = shape
    g = torch.Generator().manual_seed(R + N + K)
    x = torch.randn(R, N, generator=g).cuda()
    y = torch.randn(R, K, generator=g).cuda()
    init = torch.randn(N, K, generator=g).cuda()
    dw = init.clone()
    st = torch.cuda.current_stream().cuda_stream
    nat.check(nat.lib.mpmae_gemm_wgrad(backend, C.c_void_p(x.data_ptr()), C.c_void_p(y.data_ptr()), C.c_void_p(dw.data_ptr()),
                                       R, N, K, C.c_void_p(st)), "gemm_wgrad")
    torch.cuda.synchronize()
    ref = x.double().t() @ y.double()
    err = float(((dw - init).double() - ref).norm() / ref.norm())
    assert err < tol, (shape, backend, err)
    if backend != 0:   # localisation: a single non-zero row must land exactly
        x2 = torch.zeros_like(x)
        x2[R // 3] = 1.0
        dw2 = torch.zeros(N, K, device="cuda")
        nat.check(nat.lib.mpmae_gemm_wgrad(backend, C.c_void_p(x2.data_ptr()), C.c_void_p(y.data_ptr()),
                                           C.c_void_p(dw2.data_ptr()), R, N, K, C.c_void_p(st)), "gemm_wgrad")
        torch.cuda.synchronize()
        assert gu_rel(dw2, y[R // 3].repeat(N, 1)) < tol


def gu_rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))
