"""Stand-alone GEMM entry of the C ABI (mpmae_gemm_rows): fp32 SIMT tiles and the tcgen05/TMA/TMEM path against
torch fp64, over the ragged shapes the model produces (K = 40, N = 40, M not a multiple of 128, ...)."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [(1000, 160, 40), (4864, 40, 160), (300, 48, 64), (12544, 2816, 512), (777, 320, 80), (128, 16, 8),
          (2500, 640, 160), (4864, 512, 320), (19456, 80, 320), (65, 1280, 320), (4097, 96, 384)]


def _scratch(N, K):
    from mmearth_train_b200._native import gemm_scratch_floats   # 2 * N * ceil32(K): the split weight is padded to 32-column groups
    return gemm_scratch_floats(N, K)


def run(nat, backend, a, b, bias):
    M, K = a.shape
    N = b.shape[0]
    out = torch.full((M, N), float("nan"), device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    scratch = torch.empty(_scratch(N, K), device="cuda")
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


# 3xTF32 (backend 1): the tensor core truncates when it aligns a product to the fp32 accumulator, so the error grows with the
# rows one work item accumulates (R / splits): 1.9e-5 at 2.5 k rows per item, 2.3e-5 at 3.1 k (the (12544, 2048, 512) case since
# the split count is chosen to fill ONE wave of 148 CTAs: 32 tiles x 4 splits instead of 32 x 5 = 160 items in two waves).
@pytest.mark.parametrize("backend,tol", [(0, 2e-6), (2, 2e-3), (1, 3e-5)])
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


def _gelu_ref(x):
    return 0.5 * x * (1.0 + torch.erf(x / 2.0 ** 0.5))


def _gelu_grad_ref(x):
    return 0.5 * (1.0 + torch.erf(x / 2.0 ** 0.5)) + x * torch.exp(-0.5 * x * x) / (2.0 * torch.pi) ** 0.5


# (M, N, K, group_rows): full-size stage-2 shape (tile count just over a multiple of 148: tail splitting), a wide-K
# grouped shape (per-sample statistics on the fast path), a ragged one (generic epilogue path)
EPI_SHAPES = [(19456, 640, 160, 0), (12544 // 4, 512, 256, 49), (2500, 160, 40, 0), (38912, 160, 40, 0)]


@pytest.mark.parametrize("backend,tol", [(0, 2e-5), (1, 1e-4), (3, 2e-4)])
@pytest.mark.parametrize("shape", EPI_SHAPES)
def test_gemm_epilogues(native_lib, shape, backend, tol):
    """mpmae_gemm_epi modes 1 (bias + GELU + sum h^2) and 3 (GELU / GRN backward + column sums) against torch fp64."""
    nat = native_lib
    M, N, K, gr = shape
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g).cuda()
    b = (torch.randn(N, K, generator=g) * K ** -0.5).cuda()
    bias = torch.randn(N, generator=g).cuda()
    G = 1 if gr == 0 else (M + gr - 1) // gr
    st = torch.cuda.current_stream().cuda_stream

    def call(mode, **ptrs):
        d = nat.GemmDesc()
        for k, v in ptrs.items():
            setattr(d, k, v.data_ptr())
        d.M, d.N, d.K, d.group_rows = M, N, K, gr
        nat.check(nat.lib.mpmae_gemm_epi(mode, backend, C.byref(d), C.c_void_p(st)), "gemm_epi")
        torch.cuda.synchronize()

    scratch = torch.empty(_scratch(N, K), device="cuda")
    # mode 1
    out, out2 = torch.full((M, N), float("nan"), device="cuda"), torch.full((M, N), float("nan"), device="cuda")
    colsum = torch.zeros(G, N, device="cuda")
    call(1, a=a, b=b, bias=bias, out=out, out2=out2, colsum=colsum, scratch=scratch)
    ref_a = a.double() @ b.double().t() + bias.double()
    ref_h = _gelu_ref(ref_a)
    assert float((out.double() - ref_a).abs().max() / ref_a.abs().max()) < tol
    assert float((out2.double() - ref_h).abs().max() / ref_h.abs().max()) < tol
    rows = torch.arange(M, device="cuda") // (gr if gr else M)
    ref_cs = torch.zeros(G, N, dtype=torch.float64, device="cuda").index_add_(0, rows, ref_h ** 2)
    assert float((colsum.double() - ref_cs).norm() / ref_cs.norm()) < tol
    if gr:   # mode 3 takes one kg vector: single-group launches only
        return
    # mode 3: out = (a.b^T + kg * gelu(aux2)) * gelu'(aux2) ; colsum2 = column sums of out
    aux2 = torch.randn(M, N, generator=g).cuda()
    kg = torch.randn(1, N, generator=g).cuda() * 0.1
    out3, cs2 = torch.full((M, N), float("nan"), device="cuda"), torch.zeros(N, device="cuda")
    h2 = torch.nn.functional.gelu(aux2)          # the SIMT baseline reads h, the tcgen05 path recomputes it from aux2
    call(3, a=a, b=b, aux=h2, aux2=aux2, kg=kg, out=out3, colsum2=cs2, scratch=scratch)
    x2 = aux2.double()
    ref3 = (a.double() @ b.double().t() + kg.double() * _gelu_ref(x2)) * _gelu_grad_ref(x2)
    assert float((out3.double() - ref3).abs().max() / ref3.abs().max()) < tol
    assert float((cs2.double() - ref3.sum(0)).norm() / ref3.sum(0).norm()) < 20 * tol   # sums of sign-mixed terms


@pytest.mark.parametrize("backend,tol", [(0, 2e-5), (1, 1e-4), (3, 2e-4)])
@pytest.mark.parametrize("shape", [(38912, 40, 160, 0), (9728, 80, 320, 0), (2500, 160, 640, 0), (4864, 320, 1280, 0), (700, 96, 384, 0)])
def test_gemm_gelu_on_operand_and_acc_scale(native_lib, shape, backend, tol):
    """Round 2: `h` is never materialised.  (1) mode 0 with a_gelu: out = (gelu(a) * a_scale) . b^T + bias + resid, GELU and the
    GRN scale applied to the A operand on its way into the tensor core (pw2 of a sparse block, K = 4C incl. a K tail);
    (2) mode 3 with acc_scale: out = (a . b^T * acc_scale + kg * gelu(aux2)) * gelu'(aux2), aux (h) absent."""
    nat = native_lib
    M, N, K, _ = shape
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g).cuda()
    b = (torch.randn(N, K, generator=g) * K ** -0.5).cuda()
    bias = torch.randn(N, generator=g).cuda()
    resid = torch.randn(M, N, generator=g).cuda()
    a_scale = (1.0 + 0.3 * torch.randn(K, generator=g)).cuda()
    st = torch.cuda.current_stream().cuda_stream
    scratch = torch.empty(_scratch(N, K), device="cuda")

    def call(mode, **kw):
        d = nat.GemmDesc()
        for k, v in kw.items():
            setattr(d, k, v.data_ptr() if torch.is_tensor(v) else v)
        d.M, d.N, d.K, d.group_rows = M, N, K, 0
        nat.check(nat.lib.mpmae_gemm_epi(mode, backend, C.byref(d), C.c_void_p(st)), "gemm_epi")
        torch.cuda.synchronize()

    out = torch.full((M, N), float("nan"), device="cuda")
    call(0, a=a, b=b, bias=bias, resid=resid, out=out, scratch=scratch, a_gelu=1, a_scale=a_scale)
    ref = (_gelu_ref(a.double()) * a_scale.double()) @ b.double().t() + bias.double() + resid.double()
    assert float((out.double() - ref).abs().max() / ref.abs().max()) < tol
    out = torch.full((M, N), float("nan"), device="cuda")
    call(0, a=a, b=b, out=out, scratch=scratch, a_gelu=1)                      # no scale, no bias
    ref = _gelu_ref(a.double()) @ b.double().t()
    assert float((out.double() - ref).abs().max() / ref.abs().max()) < tol
    # mode 3 (the roles of N and K swap: da [M, K'] = dy [M, N'] . W): reuse the shapes transposed
    M3, N3, K3 = M, K, N
    dy = torch.randn(M3, K3, generator=g).cuda()
    w = (torch.randn(N3, K3, generator=g) * K3 ** -0.5).cuda()
    aux2 = torch.randn(M3, N3, generator=g).cuda()
    kg = torch.randn(1, N3, generator=g).cuda() * 0.1
    acc_scale = (1.0 + 0.3 * torch.randn(N3, generator=g)).cuda()
    out3, cs2 = torch.full((M3, N3), float("nan"), device="cuda"), torch.zeros(N3, device="cuda")
    d = nat.GemmDesc()
    for k, v in dict(a=dy, b=w, aux2=aux2, kg=kg, out=out3, colsum2=cs2, scratch=torch.empty(_scratch(N3, K3), device="cuda"),
                     acc_scale=acc_scale).items():
        setattr(d, k, v.data_ptr())
    d.M, d.N, d.K, d.group_rows = M3, N3, K3, 0
    nat.check(nat.lib.mpmae_gemm_epi(3, backend, C.byref(d), C.c_void_p(st)), "gemm_epi")
    torch.cuda.synchronize()
    x2 = aux2.double()
    ref3 = (dy.double() @ w.double().t() * acc_scale.double() + kg.double() * _gelu_ref(x2)) * _gelu_grad_ref(x2)
    assert float((out3.double() - ref3).abs().max() / ref3.abs().max()) < tol
    assert float((cs2.double() - ref3.sum(0)).norm() / ref3.sum(0).norm()) < 20 * tol


@pytest.mark.parametrize("backend,tol", [(1, 1e-4), (3, 2e-4)])
@pytest.mark.parametrize("shape", [(38912, 40, 160), (19456, 160, 640), (4864, 320, 1280), (2500, 96, 384)])
def test_gemm_in_kernel_grn_scale(native_lib, shape, backend, tol):
    """pw2 of a sparse block: A = saved pre-activation, consumed as gelu(a) * s with the batch-global GRN scale
    s = 1 + gamma * nx, nx = G / (mean G + eps), G = sqrt(gsq) (models/sparse_norm_layers.py:24-33) derived in the kernel's
    prologue from gsq = sum_rows gelu(a)^2; nx / scale / denom are written out for the backward pass."""
    nat = native_lib
    M, N, K = shape
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g).cuda()
    b = (torch.randn(N, K, generator=g) * K ** -0.5).cuda()
    bias = torch.randn(N, generator=g).cuda()
    resid = torch.randn(M, N, generator=g).cuda()
    gamma = torch.randn(K, generator=g).cuda()
    h = _gelu_ref(a.double())
    gsq = (h ** 2).sum(0)
    G = gsq.sqrt()
    den = G.mean() + 1e-6
    sc = 1 + gamma.double() * G / den
    ref = (h * sc) @ b.double().t() + bias.double() + resid.double()
    st = torch.cuda.current_stream().cuda_stream
    out = torch.full((M, N), float("nan"), device="cuda")
    nx, scale = torch.full((K,), float("nan"), device="cuda"), torch.full((K,), float("nan"), device="cuda")
    denom = torch.full((1,), float("nan"), device="cuda")
    d = nat.GemmDesc()
    for k, v in dict(a=a, b=b, bias=bias, resid=resid, out=out, scratch=torch.empty(_scratch(N, K), device="cuda"), grn_gsq=gsq.float(),
                     grn_gamma=gamma, grn_nx=nx, grn_scale=scale, grn_denom=denom).items():
        setattr(d, k, v.data_ptr())
    d.M, d.N, d.K, d.group_rows, d.a_gelu, d.grn_eps = M, N, K, 0, 1, 1e-6
    nat.check(nat.lib.mpmae_gemm_epi(0, backend, C.byref(d), C.c_void_p(st)), "gemm_epi")
    torch.cuda.synchronize()
    assert float((out.double() - ref).abs().max() / ref.abs().max()) < tol
    assert abs(float(denom) - float(den)) < 1e-5 * float(den)
    assert float((nx.double() - G / den).abs().max()) < 1e-5 * float((G / den).abs().max())
    assert float((scale.double() - sc).abs().max()) < 1e-4


@pytest.mark.parametrize("backend,tol", [(0, 2e-6), (1, 2e-5)])
@pytest.mark.parametrize("shape", [(2432, 40, 160), (608, 80, 320), (4864, 320, 1280), (9999, 160, 640)])
def test_gemm_wgrad_gelu_operand(native_lib, shape, backend, tol):
    """dW2f = dy^T . gelu(a): the activation is applied to the Y operand inside the kernel (splitter warps / on load)."""
    nat = native_lib
    R, N, K = shape
    g = torch.Generator().manual_seed(R + N + K)
    x = torch.randn(R, N, generator=g).cuda()
    y = torch.randn(R, K, generator=g).cuda()
    dw = torch.zeros(N, K, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    nat.check(nat.lib.mpmae_gemm_wgrad_act(backend, C.c_void_p(x.data_ptr()), C.c_void_p(y.data_ptr()), C.c_void_p(dw.data_ptr()),
                                           R, N, K, 1, C.c_void_p(st)), "gemm_wgrad_act")
    torch.cuda.synchronize()
    ref = x.double().t() @ _gelu_ref(y.double())
    assert float((dw.double() - ref).norm() / ref.norm()) < tol
