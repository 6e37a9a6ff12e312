import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mmearth_train_b200._native as nat
def run(backend, x, y):
    R, N = x.shape; K = y.shape[1]
    dw = torch.zeros(N, K, device="cuda")
    nat.check(nat.lib.mpmae_gemm_wgrad(backend, C.c_void_p(x.data_ptr()), C.c_void_p(y.data_ptr()), C.c_void_p(dw.data_ptr()), R, N, K, C.c_void_p(torch.cuda.current_stream().cuda_stream)), "wgrad")
    torch.cuda.synchronize()
    return dw
torch.manual_seed(0)
for (R, N, K) in [(2432, 40, 160), (2432, 160, 40), (608, 80, 320), (4096, 128, 128)]:
    x = torch.ones(R, N, device="cuda"); y = torch.ones(R, K, device="cuda")
    for be in (2, 1):
        d = run(be, x, y)
        print((R, N, K), "backend", be, "ones: min/max", float(d.min()), float(d.max()), "expected", R)
    # row-localised: y = 1, x[r] = 1 only for one row
    for r0 in (0, 5, 8, 31, 32, 100):
        x = torch.zeros(R, N, device="cuda"); x[r0] = 1.0
        y = torch.arange(K, device="cuda").float().repeat(R, 1) + 1
        d = run(2, x, y)
        ok = torch.allclose(d, y[0].repeat(N, 1))
        print("   row", r0, "ok" if ok else f"BAD: d[0,:4]={d[0,:4].tolist()} d[1,:4]={d[1,:4].tolist()} nnz={int((d!=0).sum())}")
    x = torch.randn(R, N, device="cuda"); y = torch.randn(R, K, device="cuda")
    ref = x.double().t() @ y.double()
    for be in (0, 2, 1):
        d = run(be, x, y)
        print("   random backend", be, "rel err", float((d.double() - ref).norm() / ref.norm()))
