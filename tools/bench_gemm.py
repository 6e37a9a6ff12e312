"""Micro-benchmark of the fused-epilogue GEMMs at the shapes of one cfg2 block (CUDA events, inputs >> L2)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mmearth_train_b200._native as nat

def bench(mode, backend, M, N, K, iters=10, group_rows=0):
    dev = "cuda"
    t = lambda *s: torch.randn(*s, device=dev)
    a, b, bias = t(M, K), t(N, K) * 0.1, t(N)
    out, out2 = torch.empty(M, N, device=dev), torch.empty(M, N, device=dev)
    aux, aux2 = t(M, N), t(M, N)
    G = 1 if group_rows == 0 else (M + group_rows - 1) // group_rows
    kg, colsum, colsum2 = t(G, N), torch.zeros(G, N, device=dev), torch.zeros(N, device=dev)
    scratch = torch.empty(nat.gemm_scratch_floats(N, K), device=dev)
    d = nat.GemmDesc()
    for k, v in dict(a=a, b=b, bias=bias, aux=aux, aux2=aux2, kg=kg, out=out, out2=out2, colsum=colsum, colsum2=colsum2,
                     scratch=scratch).items():
        setattr(d, k, v.data_ptr())
    d.resid = aux.data_ptr() if mode == 0 and N <= K else None
    d.M, d.N, d.K, d.group_rows = M, N, K, group_rows
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(2):
        nat.check(nat.lib.mpmae_gemm_epi(mode, backend, C.byref(d), C.c_void_p(st)), "gemm_epi")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        nat.check(nat.lib.mpmae_gemm_epi(mode, backend, C.byref(d), C.c_void_p(st)), "gemm_epi")
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    io = {0: 1 + (1 if d.resid else 0), 1: 2, 2: 2, 3: 2}[mode]
    gb = 4.0 * (M * K + N * K + M * N * io) / 1e9
    return ms, gb / (ms / 1e3), 2.0 * M * N * K / (ms / 1e3) / 1e12

if __name__ == "__main__":
    B = 256
    backend = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    for stage, (P2, Cc) in enumerate([(64, 40), (16, 80), (4, 160), (1, 320)]):
        R = B * 19 * P2
        for name, mode, N, K in (("pw1", 1, 4 * Cc, Cc), ("pw2", 0, Cc, 4 * Cc), ("da", 3, 4 * Cc, Cc), ("dvhat", 0, Cc, 4 * Cc)):
            ms, gbs, tf = bench(mode, backend, R, N, K)
            print(f"stage{stage} {name:6s} M={R:7d} N={N:5d} K={K:5d}  {ms*1e3:8.1f} us  {gbs:7.0f} GB/s  {tf:6.1f} TF/s")
    for name, mode, M, N, K, gr in (("dec_pw1", 1, B * 49, 2048, 512, 49), ("dec_pw2", 0, B * 49, 512, 2048, 0), ("heads", 0, B * 49, 2816, 512, 0),
                                    ("dec_dg", 2, B * 49, 2048, 512, 49)):
        ms, gbs, tf = bench(mode, backend, M, N, K, group_rows=gr)
        print(f"{name:14s} M={M:7d} N={N:5d} K={K:5d}  {ms*1e3:8.1f} us  {gbs:7.0f} GB/s  {tf:6.1f} TF/s")
