"""Which piece bounds each tcgen05 GEMM of a cfg2 step (all four stages + decoder), by switching pieces off.

Needs the knob build (results are invalid with knobs on; only the times mean something):
    nvcc ... -DMPMAE_TC_KNOBS=1 -o mmearth_train_b200/lib/libmpmae_knobs.so mmearth_train_b200/csrc/mpmae.cu
    MPMAE_LIB=mmearth_train_b200/lib/libmpmae_knobs.so python tools/dbg_sweep2.py
dbg bits: 1 no stats, 2 no GELU, 4 no stores, 8 no tcgen05.ld, 16 no per-element operand loads, 32 no MMA, 64 no operand split
"""
import ctypes as C
import os
import sys

os.environ.setdefault("MPMAE_TC_DBG", "0")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import mmearth_train_b200._native as nat  # noqa: E402


def bench(mode, M, N, K, a_gelu=0, group_rows=0, iters=6, backend=3, resid=False):
    dev = "cuda"
    t = lambda *s: torch.randn(*s, device=dev)
    nbuf = max(2, int(400e6 / (4.0 * M * max(N, K))))       # rotate operands so that nothing is served from L2
    nbuf = min(nbuf, 6)
    As = [t(M, K) for _ in range(nbuf)]
    auxs = [t(M, N) for _ in range(nbuf)]
    outs = [torch.empty(M, N, device=dev) for _ in range(nbuf)]
    b, bias = t(N, K) * 0.1, t(N)
    G = 1 if group_rows == 0 else (M + group_rows - 1) // group_rows
    kg, colsum, colsum2 = t(G, N), torch.zeros(G, N, device=dev), torch.zeros(N, device=dev)
    scratch = torch.empty(nat.gemm_scratch_floats(N, K), device=dev)
    gsq, gamma = torch.rand(K, device=dev) + 0.5, t(K)
    nx, sc, den = torch.empty(K, device=dev), torch.empty(K, device=dev), torch.empty(1, device=dev)
    ds = []
    for i in range(nbuf):
        d = nat.GemmDesc()
        for k, v in dict(a=As[i], b=b, bias=bias, aux=auxs[i], aux2=auxs[i], kg=kg, out=outs[i], colsum=colsum, colsum2=colsum2,
                         scratch=scratch).items():
            setattr(d, k, v.data_ptr())
        d.out2 = None
        d.resid = auxs[i].data_ptr() if resid else None
        d.M, d.N, d.K, d.group_rows, d.a_gelu = M, N, K, group_rows, a_gelu
        if a_gelu:
            d.grn_gsq, d.grn_gamma, d.grn_nx, d.grn_scale, d.grn_denom = (x.data_ptr() for x in (gsq, gamma, nx, sc, den))
            d.grn_eps = 1e-6
        ds.append(d)
    call = lambda i, st: nat.check(nat.lib.mpmae_gemm_epi(mode, backend, C.byref(ds[i % nbuf]), C.c_void_p(st)), "gemm_epi")
    return timed(call, iters)     # us; includes the (tiny) weight-split launch


GRAPH = bool(os.environ.get("SWEEP_GRAPH"))


def timed(call, iters):
    """us per call: eager launches, or (SWEEP_GRAPH=1) a CUDA graph of `iters` calls replayed -- no host time in the number"""
    st = torch.cuda.current_stream().cuda_stream
    for i in range(2):
        call(i, st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if GRAPH:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        g = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            call(0, side.cuda_stream)
            torch.cuda.synchronize()
            with torch.cuda.graph(g, stream=side):
                for i in range(iters):
                    call(i, side.cuda_stream)
        torch.cuda.synchronize()
        g.replay()
        torch.cuda.synchronize()
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters * 1e3
    e0.record()
    for i in range(iters):
        call(i, st)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def bench_tn(R, N, K, y_gelu, backend, iters=6):
    dev = "cuda"
    nbuf = 3
    xs = [torch.randn(R, N, device=dev) for _ in range(nbuf)]
    ys = [torch.randn(R, K, device=dev) for _ in range(nbuf)]
    dw = torch.zeros(N, K, device=dev)
    call = lambda i, st: nat.check(nat.lib.mpmae_gemm_wgrad_act(backend, xs[i % nbuf].data_ptr(), ys[i % nbuf].data_ptr(), dw.data_ptr(),
                                                                R, N, K, y_gelu, C.c_void_p(st)), "wgrad")
    return timed(call, iters)


if __name__ == "__main__":
    B = 256
    knobs = [0, 1, 2, 4, 8, 16, 32, 64, 96, 7, 23, 127]
    have_knobs = "knobs" in (os.environ.get("MPMAE_LIB") or "")
    if not have_knobs:
        knobs = [0]
    print("dbg bits: 1 no stats, 2 no GELU, 4 no stores, 8 no tcgen05.ld, 16 no per-element operand loads, 32 no MMA, 64 no operand split")
    print("knob:      " + "  ".join(f"{k:6d}" for k in knobs))
    only = sys.argv[1:] or None
    for stage, (P2, Cc) in enumerate([(64, 40), (16, 80), (4, 160), (1, 320)]):
        if only and f"s{stage}" not in only:
            continue
        R = B * 19 * P2
        for name, mode, N, K, ag, res in (("pw1", 1, 4 * Cc, Cc, 0, False), ("pw2", 0, Cc, 4 * Cc, 1, True), ("da", 3, 4 * Cc, Cc, 0, False),
                                          ("dvhat", 0, Cc, 4 * Cc, 0, False)):
            row = []
            for k in knobs:
                os.environ["MPMAE_TC_DBG"] = str(k)
                row.append(f"{bench(mode, R, N, K, a_gelu=ag, resid=res):6.1f}")
            print(f"s{stage} {name:6s} M={R:6d} N={N:4d} K={K:4d}: " + "  ".join(row), flush=True)
        os.environ["MPMAE_TC_DBG"] = "0"
        print(f"s{stage} dW2f(3xTF32,gelu) {bench_tn(R, Cc, 4 * Cc, 1, 3):6.1f}   dW1f(TF32) {bench_tn(R, 4 * Cc, Cc, 0, 2):6.1f}", flush=True)
    if only and "tn" in only:
        for R, N, K, name in ((311296, 40, 160, "s0"), (77824, 80, 320, "s1"), (19456, 160, 640, "s2"), (4864, 320, 1280, "s3")):
            print(f"{name} dW2f(3xTF32,gelu) {bench_tn(R, N, K, 1, 3):6.1f}   dW1f(TF32) {bench_tn(R, K, N, 0, 2):6.1f}", flush=True)
        M = B * 49
        print(f"dec dW2(TF32) {bench_tn(M, 512, 2048, 0, 2):6.1f}   dW1f(TF32) {bench_tn(M, 2048, 512, 0, 2):6.1f}   dW_pix {bench_tn(M, 2816, 512, 0, 2):6.1f}", flush=True)
        sys.exit(0)
    if not only or "dec" in only:
        M = B * 49
        for name, mode, N, K, gr in (("dec_pw1", 1, 2048, 512, 49), ("dec_pw2", 0, 512, 2048, 0), ("heads", 0, 2816, 512, 0),
                                     ("dec_dg", 2, 2048, 512, 49), ("dec_dvh", 0, 512, 2048, 0)):
            row = []
            for k in knobs:
                os.environ["MPMAE_TC_DBG"] = str(k)
                row.append(f"{bench(mode, M, N, K, group_rows=gr):6.1f}")
            print(f"{name:9s} M={M:6d} N={N:4d} K={K:4d}: " + "  ".join(row), flush=True)
        os.environ["MPMAE_TC_DBG"] = "0"
        print(f"dec dW2(TF32) {bench_tn(M, 512, 2048, 0, 2):6.1f}   dW1f(TF32) {bench_tn(M, 2048, 512, 0, 2):6.1f}", flush=True)
