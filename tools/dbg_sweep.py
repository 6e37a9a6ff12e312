"""Timing experiments on the tcgen05 GEMM (needs a library built with MPMAE_BUILD_KNOBS=1: python -c "import os;
os.environ['MPMAE_BUILD_KNOBS']='1'; from mmearth_train_b200 import build; build.build(force=True)").  The MPMAE_TC_DBG knobs
disable pieces of the kernel; results are invalid, only the times mean something: which of stats / GELU / stores / tcgen05.ld / MMA / operand split bounds each shape."""
import os, sys
os.environ.setdefault("MPMAE_TC_DBG", "0")   # the library only re-reads the knob per launch when it is set at load time
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from bench_gemm import bench

B = 256
BACKEND = int(sys.argv[1]) if len(sys.argv) > 1 else 3
shapes = [("s0 pw1", 1, B * 19 * 64, 160, 40), ("s2 pw1", 1, B * 19 * 4, 640, 160), ("s0 da", 3, B * 19 * 64, 160, 40),
          ("s0 pw2", 0, B * 19 * 64, 40, 160), ("s2 pw2", 0, B * 19 * 4, 160, 640)]
knobs = [0, 1, 2, 4, 8, 32, 64, 3, 7, 15, 96, 127]
print("dbg bits: 1 no stats, 2 no GELU, 4 no stores, 8 no tcgen05.ld, 32 no MMA, 64 no split")
for name, mode, M, N, K in shapes:
    row = []
    for k in knobs:
        os.environ["MPMAE_TC_DBG"] = str(k)
        ms, gbs, tf = bench(mode, BACKEND, M, N, K, iters=5)
        row.append(f"{k}:{ms * 1e3:6.1f}")
    print(f"{name:8s} M={M} N={N} K={K} us: " + "  ".join(row), flush=True)
os.environ["MPMAE_TC_DBG"] = "0"
