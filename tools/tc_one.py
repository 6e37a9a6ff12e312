"""One eager launch of a named GEMM shape (for ncu): python tools/tc_one.py s0da [s0pw1 ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from dbg_sweep2 import bench  # noqa: E402

B = 256
SH = {}
for st, (P2, Cc) in enumerate([(64, 40), (16, 80), (4, 160), (1, 320)]):
    R = B * 19 * P2
    SH[f"s{st}pw1"] = (1, R, 4 * Cc, Cc, 0, False, 0)
    SH[f"s{st}pw2"] = (0, R, Cc, 4 * Cc, 1, True, 0)
    SH[f"s{st}da"] = (3, R, 4 * Cc, Cc, 0, False, 0)
    SH[f"s{st}dvhat"] = (0, R, Cc, 4 * Cc, 0, False, 0)
SH["decpw1"] = (1, B * 49, 2048, 512, 0, False, 49)
SH["decpw2"] = (0, B * 49, 512, 2048, 0, False, 0)
for name in sys.argv[1:]:
    mode, M, N, K, ag, res, gr = SH[name]
    print(name, bench(mode, M, N, K, a_gelu=ag, group_rows=gr, resid=res, iters=2))
