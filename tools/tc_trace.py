"""Event timeline of CTA 0 of one tcgen05 GEMM launch (knob build: -DMPMAE_TC_KNOBS=1, MPMAE_LIB=...libmpmae_knobs.so).

roles: 0 producer: ring slot free (about to issue the TMA loads of a k-block)   1 MMA warp: k-block landed (full barrier)
       2 MMA warp: k-block split (about to issue its MMAs)                      3 epilogue warp 0: accumulator ready
       4 epilogue warp 0: tile drained
"""
import ctypes as C
import os
import sys

os.environ["MPMAE_TC_TRACE"] = "1"
os.environ.setdefault("MPMAE_TC_DBG", "0")
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import mmearth_train_b200._native as nat  # noqa: E402
from dbg_sweep2 import bench  # noqa: E402

SHAPES = {"s2pw1": (1, 19456, 640, 160, 0, False, 0), "s2pw2": (0, 19456, 160, 640, 1, True, 0), "s2da": (3, 19456, 640, 160, 0, False, 0),
          "s2dvhat": (0, 19456, 160, 640, 0, False, 0), "s3pw2": (0, 4864, 320, 1280, 1, True, 0),
          "decpw1": (1, 12544, 2048, 512, 0, False, 49), "decpw2": (0, 12544, 512, 2048, 0, False, 0), "s0pw1": (1, 311296, 160, 40, 0, False, 0)}
TN = {"s2dW2f": (19456, 160, 640, 1, 3), "s2dW1f": (19456, 640, 160, 0, 2), "s0dW2f": (311296, 40, 160, 1, 3), "s3dW1f": (4864, 1280, 320, 0, 2),
      "s3dW2f": (4864, 320, 1280, 1, 3)}
lib = C.CDLL(nat.LIB_PATH)
for name in sys.argv[1:] or ["s2pw2"]:
    if name in TN:
        from dbg_sweep2 import bench_tn
        M, N, K = TN[name][:3]
        us = bench_tn(*TN[name], iters=1)
    else:
        mode, M, N, K, ag, res, gr = SHAPES[name]
        us = bench(mode, M, N, K, a_gelu=ag, group_rows=gr, resid=res, iters=1)
    buf = (C.c_ulonglong * (8 * 256))()
    assert lib.mpmae_debug_tc_trace(buf) == 0
    ev = []
    for r in range(5):
        n = int(buf[r * 256])
        ev += [(int(buf[r * 256 + 1 + i]), r, i) for i in range(n)]
    ev.sort()
    t0 = ev[0][0]
    print(f"== {name} M={M} N={N} K={K}: {us:.1f} us/launch, {len(ev)} events of CTA 0 (us since its first event, 1.965 GHz)")
    names = ["slot_free", "landed", "split_done", "acc_ready", "tile_drained"]
    ev = [e for e in ev if e[0] > 0]
    t0 = ev[0][0]
    for t, r, i in ev[:int(os.environ.get("TRACE_EVENTS", "150"))]:
        print(f"  {(t - t0) / 1965.0:8.2f}  {names[r]:12s} #{i}")
