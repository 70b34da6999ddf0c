"""Timeline of the tensor-core backward's warp roles (instrumented build only).

    python -m fvta_memexqa_b200.build --out _ab/libfvta_trace.so -DFVTA_TC_TRACE
    FVTA_LIB_PATH=_ab/libfvta_trace.so python tools/bt_trace.py
"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import bench  # noqa: E402
from fvta_memexqa_b200 import _lib  # noqa: E402

dev = torch.device("cuda:0")
r = bench.attention_roofline_bwd(dev, iters=1)
print("backward whole op %.1f us" % (r["launch_ms"] * 1e3))
torch.cuda.synchronize()
lib = ctypes.CDLL(_lib.LIB_PATH)
NT = 24
buf = (ctypes.c_longlong * (5 * NT * 8))()
assert lib.fvta_bt_trace_read(buf) == 0
a = np.array(buf[:], dtype=np.int64).reshape(5, NT, 8)
t0 = a[a > 0].min()
names = {0: ("TMA", ["issue"]), 1: ("MMA", ["a_ready", "ep_done"]),
         2: ("split", ["top", "full", "acc_free", "end"]),
         3: ("row", ["top", "zq", "bk_free", "acc_full", "ep_done", "end", "pass1", "entries"]),
         4: ("update", ["top", "rec_ready", "A", "barA", "B", "barB", "end"])}
for i in range(4, 14):
    print("tile %d" % i)
    for role in range(5):
        nm, evs = names[role]
        print("  %-8s" % nm, "  ".join("%s=%d" % (e, a[role, i, j] - t0) for j, e in enumerate(evs)))
print("steady-state period (cycles per tile, tiles 8..20):")
for role in range(5):
    nm, evs = names[role]
    print("  %-8s period %.0f" % (nm, (a[role, 20, 0] - a[role, 8, 0]) / 12.0), end="")
    for j in range(1, len(evs)):
        print("  %s +%.0f" % (evs[j], np.mean([a[role, i, j] - a[role, i, j - 1] for i in range(8, 21)])), end="")
    print()
cnt = (ctypes.c_int * 8)()
assert lib.fvta_bt_count_read(cnt) == 0
print("row warps x tiles %d, with a slow lane %d, slow lanes %d, rows with tied maxima %d (all launches)" % tuple(cnt[:4]))
it = (ctypes.c_longlong * 32)()
if hasattr(lib, "fvta_bt_iter_read") and lib.fvta_bt_iter_read(it) == 0:
    v = [x for x in it[:] if x > 0]
    print("pass B iteration starts (tile 10, thread 0):", [int(b - v[0]) for b in v])
