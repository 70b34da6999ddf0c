"""Timeline of the tensor-core forward's warp roles (instrumented build only).

    python -m fvta_memexqa_b200.build --out fvta_memexqa_b200/lib/libfvta_trace.so -DFVTA_TC_TRACE
    FVTA_LIB_PATH=fvta_memexqa_b200/lib/libfvta_trace.so python tools/tc_trace.py
Prints, for the first tiles of CTA (0,0) at the C3-max shape, the SM-clock offsets of each pipeline
event and the per-role steady-state period.
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
r = bench.attention_roofline(dev, iters=1)
print("tile kernel %.1f us" % (r["launch_ms"] * 1e3))
torch.cuda.synchronize()
lib = ctypes.CDLL(_lib.LIB_PATH)
NT = 24
buf = (ctypes.c_longlong * (5 * NT * 8))()
rc = lib.fvta_tc_trace_read(buf)
assert rc == 0, rc
a = np.array(buf[:], dtype=np.int64).reshape(5, NT, 8)
t0 = a[a > 0].min()
names = {0: ("TMA", ["issue"]), 1: ("MMA", ["a_ready", "commit"]),
         2: ("split", ["top", "waited", "end"]),
         3: ("softmax", ["top", "acc_full", "logits", "end"]),
         4: ("wsum", ["top", "wg_ready", "end"])}
for i in range(0, 16):
    print("tile %d" % i)
    for role in range(5):
        nm, evs = names[role]
        print("  %-8s" % nm, "  ".join("%s=%d" % (e, a[role, i, j] - t0) for j, e in enumerate(evs)))
st = a[0, NT - 1]
cyc = st[3] - st[0]
ns = st[7] - st[4]
print("CTA (0,0): entry->exit %d cycles = %d ns (SM clock %.0f MHz); prologue %d cycles, pipeline %d, tail %d"
      % (cyc, ns, 1e3 * cyc / max(ns, 1), st[1] - st[0], st[2] - st[1], st[3] - st[2]))
print("steady-state period (cycles per tile, tiles 6..15) and busy time per role:")
for role in range(5):
    nm, evs = names[role]
    last = len(evs) - 1
    per = (a[role, 15, 0] - a[role, 6, 0]) / 9.0
    if role >= 2:
        busy = np.mean([a[role, i, last] - a[role, i, 1] for i in range(6, 16)])
        wait = np.mean([a[role, i, 1] - a[role, i, 0] for i in range(6, 16)])
        print("  %-8s period %.0f  busy %.0f  waiting %.0f" % (nm, per, busy, wait))
    else:
        print("  %-8s period %.0f" % (nm, per))
