"""Timeline of the tensor-core forward's warp roles (debug build only).

    NVCC_EXTRA=-DFVTA_TC_TRACE python fvta_memexqa_b200/build.py --force && python tools/tc_trace.py
Prints, for the first tiles of CTA (0,0), the SM-clock offsets of each pipeline event.
"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
from fvta_memexqa_b200 import _lib  # noqa: E402

dev = torch.device("cuda:0")
bench.attention_roofline(dev, iters=1)
torch.cuda.synchronize()
lib = ctypes.CDLL(_lib.LIB_PATH)
NT = 24
buf = (ctypes.c_longlong * (4 * NT * 8))()
rc = lib.fvta_tc_trace_read(buf)
assert rc == 0, rc
import numpy as np
a = np.array(buf[:], dtype=np.int64).reshape(4, NT, 8)
t0 = a[a > 0].min()
names = {0: ("TMA", ["issue"]), 1: ("MMA", ["a_rdy0", "a_rdy1", "commit"]),
         2: ("split", ["top", "waits", "free0", "chunk0", "free1", "chunk1", "bar1", "end"]),
         3: ("softmax", ["top", "waits", "ldtm", "logits", "bar2", "bar3", "unit0", "end"])}
for i in range(6, 12):
    print("tile %d" % i)
    for role in range(4):
        nm, evs = names[role]
        print("  %-8s" % nm, "  ".join("%s=%d" % (e, a[role, i, j] - t0) for j, e in enumerate(evs)))
