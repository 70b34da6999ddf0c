"""Per-step timeline of the forward recurrence (library built with NVCC_EXTRA=-DFVTA_LSTM_TRACE)."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from fvta_memexqa_b200 import ops, _lib  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
B, J, In, d = int(sys.argv[1]) if len(sys.argv) > 1 else 8, 25, 200, 50
kern = (torch.randn(In + d, 4 * d, generator=g) * 0.1).to(dev)
bias = torch.zeros(4 * d, device=dev)
x = torch.randn(B, J, In, generator=g).to(dev)
lens = torch.full((B,), J, dtype=torch.int32, device=dev)
for _ in range(3):
    ops.bilstm_forward(x, lens, kern, bias)
torch.cuda.synchronize()
lib = ctypes.CDLL(_lib.LIB_PATH)
buf = (ctypes.c_longlong * 256)()
lib.fvta_lstm_trace_read(buf)
ev = ["start", "-", "-", "gate phase done", "barrier 1", "cell phase done", "barrier 2"]
for step in range(4, 9):
    t0 = buf[step * 8]
    print("step %d: " % step + ", ".join("%s +%d" % (ev[i], buf[step * 8 + i] - t0) for i in range(3, 7)))
print("step period (cycles):", [buf[(s + 1) * 8] - buf[s * 8] for s in range(2, 12)])
