"""Warm time of the BiLSTM entry points (fvta_bilstm_fwd: projection GEMM + recurrence) at a few
sequence mixes, to separate per-step latency from throughput.  CUDA graph of 10 calls, events."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
from fvta_memexqa_b200 import ops, _lib  # noqa: E402

CASES = [("32 seqs x 25 steps (questions alone)", 32, 25, 25, 25),
         ("32 seqs x 8 steps", 32, 8, 8, 8),
         ("8 seqs x 25 steps (one CTA per direction)", 8, 25, 25, 25),
         ("1664 seqs, lengths 1..10", 1664, 10, 1, 10),
         ("1664 seqs, lengths 1..10 + 32 x 25", 1696, 25, 1, 10)]


def main():
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    In, d = 200, 50
    kern = (torch.randn(In + d, 4 * d, generator=g) * 0.1).to(dev)
    bias = torch.zeros(4 * d, device=dev)
    for name, B, J, lo, hi in CASES:
        lens = torch.randint(lo, hi + 1, (B,), generator=g).to(torch.int32)
        if B == 1696:
            lens[:32] = 25
        x = torch.randn(B, J, In, generator=g).to(dev)
        lens_d = lens.to(dev)

        for pname, path in (("CUDA cores", _lib.LSTM_FFMA), ("tcgen05   ", _lib.LSTM_TC)):
            def fwd():
                return ops.bilstm_forward(x, lens_d, kern, bias, path=path)

            with torch.no_grad():
                for _ in range(3):
                    fwd()
                torch.cuda.synchronize()
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr):
                    for _ in range(10):
                        fwd()
                gr.replay()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(5):
                    gr.replay()
                e1.record()
                torch.cuda.synchronize()
            print("%-45s %s fwd %7.1f us (projection GEMM + memset + recurrence)" % (name, pname, e0.elapsed_time(e1) * 1000 / 50))


if __name__ == "__main__":
    main()
