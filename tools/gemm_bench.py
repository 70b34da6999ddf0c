"""Warm per-launch time of fvta_gemm at the shapes of the C2 training step (CUDA events around a
loop of launches on one stream), tensor-core kernel vs FVTA_GEMM_FFMA=1 (read once per process,
so the FFMA leg runs in a child process)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

# (name, M, N, K, trans_a, trans_b, accumulate, bias)
SHAPES = [
    ("char table            ", 80, 500, 100, 0, 0, 0, 0),
    ("image transform       ", 384, 100, 2537, 0, 0, 0, 1),
    ("text x-projection     ", 3584, 200, 200, 0, 0, 0, 1),
    ("image x-projection    ", 384, 200, 100, 0, 0, 0, 1),
    ("text dWh  (TA, acc)   ", 50, 200, 7168, 1, 0, 1, 0),
    ("text d_x  (TB)        ", 3584, 200, 200, 0, 1, 0, 0),
    ("text dWx  (TA, acc)   ", 200, 200, 3584, 1, 0, 1, 0),
    ("image transform dW    ", 2537, 100, 384, 1, 0, 1, 0),
]


def leg():
    import torch
    from fvta_memexqa_b200 import ops
    dev = torch.device("cuda:0")
    out = {}
    for name, M, N, K, ta, tb, acc, bias in SHAPES:
        A = torch.randn((K, M) if ta else (M, K), device=dev)
        B = torch.randn((N, K) if tb else (K, N), device=dev)
        b = torch.randn(N, device=dev) if bias else None
        C = torch.zeros(M, N, device=dev)
        call = lambda: ops.gemm(A, B, b, None, bool(ta), bool(tb), out=C, accumulate=bool(acc))  # noqa: E731
        for _ in range(5):
            call()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(20):
                call()
        g.replay()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        out[name] = e0.elapsed_time(e1) * 1000.0 / 200.0
    print(json.dumps(out))


def one(idx):
    """Three eager launches of one shape (for ncu)."""
    import torch
    from fvta_memexqa_b200 import ops
    dev = torch.device("cuda:0")
    name, M, N, K, ta, tb, acc, bias = SHAPES[idx]
    A = torch.randn((K, M) if ta else (M, K), device=dev)
    B = torch.randn((N, K) if tb else (K, N), device=dev)
    b = torch.randn(N, device=dev) if bias else None
    C = torch.zeros(M, N, device=dev)
    for _ in range(3):
        ops.gemm(A, B, b, None, bool(ta), bool(tb), out=C, accumulate=bool(acc))
    torch.cuda.synchronize()


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "leg":
        leg()
    elif len(sys.argv) > 2 and sys.argv[1] == "one":
        one(int(sys.argv[2]))
    else:
        res = {}
        for tag, env in (("mma", {}), ("ffma", {"FVTA_GEMM_FFMA": "1"})):
            o = subprocess.run([sys.executable, __file__, "leg"], env=dict(os.environ, **env), capture_output=True, text=True)
            try:
                res[tag] = json.loads(o.stdout.strip().splitlines()[-1])
            except Exception:
                print(tag, "failed:", o.stderr[-800:])
        for name in [s[0] for s in SHAPES]:
            print("%s  %s" % (name, "  ".join("%s %6.2f us" % (t, res[t][name]) for t in res)))
