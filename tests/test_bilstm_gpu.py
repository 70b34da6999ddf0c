"""GPU parity: fvta_bilstm_fwd/bwd and fvta_gemm (through the C ABI) vs the CPU oracle
(oracle/tf_ops.py bidirectional_dynamic_rnn, fp64).  Tolerance 1e-4 relative (fp32)."""
import pytest
import torch

from oracle import tf_ops
from fvta_memexqa_b200 import ops
from tests.helpers import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.mark.parametrize("shape", [(5, 7, 3, 4, 6), (128, 100, 70, 50, 1), (33, 65, 129, 17, 0)])
def test_gemm_variants(cuda_device, shape):
    M, N, K, _, _ = shape
    g = torch.Generator().manual_seed(M)
    for ta in (False, True):
        for tb in (False, True):
            A = torch.randn((K, M) if ta else (M, K), generator=g)
            B = torch.randn((N, K) if tb else (K, N), generator=g)
            bias = torch.randn(N, generator=g)
            ref = (A.t() if ta else A).double() @ (B.t() if tb else B).double() + bias.double()
            out = ops.gemm(A.to(cuda_device), B.to(cuda_device), bias.to(cuda_device), trans_a=ta, trans_b=tb)
            assert rel_err(out, ref) < 1e-5
            out2 = ops.gemm(A.to(cuda_device), B.to(cuda_device), bias.to(cuda_device), act="tanh",
                            trans_a=ta, trans_b=tb)
            assert rel_err(out2, torch.tanh(ref)) < 1e-5
            acc = torch.ones(M, N, device=cuda_device)
            ops.gemm(A.to(cuda_device), B.to(cuda_device), trans_a=ta, trans_b=tb, out=acc, accumulate=True)
            assert rel_err(acc, ref - bias.double() + 1.0) < 1e-5


LSTM_CASES = [
    # B, J, In, d
    (5, 4, 6, 3),
    (37, 8, 200, 50),     # text reader shape (photo titles), C2
    (32, 25, 200, 50),    # question
    (16, 8, 100, 50),     # image reader
    (9, 5, 20, 128),      # d large enough that W_h does not fit shared memory
]


@pytest.mark.parametrize("case", LSTM_CASES, ids=[str(c) for c in LSTM_CASES])
def test_bilstm_parity(cuda_device, case):
    B, J, In, d = case
    dev = cuda_device
    g = torch.Generator().manual_seed(B * J)
    x = torch.randn(B, J, In, generator=g)
    lens = torch.randint(0, J + 1, (B,), generator=g)
    lens[0] = J
    if B > 1:
        lens[1] = 0       # zero-length sequence: zero outputs, zero final state
    kernel = (torch.rand(In + d, 4 * d, generator=g) * 2 - 1) * (6.0 / (In + 5 * d)) ** 0.5
    bias = torch.randn(4 * d, generator=g) * 0.1
    d_out = torch.randn(B, J, 2 * d, generator=g)
    d_last = torch.randn(B, 2 * d, generator=g)

    def oracle(dtype):
        xx = x.to(dtype).clone().requires_grad_(True)
        kk = kernel.to(dtype).clone().requires_grad_(True)
        bb = bias.to(dtype).clone().requires_grad_(True)
        (fw, bw), (fh, bh) = tf_ops.bidirectional_dynamic_rnn(xx, lens, kk, bb)
        out = torch.cat([fw, bw], -1)
        last = torch.cat([fh, bh], -1)
        ((out * d_out.to(dtype)).sum() + (last * d_last.to(dtype)).sum()).backward()
        return out, last, xx.grad, kk.grad, bb.grad

    ref = oracle(torch.float64)
    ref32 = oracle(torch.float32)
    xg = x.to(dev).requires_grad_(True)
    kg = kernel.to(dev).requires_grad_(True)
    bg = bias.to(dev).requires_grad_(True)
    out, last = ops.bilstm(xg, lens.to(torch.int32).to(dev), kg, bg)
    ((out * d_out.to(dev)).sum() + (last * d_last.to(dev)).sum()).backward()
    got = (out, last, xg.grad, kg.grad, bg.grad)
    for name, a, r, r32 in zip(("out", "last", "d_x", "d_kernel", "d_bias"), got, ref, ref32):
        tol = max(TOL, 3 * rel_err(r32, r))
        assert rel_err(a, r) < tol, (name, rel_err(a, r))
    # outputs past the length are exactly zero
    pad = (torch.arange(J).unsqueeze(0) >= lens.unsqueeze(1))
    assert out.detach().cpu()[pad].abs().max().item() == 0.0 if pad.any() else True


@pytest.mark.parametrize("shape", [(7, 6, 15, 6, 8), (300, 16, 80, 100, 100), (33, 9, 40, 12, 24)])
def test_charcnn_table_parity(cuda_device, shape):
    """fvta_charcnn_fwd/bwd (per-character table + gather-add) vs the oracle's conv1d
    (model_v2.py:52-70): outputs and the gradients of char_emb, filter and bias."""
    ntok, W, VC, cdim, cw = shape
    dev = cuda_device
    g = torch.Generator().manual_seed(ntok)
    ids = torch.randint(0, VC, (ntok, W), generator=g)
    emb = torch.randn(VC, cdim, generator=g) * 0.5
    filt = torch.randn(1, 5, cdim, cw, generator=g) * 0.1
    bias = torch.randn(cw, generator=g) * 0.1
    dy = torch.randn(ntok, cw, generator=g)

    def oracle(dtype):
        e = emb.to(dtype).clone().requires_grad_(True)
        f = filt.to(dtype).clone().requires_grad_(True)
        b = bias.to(dtype).clone().requires_grad_(True)
        x = e[ids].unsqueeze(0)                      # [1, ntok, W, cdim]
        out = tf_ops.conv1d(x, f, b)[0]
        (out * dy.to(dtype)).sum().backward()
        return out, e.grad, f.grad, b.grad

    ref = oracle(torch.float64)
    ref32 = oracle(torch.float32)
    e = emb.to(dev).requires_grad_(True)
    f = filt.to(dev).requires_grad_(True)
    b = bias.to(dev).requires_grad_(True)
    table = torch.matmul(e.unsqueeze(0), f[0])
    y = ops.CharCnnTableFn.apply(table, b, ids.to(torch.int32).to(dev))
    (y * dy.to(dev)).sum().backward()
    for name, a, r, r32 in zip(("y", "d_emb", "d_filter", "d_bias"), (y, e.grad, f.grad, b.grad), ref, ref32):
        tol = max(TOL, 3 * rel_err(r32, r))
        assert rel_err(a, r) < tol, (name, rel_err(a, r))
