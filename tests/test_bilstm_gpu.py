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


TC_LSTM_CASES = [(5, 4, 6, 3), (37, 8, 200, 50), (32, 25, 200, 50), (100, 10, 100, 50), (16, 8, 100, 64)]


@pytest.mark.parametrize("case", TC_LSTM_CASES, ids=[str(c) for c in TC_LSTM_CASES])
def test_bilstm_tcgen05_recurrence(cuda_device, case):
    """The forward recurrence on tcgen05 (csrc/bilstm_tc.cu, FvtaLstmDesc.path = FVTA_LSTM_TC; north_star c):
    outputs, final states and the reserve it leaves for the BPTT, against the fp64 oracle and against
    the CUDA-core recurrence on the same inputs.  Fixed bound 1e-4 of the largest entry."""
    from fvta_memexqa_b200 import _lib
    B, J, In, d = case
    dev = cuda_device
    g = torch.Generator().manual_seed(3 * B + J)
    x = torch.randn(B, J, In, generator=g)
    lens = torch.randint(0, J + 1, (B,), generator=g)
    lens[0] = J
    if B > 1:
        lens[1] = 0
    kernel = (torch.rand(In + d, 4 * d, generator=g) * 2 - 1) * (6.0 / (In + 5 * d)) ** 0.5
    bias = torch.randn(4 * d, generator=g) * 0.1
    (fw, bw), (fh, bh) = tf_ops.bidirectional_dynamic_rnn(x.double(), lens, kernel.double(), bias.double())
    ref_out, ref_last = torch.cat([fw, bw], -1), torch.cat([fh, bh], -1)
    with torch.no_grad():
        args = (x.to(dev), lens.to(torch.int32).to(dev), kernel.to(dev), bias.to(dev))
        out_tc, last_tc, res_tc, _ = ops.bilstm_forward(*args, path=_lib.LSTM_TC)
        out_ff, last_ff, res_ff, _ = ops.bilstm_forward(*args, path=_lib.LSTM_FFMA)
    e_out, e_last = rel_err(out_tc, ref_out), rel_err(last_tc, ref_last)
    e_res = rel_err(res_tc, res_ff)
    print("[lstm tcgen05] %s out %.2e last %.2e reserve-vs-ffma %.2e" % (case, e_out, e_last, e_res))
    assert e_out < TOL and e_last < TOL and e_res < TOL
    pad = (torch.arange(J).unsqueeze(0) >= lens.unsqueeze(1))
    if pad.any():
        assert out_tc.cpu()[pad].abs().max().item() == 0.0


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
    table = ops.CharTableFn.apply(e, f)
    y = ops.CharCnnTableFn.apply(table, b, ids.to(torch.int32).to(dev))
    (y * dy.to(dev)).sum().backward()
    for name, a, r, r32 in zip(("y", "d_emb", "d_filter", "d_bias"), (y, e.grad, f.grad, b.grad), ref, ref32):
        tol = max(TOL, 3 * rel_err(r32, r))
        assert rel_err(a, r) < tol, (name, rel_err(a, r))


@pytest.mark.gpu
@pytest.mark.parametrize("use_char,sink", [(True, False), (True, True), (False, False)])
def test_token_embed_parity(cuda_device, use_char, sink):
    """TokenEmbedFn (fvta_embed_rows + fvta_charcnn_* into one [char | word] buffer) vs the
    reference formulation: concat(trainable table, fed table) lookup (model_v2.py:590-600) next to
    the conv1d features (:617), outputs and every gradient."""
    dev = cuda_device
    ntok, W, VC, cdim, cw, VW, VF, wd = 77, 9, 30, 8, 12, 5, 40, 16
    g = torch.Generator().manual_seed(5)
    ids = torch.randint(0, VW + VF, (ntok,), generator=g)
    ids[:6] = torch.tensor([0, VW - 1, VW, VW + VF - 1, 1, 1])
    ids_c = torch.randint(0, VC, (ntok, W), generator=g)
    new = torch.randn(VW, wd, generator=g)
    fixed = torch.randn(VF, wd, generator=g)
    emb = torch.randn(VC, cdim, generator=g) * 0.5
    filt = torch.randn(1, 5, cdim, cw, generator=g) * 0.1
    bias = torch.randn(cw, generator=g) * 0.1
    dy = torch.randn(ntok, (cw if use_char else 0) + wd, generator=g)

    def oracle(dtype):
        n_ = new.to(dtype).clone().requires_grad_(True)
        e = emb.to(dtype).clone().requires_grad_(True)
        f = filt.to(dtype).clone().requires_grad_(True)
        b = bias.to(dtype).clone().requires_grad_(True)
        A = torch.cat([n_, fixed.to(dtype)], 0)[ids]
        if use_char:
            xc = tf_ops.conv1d(e[ids_c].unsqueeze(0), f, b)[0]
            A = torch.cat([xc, A], -1)
        (A * dy.to(dtype)).sum().backward()
        return (A, n_.grad) + ((e.grad, f.grad, b.grad) if use_char else ())

    ref, ref32 = oracle(torch.float64), oracle(torch.float32)
    n_ = new.to(dev).requires_grad_(True)
    e = emb.to(dev).requires_grad_(True)
    f = filt.to(dev).requires_grad_(True)
    b = bias.to(dev).requires_grad_(True)
    g_new = torch.zeros(VW, wd, device=dev)
    g_bias = torch.zeros(cw, device=dev)
    g_emb, g_filt = torch.zeros_like(e), torch.zeros_like(f)
    table = ops.CharTableFn.apply(e, f, (g_emb, g_filt) if sink else None) if use_char else None
    out = ops.TokenEmbedFn.apply(n_, fixed.to(dev), ids.to(dev), table, b if use_char else None,
                                 ids_c.to(torch.int32).to(dev) if use_char else None,
                                 (g_new, g_bias if use_char else None) if sink else None)
    (out * dy.to(dev)).sum().backward()
    got = (out, g_new if sink else n_.grad) + \
        ((g_emb if sink else e.grad, g_filt if sink else f.grad, g_bias if sink else b.grad) if use_char else ())
    if sink:
        assert n_.grad is None and b.grad is None and e.grad is None and f.grad is None
    for name, a, r, r32 in zip(("out", "d_new", "d_emb", "d_filter", "d_bias"), got, ref, ref32):
        tol = max(TOL, 3 * rel_err(r32, r))
        assert rel_err(a, r) < tol, (name, rel_err(a, r))
    # the word-vector columns are copies: bit-exact
    assert torch.equal(out[:, (cw if use_char else 0):].cpu(), torch.cat([new, fixed], 0)[ids])


@pytest.mark.gpu
@pytest.mark.parametrize("add_tanh,sink,need_dx", [(True, True, False), (True, False, True), (False, True, True)])
def test_linear_backward_parity(cuda_device, add_tanh, sink, need_dx):
    """LinearFn (fvta_gemm + fvta_bias_act_grad) vs torch: y, dx, dW, db (model_v2.py:75-100)."""
    dev = cuda_device
    g = torch.Generator().manual_seed(11)
    M, K, N = 203, 37, 52
    x, W, b, dy = (torch.randn(*s, generator=g) for s in ((M, K), (K, N), (N,), (M, N)))
    W = W * 0.2

    def oracle(dtype):
        x_, W_, b_ = (t.to(dtype).clone().requires_grad_(True) for t in (x, W, b))
        y = x_ @ W_ + b_
        y = torch.tanh(y) if add_tanh else y
        (y * dy.to(dtype)).sum().backward()
        return y, x_.grad, W_.grad, b_.grad

    ref, ref32 = oracle(torch.float64), oracle(torch.float32)
    x_ = x.to(dev).requires_grad_(need_dx)
    W_, b_ = W.to(dev).requires_grad_(True), b.to(dev).requires_grad_(True)
    gW, gb = torch.zeros(K, N, device=dev), torch.zeros(N, device=dev)
    y = ops.linear(x_, W_, b_, add_tanh, *((gW, gb) if sink else (None, None)))
    (y * dy.to(dev)).sum().backward()
    got = (y, x_.grad if need_dx else ref[1], gW if sink else W_.grad, gb if sink else b_.grad)
    for name, a, r, r32 in zip(("y", "dx", "dW", "db"), got, ref, ref32):
        tol = max(TOL, 3 * rel_err(r32, r))
        assert rel_err(a.cpu().double(), r) < tol, (name, rel_err(a.cpu().double(), r))
