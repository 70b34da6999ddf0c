"""CPU: pins of the oracle itself.  The reference ships no tests or golden vectors
and TensorFlow cannot run here, so the oracle is pinned by (1) literal == collapsed, forward
and gradients; (2) finite differences; (3) frozen golden vectors; (4) hand-computable known
answers; (5) independent implementations of the primitives (torch.nn.LSTM, torch.optim.Adadelta,
F.conv2d)."""
import os

import numpy as np
import pytest
import torch

from oracle import literal, collapsed, tf_ops, full
from fvta_memexqa_b200.config import fvta_config, set_shapes
from fvta_memexqa_b200.params import ParamStore, NEVER_UPDATED
from tests.helpers import attn_case

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _block_outputs(cfg, params, inp, dtype):
    pp = {k: v.to(dtype).clone().requires_grad_(True) for k, v in params.items()}
    hall = inp["hall"].to(dtype).clone().requires_grad_(True)
    hq = inp["hq"].to(dtype).clone().requires_grad_(True)
    lq = inp["lq"].to(dtype).clone().requires_grad_(True)
    lit = literal.fvta_block(hall, inp["hall_mask"], hq, lq, inp["q_mask"], pp, cfg)
    col = collapsed.fvta_attention(hall, inp["hall_mask"], hq, inp["q_mask"], lq, pp, cfg)
    leaves = [hall, hq, lq] + [pp[k] for k in sorted(pp) if k.startswith(("time_warp/W", "attention/"))]
    return lit, col, leaves


@pytest.mark.parametrize("simi", [1, 2, 3, 4])
@pytest.mark.parametrize("warp_type", [1, 2, 3, 4, 5])
@pytest.mark.parametrize("twa", [False, True])
def test_literal_equals_collapsed(simi, warp_type, twa):
    cfg, params, inp = attn_case(2, 6, 2, 5, 6, 4, simi=simi, warp_type=warp_type, time_warp_att=twa, seed=simi)
    for dtype, tol in ((torch.float64, 1e-10), (torch.float32, 2e-4)):
        lit, col, leaves = _block_outputs(cfg, params, inp, dtype)
        assert (lit["g1_all"] - col["g1_all"]).abs().max() < tol
        valid = lit["att_logits"] > -1e29
        assert ((lit["att_logits"] - col["att_logits"]) * valid).abs().max() < tol
        assert torch.equal(lit["att_logits"][~valid], col["att_logits"][~valid])
        # model.C is constant along its last axis as written (model_v2.py:979-980)
        assert (lit["C"] - col["c"].unsqueeze(2)).abs().max() < tol
        assert (lit["warp_h"] - col["warp_h"]).abs().max() < tol * 10
        w = torch.randn(lit["g1_all"].shape, generator=torch.Generator().manual_seed(0)).to(dtype)
        g1 = torch.autograd.grad((lit["g1_all"] * w).sum(), leaves, allow_unused=True)
        g2 = torch.autograd.grad((col["g1_all"] * w).sum(), leaves, allow_unused=True)
        for a, b in zip(g1, g2):
            if a is None and b is None:
                continue
            a = torch.zeros_like(b) if a is None else a
            b = torch.zeros_like(a) if b is None else b
            assert (a - b).abs().max() < tol * 100


def test_no_time_warp_and_question_attention():
    cfg, params, inp = attn_case(2, 6, 2, 5, 6, 4, use_time_warp=False, seed=4)
    lit, col, _ = _block_outputs(cfg, params, inp, torch.float64)
    assert (lit["g1_all"] - col["g1_all"]).abs().max() < 1e-12
    # question attention (model_v2.py:1044) literal vs collapsed
    p = {k: v.double() for k, v in params.items()}
    W, b = p["question_emb/question_att/att_logits/W"], p["question_emb/question_att/att_logits/b"]
    hq, g1 = inp["hq"].double(), lit["g1_all"].detach()
    gq_l, a_l = literal.attention(hq, g1.unsqueeze(1), inp["q_mask"], torch.ones(2, 1, dtype=torch.bool),
                                  simiMatrix=2, W=W, b=b, add_tanh=True)
    gq_c, a_c = collapsed.question_attention(hq, g1, inp["q_mask"], 2, W, b, True)
    assert (gq_l - gq_c).abs().max() < 1e-12
    valid = a_l > -1e29
    assert ((a_l - a_c) * valid).abs().max() < 1e-12


def test_finite_differences_literal():
    """Autograd of the literal oracle vs central differences (fp64)."""
    cfg, params, inp = attn_case(1, 6, 1, 4, 4, 3, add_tanh=False, seed=2, density=1.0, q_density=1.0)
    p = {k: v.double() for k, v in params.items()}

    def f(hall, hq, lq):
        return literal.fvta_block(hall, inp["hall_mask"], hq, lq, inp["q_mask"], p, cfg)["g1_all"].sum()
    args = [inp[k].double().clone().requires_grad_(True) for k in ("hall", "hq", "lq")]
    assert torch.autograd.gradcheck(f, args, eps=1e-6, atol=1e-6, rtol=1e-5)


def test_known_answers():
    cfg, params, inp = attn_case(2, 6, 2, 5, 6, 4, seed=3)
    p = {k: v.double() for k, v in params.items()}
    # all-zero context -> every logit equal within a row -> uniform softmax -> g1_all = 0
    z = collapsed.fvta_attention(torch.zeros_like(inp["hall"]).double(), inp["hall_mask"], inp["hq"].double(),
                                 inp["q_mask"], inp["lq"].double(), p, cfg)
    assert z["g1_all"].abs().max() == 0
    # window counts at the edges and for T < 7: cnt = min(t,3)+min(T-1-t,3)+1
    assert collapsed.window_count(10, 5).tolist() == [4, 5, 6, 7, 7, 7, 7, 6, 5, 4]
    assert collapsed.window_count(4, 5).tolist() == [4, 4, 4, 4]
    assert collapsed.window_count(3, 3).tolist() == [1, 2, 3] and collapsed.window_count(3, 4).tolist() == [3, 2, 1]
    # band matrix row sums agree with matrix_band_part
    for T in (1, 2, 5, 9):
        assert torch.equal(tf_ops.matrix_band_part(T, 3, 3, torch.float64).sum(1), collapsed.window_count(T, 5, 3.0, torch.float64))
    # exp_mask in fp32: a masked tanh logit is exactly float32(-1e30); fully masked row -> uniform
    v = tf_ops.exp_mask(torch.tanh(torch.randn(8)), torch.zeros(8, dtype=torch.bool))
    assert torch.all(v == torch.tensor(-1e30, dtype=torch.float32))
    assert torch.allclose(tf_ops.softmax(v), torch.full((8,), 0.125))


def test_reduce_max_gradient_splits_ties():
    x = torch.tensor([[1.0, 3.0, 3.0, 2.0]], dtype=torch.float64, requires_grad=True)
    torch.amax(x, dim=1).sum().backward()
    assert x.grad.tolist() == [[0.0, 0.5, 0.5, 0.0]]


def test_dynamic_rnn_semantics_and_torch_lstm():
    """BasicLSTMCell/dynamic_rnn restatement vs torch.nn.LSTM (independent implementation):
    TF gate order i,j,f,o + forget_bias 1  <->  torch order i,f,g,o with +1 in the f bias."""
    torch.manual_seed(0)
    B, J, In, d = 4, 6, 5, 3
    x = torch.randn(B, J, In, dtype=torch.float64)
    lens = torch.tensor([6, 3, 0, 1])
    kernel = torch.randn(In + d, 4 * d, dtype=torch.float64) * 0.3
    bias = torch.randn(4 * d, dtype=torch.float64) * 0.1
    out, last = tf_ops.dynamic_rnn(x, lens, kernel, bias)
    lstm = torch.nn.LSTM(In, d, batch_first=True).double()
    i, j, f, o = kernel.chunk(4, dim=1)
    bi, bj, bf, bo = bias.chunk(4)
    with torch.no_grad():
        w = torch.cat([i, f, j, o], dim=1)
        lstm.weight_ih_l0.copy_(w[:In].t())
        lstm.weight_hh_l0.copy_(w[In:].t())
        lstm.bias_ih_l0.copy_(torch.cat([bi, bf + 1.0, bj, bo]))
        lstm.bias_hh_l0.zero_()
    for b in range(B):
        L = int(lens[b])
        if L == 0:
            assert out[b].abs().max() == 0 and last[b].abs().max() == 0
            continue
        ref, (hn, _) = lstm(x[b:b + 1, :L])
        assert (out[b, :L] - ref[0]).abs().max() < 1e-12
        assert L == J or out[b, L:].abs().max() == 0
        assert (last[b] - hn[0, 0]).abs().max() < 1e-12
    # backward direction = forward over the reversed valid prefix, written back in place
    (fw, bw), (fh, bh) = tf_ops.bidirectional_dynamic_rnn(x, lens, kernel, bias)
    ref, (hn, _) = lstm(torch.flip(x[1:2, :3], dims=[1]))
    assert (bw[1, :3] - torch.flip(ref[0], dims=[0])).abs().max() < 1e-12
    assert (bh[1] - hn[0, 0]).abs().max() < 1e-12


def test_conv1d_vs_torch_conv2d():
    torch.manual_seed(1)
    x = torch.randn(3, 4, 9, 6, dtype=torch.float64)              # [Num, JX, W, cdim]
    filt = torch.randn(1, 5, 6, 7, dtype=torch.float64)           # HWIO
    bias = torch.randn(7, dtype=torch.float64)
    out = tf_ops.conv1d(x, filt, bias)
    ref = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2), filt.permute(3, 2, 0, 1)) + bias.view(1, -1, 1, 1)
    ref = torch.relu(ref).amax(dim=3).permute(0, 2, 1)
    assert (out - ref).abs().max() < 1e-12


def test_adadelta_vs_torch_optim():
    torch.manual_seed(2)
    p = torch.nn.Parameter(torch.randn(10, dtype=torch.float64))
    opt = torch.optim.Adadelta([p], lr=0.5, rho=0.95, eps=1e-8)
    var, acc, acc_u = p.detach().clone(), torch.zeros(10, dtype=torch.float64), torch.zeros(10, dtype=torch.float64)
    for _ in range(3):
        g = torch.randn(10, dtype=torch.float64)
        p.grad = g.clone()
        opt.step()
        var, acc, acc_u = tf_ops.adadelta_update(var, g, acc, acc_u, 0.5)
        assert (var - p.detach()).abs().max() < 1e-12


def test_golden_tiny_step():
    """The frozen tiny-shape training step (tools/make_golden.py) is reproduced by both the
    literal and the collapsed whole-model oracles."""
    z = np.load(os.path.join(GOLD, "fvta_tiny.npz"))
    cfg = fvta_config(batch_size=2, hidden_size=4, char_emb_size=6, char_out_size=5, image_feat_dim=11,
                      image_trans_dim=7)
    set_shapes(cfg, M=2, JI=2, JXP=3, JD=4, JXA=3, JT=2, JG=2, JQ=4, JA=3, W=6, word_vocab_size=20,
               char_vocab_size=15, word_emb_size=8)
    feed = {k[5:]: torch.as_tensor(z[k]) for k in z.files if k.startswith("feed/")}
    params = {k[6:]: torch.as_tensor(z[k]) for k in z.files if k.startswith("param/")}
    for block in ("literal", "collapsed"):
        loss, grads, new_params, _, out = full.train_step(feed, params, {}, cfg, block=block,
                                                          never_updated=NEVER_UPDATED)
        assert abs(loss.item() - float(z["out/loss"])) < 1e-12
        for k in ("g1_all", "gq", "logits", "yp"):
            assert (out[k].detach() - torch.as_tensor(z["out/" + k])).abs().max() < 1e-11, k
        for k, g in grads.items():
            if g is not None:
                assert (g - torch.as_tensor(z["grad/" + k])).abs().max() < 1e-10, k
        for k, v in new_params.items():
            assert (v - torch.as_tensor(z["new_param/" + k])).abs().max() < 1e-10, k
    # window_t receives no gradient and is never updated (model_v2.py:334-335; trainer.py:23-27)
    k = "time_warp/time_warp_C/time_warp_window_t"
    assert float(z["new_param/" + k]) == 3.0 and ("grad/" + k) not in z.files


def test_golden_attention_c2():
    z = np.load(os.path.join(GOLD, "attn_c2.npz"))
    cfg, params, inp = attn_case(2, 6, 4, 64, 100, 25, seed=77)
    assert np.array_equal(inp["hall"].numpy(), z["hall"])        # the generator is frozen too
    p = {k: v.double() for k, v in params.items()}
    o = collapsed.fvta_attention(inp["hall"].double(), inp["hall_mask"], inp["hq"].double(), inp["q_mask"],
                                 inp["lq"].double(), p, cfg)
    assert np.abs(o["g1_all"].numpy() - z["g1_all"]).max() < 1e-12


def test_padded_batch_rows_count_in_the_mean():
    """model_v2.py:1090: reduce_mean over the static N; all-False label rows add 0 (C.6)."""
    logits = torch.randn(4, 4, dtype=torch.float64)
    y = torch.zeros(4, 4, dtype=torch.bool)
    y[0, 1] = y[1, 2] = True
    loss = literal.build_loss(logits, y)
    per = tf_ops.softmax_cross_entropy_with_logits(logits[:2], y[:2].double())
    assert abs(loss.item() - per.sum().item() / 4) < 1e-12
