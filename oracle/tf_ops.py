"""CPU restatement of the TensorFlow-1 primitives the FVTA hot path is built from.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker / the CPU baseline.

PARITY UNPINNED: TensorFlow 1.x cannot be imported in this environment (no wheel, no
network, Python 3.12) and the reference ships no tests or golden vectors (SURVEY.md
section 4, section 8c).  Every function here follows the reference call site it cites
plus the documented TF-1 semantics of the primitive; none of it has been checked against a
live TensorFlow.

All functions are dtype-generic (fp32 or fp64 torch CPU tensors) and differentiable through
torch autograd, which supplies the oracle gradients.  ``torch.amax`` is used for every
``tf.reduce_max`` because, like TF, it splits the gradient equally among tied maxima
(``torch.max(dim)`` does not).
"""
import torch

# utils.py:205 (also model_v2.py:10)
VERY_NEGATIVE_NUMBER = -1e30


def exp_mask(val, mask):
    """utils.py:210-213 -- ``val + (1 - cast(mask, float)) * -1e30``.

    The addition is carried out in ``val``'s dtype, so in fp32 a masked ``tanh`` logit
    becomes exactly ``float32(-1e30)`` (ulp(1e30) >> 1).
    """
    return val + (1.0 - mask.to(val.dtype)) * VERY_NEGATIVE_NUMBER


def softmax(logits):
    """model_v2.py:23-28 -- softmax over the last axis (flatten/reconstruct are views)."""
    return torch.softmax(logits, dim=-1)


def softsel(target, logits):
    """model_v2.py:39-48 -- ``sum_J softmax(logits)[..., J, None] * target[..., J, :]``.

    Materialises the broadcast product exactly as the reference does.
    """
    a = softmax(logits)
    return (a.unsqueeze(-1) * target).sum(dim=target.dim() - 2)


def linear(x, W, b, add_tanh=False):
    """model_v2.py:75-100 -- flatten to 2-D, ``x @ W + b``, optional tanh, reconstruct."""
    flat = x.reshape(-1, x.shape[-1])
    out = flat @ W + b
    if add_tanh:
        out = torch.tanh(out)
    return out.reshape(*x.shape[:-1], W.shape[1])


def l2_normalize(x, dim=-1, epsilon=1e-12):
    """tf.nn.l2_normalize [TF]: ``x * rsqrt(max(sum(x^2), eps))`` (model_v2.py:252-253)."""
    sq = (x * x).sum(dim=dim, keepdim=True)
    return x * torch.rsqrt(torch.clamp(sq, min=epsilon))


def matrix_band_part(T, num_lower, num_upper, dtype):
    """tf.matrix_band_part(ones([T,T]), lo, hi) [TF]: keep ``i-j<=lo`` and ``j-i<=hi``;
    a negative bound means unbounded (model_v2.py:321, :330, :337)."""
    i = torch.arange(T).unsqueeze(1)
    j = torch.arange(T).unsqueeze(0)
    keep = torch.ones(T, T, dtype=torch.bool)
    if num_lower >= 0:
        keep &= (i - j) <= num_lower
    if num_upper >= 0:
        keep &= (j - i) <= num_upper
    return keep.to(dtype)


def conv1d(x, filt, bias, keep_prob=1.0, is_train=False, drop_mask=None):
    """model_v2.py:52-70 -- char-CNN.

    ``x`` [Num, JX, W, cdim]; ``filt`` [1, height, cdim, cwdim] (TF HWIO); VALID, stride 1;
    ``relu(conv + bias)`` then ``reduce_max`` over the W-height+1 positions.
    Input dropout only when ``is_train`` (``drop_mask`` supplies the kept/(keep_prob) mask so
    a test can pin it; with keep_prob == 1.0 it is the identity).
    """
    if is_train and keep_prob < 1.0:
        if drop_mask is None:
            raise ValueError("oracle conv1d: pass drop_mask to pin dropout")
        x = x * drop_mask
    height = filt.shape[1]
    # windows: [Num, JX, W-height+1, cdim, height]
    win = x.unfold(2, height, 1)
    xc = torch.einsum("bjwch,hco->bjwo", win, filt[0]) + bias
    xc = torch.relu(xc)
    return torch.amax(xc, dim=2)


def basic_lstm_cell(x, c, h, kernel, bias, forget_bias=1.0):
    """tf.nn.rnn_cell.BasicLSTMCell [TF] (model_v2.py:652-653): gate order i, j, f, o;
    ``c' = c*sigmoid(f+forget_bias) + sigmoid(i)*tanh(j)``; ``h' = tanh(c')*sigmoid(o)``;
    ``kernel`` is ``[in+d, 4d]`` applied to ``concat([x, h])``."""
    gates = torch.cat([x, h], dim=1) @ kernel + bias
    i, j, f, o = gates.chunk(4, dim=1)
    new_c = c * torch.sigmoid(f + forget_bias) + torch.sigmoid(i) * torch.tanh(j)
    new_h = torch.tanh(new_c) * torch.sigmoid(o)
    return new_c, new_h


def dynamic_rnn(x, seq_len, kernel, bias):
    """tf.nn.dynamic_rnn(sequence_length=...) [TF]: outputs are zero past the length, the
    returned state is the state at the last valid step, zero initial state.
    ``x`` [B, J, in]; returns (outputs [B, J, d], final_h [B, d])."""
    B, J, _ = x.shape
    d = kernel.shape[1] // 4
    c = x.new_zeros(B, d)
    h = x.new_zeros(B, d)
    outs = []
    for t in range(J):
        nc, nh = basic_lstm_cell(x[:, t], c, h, kernel, bias)
        live = (t < seq_len).to(x.dtype).unsqueeze(1)
        c = live * nc + (1 - live) * c
        h = live * nh + (1 - live) * h
        outs.append(live * nh)
    if J == 0:
        return x.new_zeros(B, 0, d), h
    return torch.stack(outs, dim=1), h


def reverse_sequence(x, seq_len):
    """tf.reverse_sequence [TF]: reverse the first ``seq_len[b]`` steps of each row, leave
    the rest in place."""
    B, J = x.shape[0], x.shape[1]
    t = torch.arange(J).unsqueeze(0).expand(B, J)
    L = seq_len.unsqueeze(1)
    idx = torch.where(t < L, L - 1 - t, t)
    return torch.gather(x, 1, idx.reshape(B, J, *([1] * (x.dim() - 2))).expand_as(x))


def bidirectional_dynamic_rnn(x, seq_len, kernel_fw, bias_fw, kernel_bw=None, bias_bw=None):
    """tf.nn.bidirectional_dynamic_rnn(cell, cell, x, sequence_length) [TF]
    (model_v2.py:694 and the seven other call sites).  The reference passes the SAME cell
    object for both directions, so by default bw shares fw's weights.
    Returns ((fw_out, bw_out), (fw_h, bw_h))."""
    if kernel_bw is None:
        kernel_bw, bias_bw = kernel_fw, bias_fw
    fw_out, fw_h = dynamic_rnn(x, seq_len, kernel_fw, bias_fw)
    x_rev = reverse_sequence(x, seq_len)
    bw_out_rev, bw_h = dynamic_rnn(x_rev, seq_len, kernel_bw, bias_bw)
    bw_out = reverse_sequence(bw_out_rev, seq_len)
    return (fw_out, bw_out), (fw_h, bw_h)


def softmax_cross_entropy_with_logits(logits, labels):
    """tf.nn.softmax_cross_entropy_with_logits [TF] (model_v2.py:1088):
    ``-sum_c labels * log_softmax(logits)`` per row; an all-zero label row gives 0."""
    return -(labels * torch.log_softmax(logits, dim=-1)).sum(dim=-1)


def adadelta_update(var, grad, accum, accum_update, lr, rho=0.95, epsilon=1e-8):
    """tf.train.AdadeltaOptimizer [TF defaults rho=0.95, eps=1e-8] (trainer.py:16):
    accum = rho*accum + (1-rho)*g^2
    update = sqrt(accum_update+eps) * rsqrt(accum+eps) * g
    var -= lr*update ;  accum_update = rho*accum_update + (1-rho)*update^2
    Returns the new (var, accum, accum_update)."""
    accum = rho * accum + (1 - rho) * grad * grad
    update = torch.sqrt(accum_update + epsilon) * torch.rsqrt(accum + epsilon) * grad
    var = var - lr * update
    accum_update = rho * accum_update + (1 - rho) * update * update
    return var, accum, accum_update
