"""O(T) CPU restatement of the FVTA block -- the algebra the CUDA kernels implement.

TEST INFRASTRUCTURE ONLY and PARITY UNPINNED (see ``oracle/tf_ops.py``).

``oracle/literal.py`` materialises what the reference materialises; this file computes the
same numbers without the ``[N,K,T,T,*]`` and ``[N,K,T,JQ,*]`` tensors:

* time warp (model_v2.py:953-1009): because ``hall_t1 == hall_t2`` as written (:979-980) the
  correlation is constant along t2 and the warped context is a per-position gate,
  ``h~[n,k,t,:] = h[n,k,t,:] * tanh(z[n,t]) * cnt(t)``;
* similarity (model_v2.py:241-257): ``W . [h*q ; (h-q)^2]`` etc. factor into one
  ``[T x D] . [D x JQ]`` contraction plus per-row and per-column terms (Appendix A.2).

``tests/test_oracle.py`` asserts literal == collapsed (forward and autograd gradients) in fp64
to 1e-10 and in fp32 to 1e-5, which is what lets the collapsed form stand in for the literal
one at shapes where the literal one does not fit in memory.
"""
import math

import torch

from .tf_ops import VERY_NEGATIVE_NUMBER


def window_count(T, warp_type, window_t=3.0, dtype=torch.float32):
    """Row sums of the indicator matrix of time_indication_func (model_v2.py:301-341):
    how many t2 each t1 keeps.  type 1: T; 2: 1; 3: t+1; 4: T-t; 5: band(ceil(window_t))."""
    t = torch.arange(T, dtype=dtype)
    if warp_type == 1:
        return torch.full((T,), float(T), dtype=dtype)
    if warp_type == 2:
        return torch.ones(T, dtype=dtype)
    if warp_type == 3:
        return t + 1
    if warp_type == 4:
        return T - t
    if warp_type == 5:
        w = float(int(math.ceil(float(window_t.detach() if hasattr(window_t, 'detach') else window_t))))
        return torch.clamp(t, max=w) + torch.clamp(T - 1 - t, max=w) + 1
    raise Exception("time warping type not implemented")


def time_warp_gate(hall_t, lq, W_WH, b_WH, W_WC, b_WC, warp_type, window_t=3.0):
    """``hall_t`` [N,K,T,D] -> (g [N,T] = tanh(z)*cnt, tanh(z) [N,T]).
    z[n,t] = sum_k u.h[n,k,t]^2 + K*((b_WH + lq[n]).W_WC + b_WC),  u = W_WH[:D] @ W_WC."""
    N, K, T, D = hall_t.shape
    wc = W_WC[:, 0]
    u = W_WH[:D] @ wc
    const = (b_WH + lq) @ wc + b_WC[0]
    z = ((hall_t * hall_t) @ u).sum(1) + K * const.unsqueeze(1)
    tz = torch.tanh(z)
    return tz * window_count(T, warp_type, window_t, hall_t.dtype), tz


def similarity_terms(simiMatrix, W, b, D, dtype):
    """Split ``att_logits/W`` into the factorised coefficients:
    returns (w_mul [D] multiplying h*q, w_lin [D] for W.h, w_quad [D] for W.h^2,
             wq_lin [D] for W.q, wq_quad [D] for W.q^2, bias)."""
    z = torch.zeros(D, dtype=dtype)
    if simiMatrix == 1:      # [h, q, h*q]
        Wh, Wq, Whq = W[:D, 0], W[D:2 * D, 0], W[2 * D:, 0]
        return Whq, Wh, z, Wq, z, b[0]
    if simiMatrix == 2:      # [h*q, (h-q)^2]
        W1, W2 = W[:D, 0], W[D:, 0]
        return W1 - 2 * W2, z, W2, z, W2, b[0]
    if simiMatrix == 3:      # [h, q, (h-q)^2, h*q]
        Wh, Wq, W2, W1 = W[:D, 0], W[D:2 * D, 0], W[2 * D:3 * D, 0], W[3 * D:, 0]
        return W1 - 2 * W2, Wh, W2, Wq, W2, b[0]
    raise SystemExit("similarity matrix not implemented")


def attention_logits(h, g, hq, simiMatrix, W, b, add_tanh):
    """Unmasked logits of attention_3d on the gated context, h [N,K,T,D] un-warped,
    g [N,T] (ones when there is no time warp)."""
    N, K, T, D = h.shape
    gg = g.unsqueeze(1).unsqueeze(-1)  # [N,1,T,1]
    if simiMatrix == 4:
        hw = h * gg
        hn = hw * torch.rsqrt(torch.clamp((hw * hw).sum(-1, keepdim=True), min=1e-12))
        qn = hq * torch.rsqrt(torch.clamp((hq * hq).sum(-1, keepdim=True), min=1e-12))
        return torch.einsum("nktd,njd->nktj", hn, qn)
    w_mul, w_lin, w_quad, wq_lin, wq_quad, bias = similarity_terms(simiMatrix, W, b, D, h.dtype)
    P = torch.einsum("nktd,njd->nktj", h, hq * w_mul)
    row = gg[..., 0] * (h @ w_lin) + gg[..., 0] ** 2 * ((h * h) @ w_quad)       # [N,K,T]
    col = hq @ wq_lin + (hq * hq) @ wq_quad                                    # [N,JQ]
    S = gg * P + row.unsqueeze(-1) + col.unsqueeze(1).unsqueeze(1) + bias
    return torch.tanh(S) if add_tanh else S


def fvta_attention(hall, hall_mask, hq, q_mask, lq, params, config):
    """time warp + attention_3d in O(N*K*T*(D+JQ)) memory.
    Returns dict(g1_all [N,D], att_logits [N,K,T,JQ], warp_h [N,K,M,JMAX,D] | None,
                 c [N,T] (= any column of model.C) | None, gate [N,T])."""
    N, K, M, JMAX, D = hall.shape
    h = hall.reshape(N, K, -1, D)
    T = h.shape[2]
    out = {}
    if config.use_time_warp:
        g, tz = time_warp_gate(h, lq, params["time_warp/WH/W"], params["time_warp/WH/b"],
                               params["time_warp/WC/W"], params["time_warp/WC/b"], config.warp_type,
                               params.get("time_warp/time_warp_C/time_warp_window_t", 3.0))
        out["c"] = tz
        out["warp_h"] = (h * g.unsqueeze(1).unsqueeze(-1)).reshape(N, K, M, JMAX, D)
    else:
        g = torch.ones(N, T, dtype=h.dtype)
        out["c"] = None
        out["warp_h"] = None
    out["gate"] = g
    a = attention_logits(h, g, hq, config.simiMatrix, params.get("attention/all/att_logits/W"),
                         params.get("attention/all/att_logits/b"), config.add_tanh)
    if hall_mask is not None and q_mask is not None:
        mask = hall_mask.reshape(N, K, T).unsqueeze(-1) & q_mask.unsqueeze(1).unsqueeze(1)
        a = a + (1.0 - mask.to(a.dtype)) * VERY_NEGATIVE_NUMBER
    m = torch.amax(a, dim=3)
    if config.use_time_warp_att:  # model_v2.py:269-275: m * sum_t' C[n,t,t'] = m * g
        m = m * g.unsqueeze(1)
    alpha = torch.softmax(m, dim=-1)
    ht = h * g.unsqueeze(1).unsqueeze(-1)
    r = torch.einsum("nkt,nktd->nkd", alpha, ht)
    beta = torch.softmax(torch.amax(a, dim=(3, 2)), dim=-1)
    out["g1_all"] = torch.einsum("nk,nkd->nd", beta, r)
    out["att_logits"] = a
    return out


def question_attention(hq, g1_all, q_mask, simiMatrix, W, b, add_tanh):
    """model_v2.py:1044 -> :125-201 with hinfo=hq [N,JQ,D], hq=g1_all[:,None,:]:
    returns (gq [N,D], q_att_logits [N,JQ,1])."""
    N, JQ, D = hq.shape
    a = attention_logits(hq.unsqueeze(1), torch.ones(N, JQ, dtype=hq.dtype), g1_all.unsqueeze(1),
                         simiMatrix, W, b, add_tanh)[:, 0]          # [N,JQ,1]
    a = a + (1.0 - q_mask.unsqueeze(-1).to(a.dtype)) * VERY_NEGATIVE_NUMBER
    alpha = torch.softmax(torch.amax(a, dim=2), dim=-1)
    return torch.einsum("nj,njd->nd", alpha, hq), a
