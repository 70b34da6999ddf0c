"""Literal CPU restatement of the FVTA graph in the reference's ``model_v2.py``.

TEST INFRASTRUCTURE ONLY (see ``oracle/tf_ops.py`` for who may import this) and
PARITY UNPINNED (TensorFlow cannot run here; the reference has no tests or golden vectors).

"Literal" means op-for-op: every ``tf.tile`` / ``tf.expand_dims`` / ``tf.concat`` the
reference performs is materialised at the same shape, including the ``[N,K,T,T,4d]`` tensor
of the time-warp block, so this file is only usable at small shapes (or batch 1).  Its job
is to pin the as-written semantics -- including the reference's quirks
-- against which ``oracle/collapsed.py`` and the CUDA kernels are checked.

Parameters travel in a dict keyed by the TensorFlow variable names the reference's scopes
produce.
"""
import torch

from .tf_ops import (exp_mask, softsel, linear, l2_normalize, matrix_band_part, conv1d,
                     bidirectional_dynamic_rnn, softmax_cross_entropy_with_logits)


def _similarity_logits(h_aug, q_aug, simiMatrix, W, b, add_tanh, cat_dim):
    """The four ``simiMatrix`` branches shared by model_v2.py:154-170 and :241-257."""
    if simiMatrix == 1:
        return linear(torch.cat([h_aug, q_aug, h_aug * q_aug], cat_dim), W, b, add_tanh).squeeze(cat_dim)
    if simiMatrix == 2:
        return linear(torch.cat([h_aug * q_aug, (h_aug - q_aug) * (h_aug - q_aug)], cat_dim),
                      W, b, add_tanh).squeeze(cat_dim)
    if simiMatrix == 3:
        return linear(torch.cat([h_aug, q_aug, (h_aug - q_aug) * (h_aug - q_aug), h_aug * q_aug], cat_dim),
                      W, b, add_tanh).squeeze(cat_dim)
    if simiMatrix == 4:
        return (l2_normalize(h_aug, -1) * l2_normalize(q_aug, -1)).sum(cat_dim)
    # model_v2.py:168-170 / :255-257 print and sys.exit()
    raise SystemExit("similarity matrix not implemented")


def attention(hinfo, hq, hinfo_mask=None, hq_mask=None, simiMatrix=1, W=None, b=None,
              add_tanh=False, bidirect=False):
    """model_v2.py:125-201.  ``hinfo`` [N, ..., w], ``hq`` [N, JQ, w] ->
    (h_a [N, w] (or [N, 2w] when bidirect), a_logits [N, V, JQ])."""
    N, w = hinfo.shape[0], hinfo.shape[-1]
    JQ = hq.shape[1]
    hinfo = hinfo.reshape(N, -1, w)
    if hinfo_mask is not None:
        hinfo_mask = hinfo_mask.reshape(N, -1)
    V = hinfo.shape[1]
    h_aug = hinfo.unsqueeze(2).repeat(1, 1, JQ, 1)
    q_aug = hq.unsqueeze(1).repeat(1, V, 1, 1)
    a_logits = _similarity_logits(h_aug, q_aug, simiMatrix, W, b, add_tanh, 3)
    if hinfo_mask is not None and hq_mask is not None:
        mask = hinfo_mask.unsqueeze(2).repeat(1, 1, JQ) & hq_mask.unsqueeze(1).repeat(1, V, 1)
        a_logits = exp_mask(a_logits, mask)
    h_a = softsel(hinfo, torch.amax(a_logits, dim=2))
    if bidirect:  # model_v2.py:184-192
        q_a = softsel(q_aug, a_logits).mean(dim=1)
        h_a = torch.cat([h_a, q_a], dim=1)
    return h_a, a_logits


def attention_3d(hinfo, hq, hinfo_mask=None, hq_mask=None, simiMatrix=1, W=None, b=None,
                 add_tanh=False, time_warp_att=False, C=None, bidirect=False):
    """model_v2.py:210-298.  ``hinfo`` [N,K,M,JMAX,w], ``hq`` [N,JQ,w] ->
    (h_a [N,w], a_logits [N,K,T,JQ]) with T = M*JMAX."""
    N, K, w = hinfo.shape[0], hinfo.shape[1], hinfo.shape[-1]
    JQ = hq.shape[1]
    hinfo = hinfo.reshape(N, K, -1, w)
    if hinfo_mask is not None:
        hinfo_mask = hinfo_mask.reshape(N, K, -1)
    T = hinfo.shape[2]
    h_aug = hinfo.unsqueeze(3).repeat(1, 1, 1, JQ, 1)
    q_aug = hq.unsqueeze(1).unsqueeze(1).repeat(1, K, T, 1, 1)
    a_logits = _similarity_logits(h_aug, q_aug, simiMatrix, W, b, add_tanh, 4)
    if hinfo_mask is not None and hq_mask is not None:
        mask = (hinfo_mask.unsqueeze(3).repeat(1, 1, 1, JQ)
                & hq_mask.unsqueeze(1).unsqueeze(1).repeat(1, K, T, 1))
        a_logits = exp_mask(a_logits, mask)
    a_logits_maxed = torch.amax(a_logits, dim=3)  # [N,K,T]
    if time_warp_att:  # model_v2.py:269-275 (C=None crashes in the reference too)
        a_temp = a_logits_maxed.unsqueeze(3).repeat(1, 1, 1, C.shape[-1])
        C_ex = C.unsqueeze(1).repeat(1, K, 1, 1)
        a_logits_maxed = (a_temp * C_ex).sum(-1)
    h_a = softsel(softsel(hinfo, a_logits_maxed), torch.amax(a_logits, dim=(3, 2)))
    if bidirect:
        # model_v2.py:286-289 concatenates [N,w] with [N,T,w]: a shape error in TF too.
        raise NotImplementedError("attention_3d(bidirect=True) is shape-broken in the reference")
    return h_a, a_logits


def time_indication_func(C, warp_type=1, window_t=3.0):
    """model_v2.py:301-341 -> (C', window_t or None).  ``window_t`` reaches the graph only
    through ceil->int64 (no gradient)."""
    if warp_type == 1:
        return C, None
    N, T = C.shape[0], C.shape[-1]
    if warp_type == 2:
        one = torch.diag(torch.ones(T, dtype=C.dtype)).unsqueeze(0).repeat(N, 1, 1)
        return one * C, None
    if warp_type == 3:
        past = matrix_band_part(T, -1, 0, C.dtype).unsqueeze(0).repeat(N, 1, 1)
        return past * C, None
    if warp_type == 4:
        future = matrix_band_part(T, 0, -1, C.dtype).unsqueeze(0).repeat(N, 1, 1)
        return future * C, None
    if warp_type == 5:
        import math
        w = int(math.ceil(float(window_t.detach() if hasattr(window_t, 'detach') else window_t)))
        win = matrix_band_part(T, w, w, C.dtype).unsqueeze(0).repeat(N, 1, 1)
        return win * C, window_t
    raise Exception("time warping type not implemented")


def time_warp(hall, lq, W_WH, b_WH, W_WC, b_WC, warp_type, window_t=3.0):
    """model_v2.py:953-1009.  ``hall`` [N,K,M,JMAX,2d], ``lq`` [N,2d] ->
    (warped hall [N,K,M,JMAX,2d], C_logits [N,T,T] (= model.C), C [N,T,T], window_t).
    ``hall_t1`` and ``hall_t2`` are built by the same expression, exactly as written
    (:979-980)."""
    N, K, M, JMAX, D = hall.shape
    hall_t = hall.reshape(N, K, -1, D)
    T = hall_t.shape[-2]
    hall_t1 = hall_t.unsqueeze(3).repeat(1, 1, 1, T, 1)
    hall_t2 = hall_t.unsqueeze(3).repeat(1, 1, 1, T, 1)
    WQ_t = lq.unsqueeze(1).unsqueeze(1).unsqueeze(1).repeat(1, K, T, T, 1)
    WH = linear(torch.cat([hall_t1 * hall_t2, (hall_t1 - hall_t2) * (hall_t1 - hall_t2)], 4), W_WH, b_WH)
    C_logits = linear(WH + WQ_t, W_WC, b_WC)
    C_logits = torch.tanh(C_logits.squeeze(4).sum(1)).reshape(N, T, T)
    C, win = time_indication_func(C_logits, warp_type, window_t)
    hall_tt = hall_t.unsqueeze(3).repeat(1, 1, 1, T, 1)
    C_ex = C.unsqueeze(1).repeat(1, K, 1, 1).unsqueeze(-1)
    h_warped = (hall_tt * C_ex).sum(-2)
    return h_warped.reshape(N, K, M, JMAX, D), C_logits, C, win


def context_tensor(hat, had, hwhen, hwhere, hpts, hpis, at_mask, ad_mask, when_mask, where_mask,
                   pts_mask, pis_mask, no_photo=False):
    """model_v2.py:862-914: zero-pad each modality to JMAX and stack on axis 1 in the order
    [at, ad, when, where, pts, pis] -> (hall [N,K,M,JMAX,2d], hall_mask [N,K,M,JMAX])."""
    N, M, JXA, D = hat.shape
    JD, JT, JG = had.shape[2], hwhen.shape[2], hwhere.shape[2]
    JI, JXP = hpts.shape[2], hpts.shape[3]
    JMAX = max(JXA, JD, JT, JG, JI * JXP, JI)

    def pad(x, J):
        return torch.nn.functional.pad(x, (0, 0, 0, JMAX - J))

    def padm(m, J):
        return torch.nn.functional.pad(m, (0, JMAX - J))

    hs = [pad(hat, JXA), pad(had, JD), pad(hwhen, JT), pad(hwhere, JG),
          pad(hpts.reshape(N, M, -1, D), JI * JXP)]
    ms = [padm(at_mask, JXA), padm(ad_mask, JD), padm(when_mask, JT), padm(where_mask, JG),
          padm(pts_mask.reshape(N, M, -1), JI * JXP)]
    if not no_photo:
        hs.append(pad(hpis, JI))
        ms.append(padm(pis_mask, JI))
    return torch.stack(hs, dim=1), torch.stack(ms, dim=1)


def fvta_block(hall, hall_mask, hq, lq, q_mask, params, config):
    """model_v2.py:953-1050: time warp -> attention_3d -> question attention.
    Returns a dict with hall_warped, C, C_win, att_logits, g1_all, gq, q_att_logits."""
    out = {}
    N = hall.shape[0]
    C = None
    if config.use_time_warp:
        hall, C_logits, C, win = time_warp(
            hall, lq, params["time_warp/WH/W"], params["time_warp/WH/b"],
            params["time_warp/WC/W"], params["time_warp/WC/b"], config.warp_type,
            params.get("time_warp/time_warp_C/time_warp_window_t", 3.0))
        out["C"] = C_logits
        out["C_win"] = win
        out["warp_h"] = hall
    g1_all, att_logits = attention_3d(
        hall, hq, hall_mask, q_mask, simiMatrix=config.simiMatrix,
        W=params.get("attention/all/att_logits/W"), b=params.get("attention/all/att_logits/b"),
        add_tanh=config.add_tanh, bidirect=config.use_bidirection,
        time_warp_att=config.use_time_warp_att, C=C)
    out["att_logits"] = att_logits
    out["g1_all"] = g1_all
    if config.use_question_att:
        gq, q_att_logits = attention(
            hq, g1_all.unsqueeze(1), q_mask, torch.ones(N, 1, dtype=torch.bool),
            simiMatrix=config.simiMatrix,
            W=params.get("question_emb/question_att/att_logits/W"),
            b=params.get("question_emb/question_att/att_logits/b"),
            add_tanh=config.add_tanh, bidirect=config.use_bidirection)
        out["q_att_logits"] = q_att_logits
    else:
        gq = lq
    out["gq"] = gq
    return out


def output_layer(gq, g1_all, gchoices, params, config):
    """model_v2.py:1053-1083 -> (logits [N,C], yp [N,C])."""
    C = gchoices.shape[1]
    g1_a_t = g1_all.unsqueeze(1).repeat(1, C, 1)
    gq = gq.unsqueeze(1).repeat(1, C, 1)
    W, b = params["output/choicelogits/W"], params["output/choicelogits/b"]
    if config.use_eu_output:
        x = torch.cat([gq, g1_a_t, gchoices, g1_a_t * gchoices, gq * gchoices,
                       (g1_a_t - gchoices) * (g1_a_t - gchoices), (gq - gchoices) * (gq - gchoices)], 2)
        logits = linear(x, W, b, add_tanh=config.add_tanh)
    else:
        x = torch.cat([gq, g1_a_t, gchoices, g1_a_t * gchoices, gq * gchoices], 2)
        logits = linear(x, W, b)
    logits = logits.squeeze(2)
    return logits, torch.softmax(logits, dim=-1)


def build_loss(logits, y, params=None, wd=None):
    """model_v2.py:1085-1096: mean over the STATIC batch N of softmax-CE (padded rows have
    all-False labels -> 0 loss but still count in the mean) + optional ``wd * l2_loss(var)``
    for every trainable variable reached by an ``add_wd`` call (model_v2.py:347-354)."""
    loss = softmax_cross_entropy_with_logits(logits, y.to(logits.dtype)).mean()
    if wd is not None and wd != 0.0 and params is not None:
        for name in wd_variable_names(params):
            loss = loss + wd * 0.5 * (params[name] ** 2).sum()
    return loss


def wd_variable_names(params):
    """Variables that receive weight decay, by the scopes ``add_wd`` is called in.

    add_wd(scope=current) collects TRAINABLE variables under the current scope:
    conv1d (:66-67, scope emb/conv/conv1d -- once per call, 7 calls, same variables each
    time), linear when it is handed ``wd`` (:96-97; only image_trans_linear :645 is),
    attention/attention_3d (:196-197, :293-294, the attention scope which contains
    att_logits; their inner ``linear`` is NOT given wd), reader (:834-836, scope "reader" =
    both LSTM cells).  Each call appends one more loss term, so a variable's multiplicity is
    the number of calls; returned with repetition."""
    names = []
    for n in params:
        if n.startswith("emb/conv/conv1d/"):
            names += [n] * 7
        elif n.startswith("emb/image/image_transform/image_trans_linear/"):
            names += [n]
        elif n.startswith("reader/"):
            names += [n]
        elif n.startswith("attention/all/att_logits/") or n.startswith("question_emb/question_att/att_logits/"):
            names += [n]
    return names


def embed(feed, params, config, is_train=False):
    """model_v2.py:524-646: char-CNN + word embedding (+ concat), image lookup + linear."""
    f = feed
    dtype = params["emb/word/var/word_emb_mat"].dtype
    word_emb_mat = torch.cat([params["emb/word/var/word_emb_mat"], f["existing_emb_mat"].to(dtype)], 0)
    names = ["at", "ad", "when", "where", "pts", "q", "choices"]
    x = {}
    for nme in names:
        A = word_emb_mat[f[nme].long()]
        if config.use_char:
            char_emb = params["emb/var/char_emb"]
            Ac = char_emb[f[nme + "_c"].long()]  # [..., J, W, cdim]
            lead = Ac.shape[:-3]
            Ac = Ac.reshape(-1, Ac.shape[-3], Ac.shape[-2], Ac.shape[-1])
            xc = conv1d(Ac, params["emb/conv/conv1d/filter"], params["emb/conv/conv1d/bias"],
                        config.keep_prob, is_train)
            xc = xc.reshape(*lead, xc.shape[-2], xc.shape[-1])
            A = torch.cat([xc, A], dim=-1)
        x[nme] = A
    xpis = f["image_emb_mat"].to(dtype)[f["pis"].long()]
    if config.use_image_trans:
        xpis = linear(xpis, params["emb/image/image_transform/image_trans_linear/W"],
                      params["emb/image/image_transform/image_trans_linear/b"], add_tanh=config.add_tanh)
    x["pis"] = xpis
    return x


def reader(x, feed, params, config):
    """model_v2.py:649-836: the eight bidirectional_dynamic_rnn call sites."""
    f = feed
    kt, bt = params["reader/text/utext/fw/basic_lstm_cell/kernel"], params["reader/text/utext/fw/basic_lstm_cell/bias"]
    ki, bi = params["reader/image/uimage/fw/basic_lstm_cell/kernel"], params["reader/image/uimage/fw/basic_lstm_cell/bias"]
    kt_bw = params.get("reader/text/utext/bw/basic_lstm_cell/kernel")
    bt_bw = params.get("reader/text/utext/bw/basic_lstm_cell/bias")
    ki_bw = params.get("reader/image/uimage/bw/basic_lstm_cell/kernel")
    bi_bw = params.get("reader/image/uimage/bw/basic_lstm_cell/bias")
    d = config.hidden_size
    out = {}

    def run(name, k, b, kb, bb):
        xx = x[name]
        mask = f[name + "_mask"]
        lens = mask.to(torch.int64).sum(-1)
        flat_x = xx.reshape(-1, xx.shape[-2], xx.shape[-1])
        (fw, bw), (fw_h, bw_h) = bidirectional_dynamic_rnn(flat_x, lens.reshape(-1), k, b, kb, bb)
        h = torch.cat([fw, bw], -1).reshape(*xx.shape[:-1], 2 * d)
        l = torch.cat([fw_h, bw_h], -1).reshape(*xx.shape[:-2], 2 * d)
        return h, l

    out["hq"], out["lq"] = run("q", kt, bt, kt_bw, bt_bw)
    for nme in ["at", "ad", "when", "where", "pts", "choices"]:
        out["h" + nme], out["l" + nme] = run(nme, kt, bt, kt_bw, bt_bw)
    out["hpis"], out["lpis"] = run("pis", ki, bi, ki_bw, bi_bw)
    return out


def forward(feed, params, config, is_train=False):
    """model_v2.py:493-1096, literal.  ``feed`` maps the placeholder names (model_v2.py:420-
    477) to torch tensors.  Returns every tensor the Model exposes plus the loss."""
    x = embed(feed, params, config, is_train)
    r = reader(x, feed, params, config)
    hall, hall_mask = context_tensor(
        r["hat"], r["had"], r["hwhen"], r["hwhere"], r["hpts"], r["hpis"],
        feed["at_mask"], feed["ad_mask"], feed["when_mask"], feed["where_mask"],
        feed["pts_mask"], feed["pis_mask"], no_photo=config.no_photo)
    out = {"hall": hall, "hall_mask": hall_mask, "hq": r["hq"], "lq": r["lq"],
           "lchoices": r["lchoices"]}
    out.update(fvta_block(hall, hall_mask, r["hq"], r["lq"], feed["q_mask"], params, config))
    logits, yp = output_layer(out["gq"], out["g1_all"], r["lchoices"], params, config)
    out["logits"], out["yp"] = logits, yp
    if "y" in feed:
        out["loss"] = build_loss(logits, feed["y"], params, config.wd)
    return out
