"""The FVTA model behind the reference's model-construction interface.

Mirrors the interface of the reference's ``model_v2.py``: ``get_model(config) -> Model``
(model_v2.py:12-17); ``Model`` exposes ``global_step, loss, logits, yp, summary``, one
attribute per placeholder (model_v2.py:420-477), the visualisation tensors
``C, C_win, att_logits, q_att_logits, warp_h, hall, JXP, g1_all, gq`` and
``get_feed_dict(batch, is_train)`` (model_v2.py:1099).  The free functions ``attention_3d``,
``attention`` and ``time_indication_func`` keep the reference signatures
(model_v2.py:210, :125, :301).

Execution is NOT a port of the TF graph: the hot path -- time-warp gate, similarity,
two-level masked softmax, weighted sum (model_v2.py:953-1024), the question attention
(:1044), the BiLSTM reader (:649-836) and Adadelta (trainer.py) -- runs in the hand-written
sm_100a kernels of libfvta_b200.so (include/fvta_b200.h); torch tensors provide device
memory, streams and autograd plumbing between them.  There is no CPU path: constructing a
Model without a CUDA device raises.
"""
import math

import torch

from . import ops
from ._lib import FvtaError
from .params import ParamStore, att_w_rows

VERY_NEGATIVE_NUMBER = -1e30  # model_v2.py:10

PLACEHOLDERS = [
    "at", "at_c", "at_mask", "ad", "ad_c", "ad_mask", "when", "when_c", "when_mask",
    "where", "where_c", "where_mask", "pts", "pts_c", "pts_mask", "pis", "pis_mask",
    "q", "q_c", "q_mask", "choices", "choices_c", "choices_mask", "y",
    "existing_emb_mat", "image_emb_mat", "is_train"]


def select_get_model(config):
    """main.py:156-165: pick the model module from the flags.  Only the FVTA module
    (``--use_3d`` -> model_v2) has a B200 implementation; the others are the paper's
    baselines and are out of scope."""
    if getattr(config, "dmnplus", False) or getattr(config, "mcb", False) or not config.use_3d:
        raise NotImplementedError(
            "only the FVTA model (--use_3d, reference model_v2.py) is implemented; "
            "model.py / model_mcb.py / model_dmnplus.py are baselines outside the FVTA hot path")
    return get_model


def get_model(config):
    """model_v2.py:12-17."""
    return Model(config, "model_%s" % config.modelname)


# ---------------------------------------------------------------------------------------------
# reference-signature free functions
# ---------------------------------------------------------------------------------------------
class VariableStore(dict):
    """Stand-in for TF variable scopes: ``scope`` + variable name -> tensor.  Variables are
    created on first use with the reference initialiser (``linear``: truncated normal 0.1 /
    zeros, model_v2.py:88-89) and reused afterwards."""

    def __init__(self, device="cuda", seed=0):
        super().__init__()
        self.device = torch.device(device)
        self.gen = torch.Generator().manual_seed(seed)

    def get(self, name, shape, kind):
        if name not in self:
            if kind == "trunc":
                t = torch.empty(shape)
                torch.nn.init.trunc_normal_(t, 0.0, 0.1, -0.2, 0.2, generator=self.gen)
            else:
                t = torch.zeros(shape)
            self[name] = t.to(self.device).requires_grad_(True)
        return self[name]


_default_store = None


def default_store():
    global _default_store
    if _default_store is None:
        _default_store = VariableStore()
    return _default_store


def _att_params(store, scope, simiMatrix, D):
    rows = att_w_rows(simiMatrix, D)
    if rows == 0:
        return None, None
    return (store.get(scope + "/att_logits/W", (rows, 1), "trunc"),
            store.get(scope + "/att_logits/b", (1,), "zeros"))


def attention_3d(hinfo, hq, hinfo_mask=None, hq_mask=None, simiMatrix=1, wd=None, add_tanh=False,
                 time_warp_att=False, C=None, bidirect=False, scope=None, store=None):
    """model_v2.py:210-298.  hinfo [N,K,M,JMAX,w], hq [N,JQ,w] -> (h_a [N,w], a_logits [N,K,T,JQ]).
    ``scope`` names the parameter set (``<scope>/att_logits/{W,b}``) inside ``store``.
    ``wd`` is accepted for signature parity; weight decay is applied by the Model's loss."""
    if simiMatrix not in (1, 2, 3, 4):
        print("similarity matrix not implemented")  # model_v2.py:255-257
        raise SystemExit(1)
    if bidirect:
        # model_v2.py:286-289 concatenates [N,w] with [N,T,w]; it cannot run in the reference either
        raise NotImplementedError("attention_3d(bidirect=True) is shape-broken in the reference")
    if time_warp_att:
        # model_v2.py:269-275 needs the dense C [N,T,T]; the fused path derives the same row sums
        # from the time-warp gate inside Model (config.use_time_warp_att with use_time_warp).
        raise NotImplementedError("time_warp_att with an external C is only available through Model")
    store = store or default_store()
    N, K, w = hinfo.shape[0], hinfo.shape[1], hinfo.shape[-1]
    h = hinfo.reshape(N, K, -1, w)
    hm = hinfo_mask.reshape(N, K, -1) if hinfo_mask is not None else None
    W, b = _att_params(store, scope or "attention_2vector", simiMatrix, w)
    cfg = ops.AttnConfig(simiMatrix, add_tanh, False, 1, 3.0, False)
    g1, a, _, _ = ops.FvtaAttentionFn.apply(cfg, (True, False, False), h, hm, hq, hq_mask, None, W, b,
                                            None, None, None, None)
    return g1, a


def attention(hinfo, hq, hinfo_mask=None, hq_mask=None, simiMatrix=1, wd=None, add_tanh=False,
              bidirect=False, scope=None, store=None):
    """model_v2.py:125-201.  hinfo [N,...,w], hq [N,JQ,w] -> (h_a [N,w], a_logits [N,V,JQ]):
    ``attention_3d`` with one sequence and no window."""
    if simiMatrix not in (1, 2, 3, 4):
        print("similarity matrix not implemented")  # model_v2.py:168-170
        raise SystemExit(1)
    if bidirect:
        raise NotImplementedError("bidirect=True (BiDAF baseline, README 4) is outside the FVTA path")
    store = store or default_store()
    N, w = hinfo.shape[0], hinfo.shape[-1]
    h = hinfo.reshape(N, 1, -1, w)
    hm = hinfo_mask.reshape(N, 1, -1) if hinfo_mask is not None else None
    W, b = _att_params(store, scope or "attention_2vector", simiMatrix, w)
    cfg = ops.AttnConfig(simiMatrix, add_tanh, False, 1, 3.0, False)
    g1, a, _, _ = ops.FvtaAttentionFn.apply(cfg, (True, False, False), h, hm, hq, hq_mask, None, W, b,
                                            None, None, None, None)
    return g1, a.reshape(N, -1, hq.shape[1])


def time_indication_func(C, warp_type=1, scope=None, window_t=3.0):
    """model_v2.py:301-341: mask the [N,T,T] correlation by window type -> (C', window_t|None).
    Kept for signature parity and visualisation; the fused kernels never build [N,T,T] (they use
    the row sums ``window_count``)."""
    if warp_type == 1:
        return C, None
    T = C.shape[-1]
    i = torch.arange(T, device=C.device).unsqueeze(1)
    j = torch.arange(T, device=C.device).unsqueeze(0)
    if warp_type == 2:
        keep = i == j
    elif warp_type == 3:
        keep = j <= i
    elif warp_type == 4:
        keep = j >= i
    elif warp_type == 5:
        w = int(math.ceil(float(window_t)))
        keep = (i - j).abs() <= w
    else:
        raise Exception("time warping type not implemented")
    return C * keep.to(C.dtype).unsqueeze(0), (window_t if warp_type == 5 else None)


# ---------------------------------------------------------------------------------------------
# Model
# ---------------------------------------------------------------------------------------------
class Model:
    """model_v2.py:361-1565 (FVTA configuration)."""

    def __init__(self, config, scope, device=None, seed=None):
        if not torch.cuda.is_available():
            raise FvtaError("the FVTA model needs a CUDA device (sm_100a); there is no CPU path")
        from . import _lib
        _lib.load()
        self.scope = scope
        self.config = config
        self.device = torch.device(device or "cuda:%d" % torch.cuda.current_device())
        if not hasattr(config, "num_choice"):
            config.num_choice = 4  # model_v2.py:381 reads it; main.py never defines it
        if getattr(config, "use_bidirection", False):
            raise NotImplementedError("--use_bidirection is shape-broken in model_v2.attention_3d (:286-289)")
        if config.use_time_warp_att and not config.use_time_warp:
            # model_v2.py:919/:270: C is None -> tf.shape(None) raises
            raise ValueError("--use_time_warp_att requires --use_time_warp (C is None otherwise)")
        self.global_step = 0
        self.N = config.batch_size
        self.VW, self.VC, self.W = config.word_vocab_size, config.char_vocab_size, config.max_word_size
        self.cd, self.wd, self.cwd = config.char_emb_size, config.word_emb_size, config.char_out_size
        self.idim = config.image_feat_dim
        self.num_choice = config.num_choice
        seed = getattr(config, "seed", 1234) if seed is None else seed
        self.params = ParamStore(config, seed=seed, device=self.device)
        emb = getattr(config, "emb_mat", None)   # main.py:308
        if emb is not None:
            self.params.view("emb/word/var/word_emb_mat").copy_(torch.as_tensor(emb, dtype=torch.float32))
        # leaf views over the flat buffers: autograd accumulates straight into grad_flat
        self.vars = {}
        for name in self.params.offsets:
            p = self.params.view(name).detach().requires_grad_(True)
            p.grad = self.params.grad(name)
            self.vars[name] = p
        for name in PLACEHOLDERS:          # placeholder attributes double as feed-dict keys
            setattr(self, name, name)
        # visualisation attributes (model_v2.py:384-399)
        self.C = self.C_win = self.att_logits = self.q_att_logits = self.JXP = -1
        self.hat_len = self.had_len = self.hwhen_len = self.hwhere_len = self.hpis_len = self.hpts_len = -1
        self.warp_h = self.hall = -1
        self.logits = self.yp = self.loss = None
        self.summary = None                # tf.summary.merge_all() with no summaries registered
        # window_t never changes (its gradient is None): read it once, not per step
        self.window_t = 3.0
        if "time_warp/time_warp_C/time_warp_window_t" in self.vars:
            self.window_t = float(self.params.view("time_warp/time_warp_C/time_warp_window_t"))

    # ------------------------------------------------------------------ feed handling
    def get_feed_dict(self, batch, is_train=False):
        """model_v2.py:1099-1565: batch -> {placeholder: array}.  Batches that already carry
        tensors (synthetic data, the vectorised input pipeline) pass through unchanged."""
        feed = getattr(batch, "feed", None)
        if feed is None and isinstance(batch, dict):
            feed = batch
        if feed is None:
            from .data import build_feed_dict
            feed = build_feed_dict(self, batch, is_train)
        feed = dict(feed)
        feed["is_train"] = is_train
        return feed

    def to_device(self, feed):
        """Host->device copy of a feed dict (non-blocking when the host tensors are pinned)."""
        out = {}
        for k, v in feed.items():
            if isinstance(v, torch.Tensor):
                out[k] = v.to(self.device, non_blocking=True)
            elif hasattr(v, "dtype") and hasattr(v, "shape"):   # numpy
                out[k] = torch.as_tensor(v).to(self.device, non_blocking=True)
            else:
                out[k] = v
        return out

    # ------------------------------------------------------------------ forward
    def _embed_words(self, ids, ids_c, glove, is_train):
        cfg, v = self.config, self.vars
        VW = self.VW
        ids = ids.long()
        table = v["emb/word/var/word_emb_mat"]
        # word table = concat(trainable [VW], fed GloVe) (model_v2.py:596) without the concat
        # F.embedding, not advanced indexing: its backward is the sorted segmented reduction
        # (plain index backward serialises on the heavily repeated ids)
        emb = torch.nn.functional.embedding
        is_new = ids < VW
        A = torch.where(is_new.unsqueeze(-1), emb(ids.clamp(max=VW - 1), table),
                        emb((ids - VW).clamp(min=0), glove))
        if not cfg.use_char:
            return A
        Ac = emb(ids_c.long(), v["emb/var/char_emb"])                  # [..., J, W, cdim]
        if is_train and cfg.keep_prob < 1.0:                           # conv1d input dropout (:60-61)
            Ac = torch.nn.functional.dropout(Ac, 1.0 - cfg.keep_prob, True)
        filt = v["emb/conv/conv1d/filter"]                             # [1,5,cdim,cwdim]
        height, cdim, cw = filt.shape[1], filt.shape[2], filt.shape[3]
        win = Ac.unfold(-2, height, 1)                                 # [..., J, W-4, cdim, 5]
        lead = win.shape[:-2]
        flat = win.reshape(-1, cdim * height)
        Wm = filt[0].permute(1, 0, 2).reshape(cdim * height, cw)
        xc = torch.relu(torch.matmul(flat, Wm) + v["emb/conv/conv1d/bias"])
        xc = xc.reshape(*lead, cw).amax(dim=-2)                        # max over positions (:65)
        return torch.cat([xc, A], dim=-1)

    def _read(self, x, mask, which):
        v = self.vars
        pre = "reader/text/utext" if which == "text" else "reader/image/uimage"
        kernel, bias = v[pre + "/fw/basic_lstm_cell/kernel"], v[pre + "/fw/basic_lstm_cell/bias"]
        cfg = self.config
        if self._is_train and cfg.keep_prob < 1.0:                     # DropoutWrapper input keep_prob
            x = torch.nn.functional.dropout(x, 1.0 - cfg.keep_prob, True)
        lens = mask.sum(-1).to(torch.int32).reshape(-1).contiguous()
        flat = x.reshape(-1, x.shape[-2], x.shape[-1])
        out, last = ops.bilstm(flat, lens, kernel, bias)
        d2 = out.shape[-1]
        return out.reshape(*x.shape[:-1], d2), last.reshape(*x.shape[:-2], d2)

    def forward(self, feed, is_train=False, want_vis=False):
        """model_v2.py:493-1096 on device tensors.  Sets and returns loss/logits/yp."""
        cfg, v = self.config, self.vars
        f = feed
        self._is_train = is_train
        N = f["q"].shape[0]
        d = cfg.hidden_size
        D = 2 * d
        glove = f["existing_emb_mat"]
        x = {}
        for nme in ("at", "ad", "when", "where", "pts", "q", "choices"):
            x[nme] = self._embed_words(f[nme], f.get(nme + "_c"), glove, is_train)
        # image: transform the batch's unique photos once, then gather (== lookup then linear, :634-645)
        feats = f["image_emb_mat"]
        if cfg.use_image_trans:
            feats = ops.linear(feats, v["emb/image/image_transform/image_trans_linear/W"],
                               v["emb/image/image_transform/image_trans_linear/b"], cfg.add_tanh)
        xpis = torch.nn.functional.embedding(f["pis"].long(), feats)

        hq, lq = self._read(x["q"], f["q_mask"], "text")
        hat, _ = self._read(x["at"], f["at_mask"], "text")
        had, _ = self._read(x["ad"], f["ad_mask"], "text")
        hwhen, _ = self._read(x["when"], f["when_mask"], "text")
        hwhere, _ = self._read(x["where"], f["where_mask"], "text")
        hpts, _ = self._read(x["pts"], f["pts_mask"], "text")
        _, lchoices = self._read(x["choices"], f["choices_mask"], "text")
        hpis, _ = self._read(xpis, f["pis_mask"], "image")

        # context tensor (model_v2.py:862-914)
        M, JXA = hat.shape[1], hat.shape[2]
        JD, JT, JG = had.shape[2], hwhen.shape[2], hwhere.shape[2]
        JI, JXP = hpts.shape[2], hpts.shape[3]
        JMAX = max(JXA, JD, JT, JG, JI * JXP, JI)
        pad = torch.nn.functional.pad
        hs = [pad(hat, (0, 0, 0, JMAX - JXA)), pad(had, (0, 0, 0, JMAX - JD)),
              pad(hwhen, (0, 0, 0, JMAX - JT)), pad(hwhere, (0, 0, 0, JMAX - JG)),
              pad(hpts.reshape(N, M, JI * JXP, D), (0, 0, 0, JMAX - JI * JXP))]
        ms = [pad(f["at_mask"], (0, JMAX - JXA)), pad(f["ad_mask"], (0, JMAX - JD)),
              pad(f["when_mask"], (0, JMAX - JT)), pad(f["where_mask"], (0, JMAX - JG)),
              pad(f["pts_mask"].reshape(N, M, JI * JXP), (0, JMAX - JI * JXP))]
        if not cfg.no_photo:
            hs.append(pad(hpis, (0, 0, 0, JMAX - JI)))
            ms.append(pad(f["pis_mask"], (0, JMAX - JI)))
        K = len(hs)
        hall = torch.stack(hs, dim=1)                  # [N,K,M,JMAX,D]
        hall_mask = torch.stack(ms, dim=1)             # [N,K,M,JMAX]
        T = M * JMAX

        # time warp + attention_3d, fused (model_v2.py:953-1024)
        acfg = ops.AttnConfig(cfg.simiMatrix, cfg.add_tanh, cfg.use_time_warp, cfg.warp_type,
                              self.window_t, cfg.use_time_warp_att)
        tw = cfg.use_time_warp
        g1_all, a_logits, c_col, warp_h = ops.FvtaAttentionFn.apply(
            acfg, (want_vis, want_vis and tw, want_vis and tw),
            hall.reshape(N, K, T, D), hall_mask.reshape(N, K, T), hq, f["q_mask"], lq if tw else None,
            v.get("attention/all/att_logits/W"), v.get("attention/all/att_logits/b"),
            v.get("time_warp/WH/W") if tw else None, v.get("time_warp/WH/b") if tw else None,
            v.get("time_warp/WC/W") if tw else None, v.get("time_warp/WC/b") if tw else None)

        # question attention (model_v2.py:1039-1050)
        q_att = None
        if cfg.use_question_att:
            qcfg = ops.AttnConfig(cfg.simiMatrix, cfg.add_tanh, False, 1, 3.0, False)
            ones = torch.ones(N, 1, dtype=torch.bool, device=hq.device)
            gq, q_att, _, _ = ops.FvtaAttentionFn.apply(
                qcfg, (want_vis, False, False), hq.unsqueeze(1), f["q_mask"].unsqueeze(1),
                g1_all.unsqueeze(1), ones, None,
                v.get("question_emb/question_att/att_logits/W"), v.get("question_emb/question_att/att_logits/b"),
                None, None, None, None)
        else:
            gq = lq

        # output layer (model_v2.py:1053-1083)
        C = lchoices.shape[1]
        g1t = g1_all.unsqueeze(1).expand(N, C, D)
        gqt = gq.unsqueeze(1).expand(N, C, D)
        gc = lchoices
        if cfg.use_eu_output:
            feat = torch.cat([gqt, g1t, gc, g1t * gc, gqt * gc, (g1t - gc) * (g1t - gc), (gqt - gc) * (gqt - gc)], 2)
            logits = ops.linear(feat, v["output/choicelogits/W"], v["output/choicelogits/b"], cfg.add_tanh)
        else:
            feat = torch.cat([gqt, g1t, gc, g1t * gc, gqt * gc], 2)
            logits = ops.linear(feat, v["output/choicelogits/W"], v["output/choicelogits/b"], False)
        logits = logits.squeeze(2)
        self.logits = logits
        self.yp = torch.softmax(logits, dim=-1)
        self.g1_all, self.gq = g1_all, gq
        self.loss = None
        if "y" in f and f["y"] is not None:
            self.loss = self._build_loss(logits, f["y"])
        if want_vis:
            self.hall = hall
            self.att_logits = a_logits
            self.q_att_logits = q_att.reshape(N, -1, 1) if q_att is not None else -1
            self.JXP = JXP
            if tw:
                # model.C [N,T,T] is constant along its last axis as written (:979-980)
                self.C = c_col.unsqueeze(2).expand(N, T, T)
                self.warp_h = warp_h.reshape(N, K, M, JMAX, D)
                self.C_win = self.window_t if cfg.warp_type == 5 else None
        return self.loss, self.logits, self.yp

    def _build_loss(self, logits, y):
        """model_v2.py:1085-1096: mean over the static batch of softmax-CE (+ wd terms)."""
        cfg = self.config
        losses = -(y.to(logits.dtype) * torch.log_softmax(logits, dim=-1)).sum(-1)
        loss = losses.mean()
        if cfg.wd is not None and cfg.wd != 0.0:
            for name, mult in self._wd_names():
                loss = loss + mult * cfg.wd * 0.5 * (self.vars[name] ** 2).sum()
        return loss

    def _wd_names(self):
        """Variables reached by add_wd (model_v2.py:347-354) and how many times: conv1d x7 calls
        (:66), image_trans_linear (:96 via wd=), attention scopes (:293, :196), reader (:834)."""
        out = []
        for n in self.vars:
            if n.startswith("emb/conv/conv1d/"):
                out.append((n, 7))
            elif (n.startswith("emb/image/image_transform/") or n.startswith("reader/")
                  or n.startswith("attention/all/") or n.startswith("question_emb/question_att/")):
                out.append((n, 1))
        return out
