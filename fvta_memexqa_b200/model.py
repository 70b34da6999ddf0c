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
        if hasattr(batch, "step_input"):      # a batch that caches its packed, pinned step input
            return batch.step_input(self, is_train)
        feed = getattr(batch, "feed", None)
        if feed is None and isinstance(batch, dict):
            feed = batch
        if feed is None:
            from .data import build_feed_dict
            feed = build_feed_dict(self, batch, is_train)
        feed = dict(feed)
        feed["is_train"] = is_train
        return feed

    def set_resident_tables(self, image_emb_mat=None, existing_emb_mat=None):
        """Keep the photo-feature table and / or the GloVe matrix on the device for the whole run instead
        of feeding them with every batch (model_v2.py:468-477 feeds both per step: 5 MB of the C2 step
        input is the batch's photo rows).  A feed that omits a table uses the resident one; ``pis`` then
        indexes the resident photo table (pipeline.TokenizedStore.feed(resident_photos=True)), whose rows
        the image transform reads in place (fvta_gemm_gather).  Same arithmetic, same results."""
        if not hasattr(self, "resident"):
            self.resident = {}
        for k, v in (("image_emb_mat", image_emb_mat), ("existing_emb_mat", existing_emb_mat)):
            if v is not None:
                self.resident[k] = torch.as_tensor(v, dtype=torch.float32).to(self.device).contiguous()

    def _table(self, f, name):
        t = f.get(name)
        if t is None:
            t = getattr(self, "resident", {}).get(name)
        if t is None:
            raise KeyError("%s is neither fed nor resident (Model.set_resident_tables)" % name)
        return t

    def host_step_input(self, feed):
        """Everything the device step reads, as ONE pinned host buffer (packing.host_step_input).  A data
        loader builds this once per batch (TensorBatch caches it; pipeline.BatchLoader builds it on worker
        processes); to_device then issues a single copy."""
        from .packing import host_step_input
        return host_step_input(feed, self.config, getattr(self, "pin_text_cap", 0), getattr(self, "pin_img_cap", 0),
                               self.params.touch_rows, self.params.touch_len)

    def to_device(self, feed, pack=None):
        """Host feed (or a cached host_step_input) -> device step input: one non-blocking copy of
        the blob, typed views into it."""
        if "_meta" in feed:
            return feed                      # already a device step input
        from .packing import blob_views
        hsi = feed if "blob" in feed else self.host_step_input(feed)
        dev_blob = hsi["blob"].to(self.device, non_blocking=True)
        out = blob_views(dev_blob, hsi["index"])
        out["_blob"] = dev_blob
        out["_meta"] = hsi["meta"]
        out["_host"] = hsi["host"]
        out["is_train"] = hsi["host"].get("is_train", False)
        return out

    @staticmethod
    def h2d_bytes(dfeed):
        return dfeed["_blob"].numel()

    # ------------------------------------------------------------------ forward
    def _sink(self, *names):
        """Gradient views of the named variables (None entries for absent ones): the custom ops
        accumulate parameter gradients straight into the flat buffer."""
        return tuple(self.params.grad(n) if n in self.params else None for n in names)

    def _embed_tokens(self, ids, ids_c, glove, is_train):
        """model_v2.py:524-620 for the packed valid tokens: word vector (trainable table for ids <
        VW, fed GloVe table otherwise, :596) and the char-CNN (:52-70) concatenated [char, word]."""
        cfg, v = self.config, self.vars
        VW = self.VW
        table = v["emb/word/var/word_emb_mat"]
        if cfg.use_char:
            filt = v["emb/conv/conv1d/filter"]                          # [1,5,cdim,cwdim]
            height, cdim, cw = filt.shape[1], filt.shape[2], filt.shape[3]
            dropped = is_train and cfg.keep_prob < 1.0                  # conv1d input dropout (:60-61)
            table_ok = (not dropped and cw <= 128 and ids_c.shape[1] <= 64
                        and ids_c.shape[1] - height + 1 <= 255)
        if not cfg.use_char or table_ok:
            # a window's conv output depends only on its `height` char ids: tabulate
            # char_emb @ filter[h] once ([height,VC,cw], a few MFLOP) and gather-add (csrc/charcnn.cu);
            # the word vectors land next to it in the same buffer (csrc/embed.cu)
            ctab = ops.CharTableFn.apply(v["emb/var/char_emb"], filt,
                                         self._sink("emb/var/char_emb", "emb/conv/conv1d/filter")) \
                if cfg.use_char else None
            return ops.TokenEmbedFn.apply(table, glove, ids, ctab, v["emb/conv/conv1d/bias"] if cfg.use_char else None,
                                          ids_c if cfg.use_char else None,
                                          self._sink("emb/word/var/word_emb_mat", "emb/conv/conv1d/bias"))
        is_new = ids < VW
        A = torch.where(is_new.unsqueeze(-1),
                        ops.embed(table, ids.clamp(max=VW - 1), self._sink("emb/word/var/word_emb_mat")[0]),
                        ops.embed(glove, (ids - VW).clamp(min=0)))
        Ac = torch.nn.functional.embedding(ids_c.long(), v["emb/var/char_emb"])   # [cap, W, cdim]
        if dropped:
            Ac = torch.nn.functional.dropout(Ac, 1.0 - cfg.keep_prob, True)
        win = Ac.unfold(1, height, 1)                                   # [cap, W-4, cdim, 5]
        npos = win.shape[1]
        flat = win.reshape(-1, cdim * height)
        Wm = filt[0].permute(1, 0, 2).reshape(cdim * height, cw)
        xc = torch.relu(torch.matmul(flat, Wm) + v["emb/conv/conv1d/bias"])
        xc = xc.reshape(-1, npos, cw).amax(dim=1)                       # max over positions (:65)
        return torch.cat([xc, A], dim=-1)

    def forward(self, feed, is_train=False, want_vis=False):
        """model_v2.py:493-1096 on a device step input (see to_device).  Sets and returns
        loss/logits/yp."""
        cfg, v = self.config, self.vars
        f = feed if "_meta" in feed else self.to_device(feed)
        meta = f["_meta"]
        self._is_train = is_train
        N, K, T, M, JMAX, JQ = meta["N"], meta["K"], meta["T"], meta["M"], meta["JMAX"], meta["JQ"]
        d = cfg.hidden_size
        D = 2 * d
        drop = (lambda t: torch.nn.functional.dropout(t, 1.0 - cfg.keep_prob, True)) \
            if (is_train and cfg.keep_prob < 1.0) else (lambda t: t)
        # embeddings of the valid tokens only (padding never reaches a consumer: dynamic_rnn ignores it)
        # The image side (lookup + transform) and the text side (char-CNN + word vectors) meet only in the
        # reader: the image side is enqueued on the reader's second stream (ops.ReaderFn runs the image
        # reader there too), so inside a captured step the two are parallel branches.  Autograd replays
        # each op's backward on the stream its forward ran on, which parallelises the backward the same way.
        main = torch.cuda.current_stream()
        side = ops._reader_side_stream(self.device)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            if cfg.use_image_trans:                                    # model_v2.py:634-645
                # lookup + linear (+tanh) as one product: the A rows are read in place from the fed table
                x_img = ops.gather_linear(self._table(f, "image_emb_mat"), f["img_tok_pidx"],
                                          v["emb/image/image_transform/image_trans_linear/W"],
                                          v["emb/image/image_transform/image_trans_linear/b"], cfg.add_tanh,
                                          *self._sink("emb/image/image_transform/image_trans_linear/W",
                                                      "emb/image/image_transform/image_trans_linear/b"))
            else:
                x_img = ops.embed(self._table(f, "image_emb_mat"), f["img_tok_pidx"])
            x_img = drop(x_img)
        x_text = drop(self._embed_tokens(f["text_tok_ids"], f.get("text_tok_cids"), self._table(f, "existing_emb_mat"), is_train))
        main.wait_stream(side)
        # the eight reader call sites (model_v2.py:694-823) + the context tensor (:862-914)
        arena, last_t, last_i = ops.ReaderFn.apply(
            x_text, x_img, f["text_seq_off"], f["text_seq_len"], f["text_seq_out"],
            f["img_seq_off"], f["img_seq_len"], f["img_seq_out"],
            v["reader/text/utext/fw/basic_lstm_cell/kernel"], v["reader/text/utext/fw/basic_lstm_cell/bias"],
            v["reader/image/uimage/fw/basic_lstm_cell/kernel"], v["reader/image/uimage/fw/basic_lstm_cell/bias"],
            meta["arena_rows"], meta["text_Jmax"], meta["img_Jmax"],
            self._sink("reader/text/utext/fw/basic_lstm_cell/kernel", "reader/text/utext/fw/basic_lstm_cell/bias",
                       "reader/image/uimage/fw/basic_lstm_cell/kernel", "reader/image/uimage/fw/basic_lstm_cell/bias"))
        hall = arena[:meta["hall_rows"]].view(N, K, M, JMAX, D)
        hall_mask = f["hall_mask"].view(N, K, M, JMAX)
        hq = arena[meta["hq_row0"]:].view(N, JQ, D)
        q0, q1 = meta["q_range"]
        c0, c1 = meta["choices_range"]
        lq = last_t[q0:q1]
        lchoices = last_t[c0:c1].view(N, meta["C"], D)
        JXP = meta["JXP"]

        # time warp + attention_3d, fused (model_v2.py:953-1024)
        acfg = ops.AttnConfig(cfg.simiMatrix, cfg.add_tanh, cfg.use_time_warp, cfg.warp_type,
                              self.window_t, cfg.use_time_warp_att)
        tw = cfg.use_time_warp
        # both attentions write their context / question gradients straight into one arena-shaped buffer
        asink = ops.ArenaGradSink(arena, meta["hall_rows"], meta["hq_row0"]) \
            if (torch.is_grad_enabled() and arena.requires_grad) else None
        g1_all, a_logits, c_col, warp_h = ops.FvtaAttentionFn.apply(
            acfg, (want_vis, want_vis and tw, want_vis and tw),
            hall.reshape(N, K, T, D), f["hall_mask"], hq, f["q_mask"], lq if tw else None,
            v.get("attention/all/att_logits/W"), v.get("attention/all/att_logits/b"),
            v.get("time_warp/WH/W") if tw else None, v.get("time_warp/WH/b") if tw else None,
            v.get("time_warp/WC/W") if tw else None, v.get("time_warp/WC/b") if tw else None,
            # fvta_attn_bwd OVERWRITES its parameter gradients: only sink them when nothing else
            # (weight decay) adds to the same slots through autograd
            None if cfg.wd else self._sink("attention/all/att_logits/W", "attention/all/att_logits/b",
                                           "time_warp/WH/W", "time_warp/WH/b", "time_warp/WC/W", "time_warp/WC/b"),
            arena if asink is not None else None, (asink, "context") if asink is not None else None)

        # question attention (model_v2.py:1039-1050)
        q_att = None
        if cfg.use_question_att:
            qcfg = ops.AttnConfig(cfg.simiMatrix, cfg.add_tanh, False, 1, 3.0, False)
            ones = self._ones_mask(N, hq.device)
            gq, q_att, _, _ = ops.FvtaAttentionFn.apply(
                qcfg, (want_vis, False, False), hq.unsqueeze(1), f["q_mask"].unsqueeze(1),
                g1_all.unsqueeze(1), ones, None,
                v.get("question_emb/question_att/att_logits/W"), v.get("question_emb/question_att/att_logits/b"),
                None, None, None, None,
                None if cfg.wd else self._sink("question_emb/question_att/att_logits/W",
                                               "question_emb/question_att/att_logits/b") + (None, None, None, None),
                None, (asink, "question") if asink is not None else None)
        else:
            gq = lq

        # output layer (model_v2.py:1053-1083) + loss (:1085-1096)
        C = lchoices.shape[1]
        yt = f.get("y")
        if cfg.use_eu_output:
            g1t = g1_all.unsqueeze(1).expand(N, C, D)
            gqt = gq.unsqueeze(1).expand(N, C, D)
            gc = lchoices
            feat = torch.cat([gqt, g1t, gc, g1t * gc, gqt * gc, (g1t - gc) * (g1t - gc), (gqt - gc) * (gqt - gc)], 2)
            logits = ops.linear(feat, v["output/choicelogits/W"], v["output/choicelogits/b"], cfg.add_tanh,
                                *self._sink("output/choicelogits/W", "output/choicelogits/b")).squeeze(2)
            self.logits = logits
            self.yp = torch.softmax(logits, dim=-1)
            self.loss = self._build_loss(logits, yt) if yt is not None else None
        else:
            # one launch: logits, yp, mean CE and every gradient of the tail (csrc/output_loss.cu)
            loss, logits, yp = ops.OutputLossFn.apply(
                gq, g1_all, lchoices, v["output/choicelogits/W"], v["output/choicelogits/b"], yt,
                self._sink("output/choicelogits/W", "output/choicelogits/b"), torch.is_grad_enabled())
            self.logits, self.yp = logits, yp
            wdl = self._wd_loss()
            self.loss = (loss if isinstance(wdl, float) and wdl == 0.0 else loss + wdl) if yt is not None else None
        self.g1_all, self.gq = g1_all, gq
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
        return losses.mean() + self._wd_loss()

    def release_graph(self):
        """Drop every reference to the autograd graph of the last forward (call after backward).
        Besides freeing the saved activations early this lets the leaves' AccumulateGrad nodes die
        with the graph, so the next forward -- possibly under CUDA-graph capture, on another stream
        -- creates fresh ones on ITS stream instead of inheriting nodes bound to a stale stream
        (autograd would otherwise order the capture against that uncaptured stream and fail)."""
        for name in ("loss", "g1_all", "gq"):
            t = getattr(self, name, None)
            if isinstance(t, torch.Tensor) and t.grad_fn is not None:
                setattr(self, name, t.detach())

    def _ones_mask(self, N, device):
        """[N,1] all-true mask of the question attention's single 'question word' (cached)."""
        m = getattr(self, "_ones_cache", None)
        if m is None or m.shape[0] != N or m.device != device:
            m = self._ones_cache = torch.ones(N, 1, dtype=torch.bool, device=device)
        return m

    def _wd_loss(self):
        cfg = self.config
        if cfg.wd is None or cfg.wd == 0.0:
            return 0.0
        loss = 0.0
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
