"""Whole-model CPU oracle at shapes where the literal time-warp tensors do not fit.

TEST INFRASTRUCTURE ONLY and PARITY UNPINNED (see ``oracle/tf_ops.py``).

Same graph as ``oracle/literal.forward`` (model_v2.py:493-1096) with the FVTA block computed
by ``oracle/collapsed.py`` -- the two blocks are asserted equal in tests/test_oracle.py.
Also hosts the oracle training step (loss -> autograd gradients -> Adadelta, trainer.py:16-27).
"""
import torch

from . import literal, collapsed
from .tf_ops import adadelta_update


def forward(feed, params, config, is_train=False, block="collapsed"):
    if block == "literal":
        return literal.forward(feed, params, config, is_train)
    x = literal.embed(feed, params, config, is_train)
    r = literal.reader(x, feed, params, config)
    hall, hall_mask = literal.context_tensor(
        r["hat"], r["had"], r["hwhen"], r["hwhere"], r["hpts"], r["hpis"],
        feed["at_mask"], feed["ad_mask"], feed["when_mask"], feed["where_mask"],
        feed["pts_mask"], feed["pis_mask"], no_photo=config.no_photo)
    out = {"hall": hall, "hall_mask": hall_mask, "hq": r["hq"], "lq": r["lq"], "lchoices": r["lchoices"]}
    blk = collapsed.fvta_attention(hall, hall_mask, r["hq"], feed["q_mask"], r["lq"], params, config)
    out["att_logits"], out["g1_all"], out["warp_h"], out["c"] = (
        blk["att_logits"], blk["g1_all"], blk["warp_h"], blk["c"])
    if config.use_question_att:
        gq, qa = collapsed.question_attention(
            r["hq"], blk["g1_all"], feed["q_mask"], config.simiMatrix,
            params.get("question_emb/question_att/att_logits/W"),
            params.get("question_emb/question_att/att_logits/b"), config.add_tanh)
        out["q_att_logits"] = qa
    else:
        gq = r["lq"]
    out["gq"] = gq
    logits, yp = literal.output_layer(gq, blk["g1_all"], r["lchoices"], params, config)
    out["logits"], out["yp"] = logits, yp
    if "y" in feed:
        out["loss"] = literal.build_loss(logits, feed["y"], params, config.wd)
    return out


# variables whose gradient TF holds as IndexedSlices (read through tf.nn.embedding_lookup:
# model_v2.py:534 char_emb, :596-600 word_emb_mat via the axis-0 concat with the fed GloVe matrix)
SPARSE_TABLES = {"emb/var/char_emb": "_c", "emb/word/var/word_emb_mat": ""}
TEXT_INPUTS = ("q", "at", "ad", "when", "where", "pts", "choices")


def looked_up_rows(feed, name, rows):
    """Rows of a sparse table that the FED id arrays address (padding ids included: the reference
    looks up the padded arrays, model_v2.py:534-546, :599-606)."""
    hit = torch.zeros(rows, dtype=torch.bool)
    for site in TEXT_INPUTS:
        ids = feed[site + SPARSE_TABLES[name]].reshape(-1).long()
        hit[ids[ids < rows]] = True
    return hit


def train_step(feed, params, state, config, block="collapsed", never_updated=()):
    """One reference training step (trainer.py:30-40): loss, d loss / d every trainable,
    Adadelta(lr=init_lr).  ``params``/``state`` are dicts of tensors; returns
    (loss, grads, new_params, new_state).  Variables whose gradient is None (window_t) are
    skipped like ``apply_gradients`` skips them.  The two embedding tables get TF's sparse
    Adadelta [TF: AdadeltaOptimizer._apply_sparse -> sparse_apply_adadelta on the de-duplicated
    IndexedSlices]: rows that were not looked up keep var, accum and accum_update."""
    leaves = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    out = forward(feed, leaves, config, is_train=True, block=block)
    loss = out["loss"]
    names = [k for k in leaves if k not in never_updated]
    grads = torch.autograd.grad(loss, [leaves[k] for k in names], allow_unused=True)
    grads = {k: (g if g is not None else None) for k, g in zip(names, grads)}
    new_params, new_state = dict(params), {}
    for k in params:
        acc, acc_u = state.get(k, (torch.zeros_like(params[k]), torch.zeros_like(params[k])))
        g = grads.get(k)
        if g is None:
            new_state[k] = (acc, acc_u)
            continue
        var, acc_n, acc_un = adadelta_update(params[k], g, acc, acc_u, config.init_lr)
        if k in SPARSE_TABLES:
            hit = looked_up_rows(feed, k, params[k].shape[0]).unsqueeze(1)
            var = torch.where(hit, var, params[k])
            acc_n = torch.where(hit, acc_n, acc)
            acc_un = torch.where(hit, acc_un, acc_u)
        new_params[k] = var
        new_state[k] = (acc_n, acc_un)
    return loss.detach(), grads, new_params, new_state, out
