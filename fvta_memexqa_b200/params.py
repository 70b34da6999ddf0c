"""Trainable variables of the FVTA model, laid out in ONE flat fp32 buffer.

Names and shapes are the TensorFlow variables the reference's scopes create (SURVEY
Appendix D; creation sites cited per entry below), so a checkpoint name map is the identity.
Every variable is a view into ``ParamStore.flat``; gradients are views into
``ParamStore.grad_flat`` with the same offsets, which is what lets data parallelism
all-reduce one buffer per step and Adadelta run as one kernel over it
(reference: trainer.py:16-27 applies one optimizer to all trainables).

Initialisers follow the reference: ``linear`` W ~ truncated_normal(stddev=0.1), b = 0
(model_v2.py:88-89); ``tf.get_variable`` with no initializer (char_emb :530, conv filter/bias
:55-56, LSTM kernel) -> glorot_uniform [TF default]; LSTM bias zeros [TF];
word_emb_mat ~ N(0, 1) (main.py:308); window_t = 3 (model_v2.py:333-334).
"""
from collections import OrderedDict
import math

import torch


def att_w_rows(simiMatrix, D):
    """Rows of ``att_logits/W`` per similarity mode (model_v2.py:241-257)."""
    return {1: 3 * D, 2: 2 * D, 3: 4 * D, 4: 0}[simiMatrix]


def param_specs(config):
    """OrderedDict name -> (shape, init kind).  Order = graph construction order."""
    d = config.hidden_size
    D = 2 * d
    wdim = config.word_emb_size
    specs = OrderedDict()
    in_text = wdim
    if config.use_char:
        specs["emb/var/char_emb"] = ((config.char_vocab_size, config.char_emb_size), "glorot")      # :530
        specs["emb/conv/conv1d/filter"] = ((1, 5, config.char_emb_size, config.char_out_size), "glorot_conv")  # :55
        specs["emb/conv/conv1d/bias"] = ((config.char_out_size,), "glorot_vec")                      # :56
        in_text = wdim + config.char_out_size
    specs["emb/word/var/word_emb_mat"] = ((config.word_vocab_size, wdim), "normal")                 # :590
    in_img = config.image_feat_dim
    if config.use_image_trans:
        specs["emb/image/image_transform/image_trans_linear/W"] = ((config.image_feat_dim, config.image_trans_dim), "trunc")  # :645
        specs["emb/image/image_transform/image_trans_linear/b"] = ((config.image_trans_dim,), "zeros")
        in_img = config.image_trans_dim
    specs["reader/text/utext/fw/basic_lstm_cell/kernel"] = ((in_text + d, 4 * d), "glorot")        # :652,:694
    specs["reader/text/utext/fw/basic_lstm_cell/bias"] = ((4 * d,), "zeros")
    specs["reader/image/uimage/fw/basic_lstm_cell/kernel"] = ((in_img + d, 4 * d), "glorot")       # :653,:823
    specs["reader/image/uimage/fw/basic_lstm_cell/bias"] = ((4 * d,), "zeros")
    if config.use_time_warp:
        specs["time_warp/WH/W"] = ((2 * D, D), "trunc")                                             # :986
        specs["time_warp/WH/b"] = ((D,), "zeros")
        specs["time_warp/WC/W"] = ((D, 1), "trunc")                                                 # :989
        specs["time_warp/WC/b"] = ((1,), "zeros")
        if config.warp_type == 5:
            specs["time_warp/time_warp_C/time_warp_window_t"] = ((), "window")                      # :334
    rows = att_w_rows(config.simiMatrix, D)
    if rows:
        specs["attention/all/att_logits/W"] = ((rows, 1), "trunc")                                  # :1020 -> :245
        specs["attention/all/att_logits/b"] = ((1,), "zeros")
        if config.use_question_att:
            specs["question_emb/question_att/att_logits/W"] = ((rows, 1), "trunc")                  # :1044 -> :158
            specs["question_emb/question_att/att_logits/b"] = ((1,), "zeros")
    out_rows = (7 if config.use_eu_output else 5) * D
    specs["output/choicelogits/W"] = ((out_rows, 1), "trunc")                                       # :1073/:1075
    specs["output/choicelogits/b"] = ((1,), "zeros")
    return specs


# ``window_t`` is declared trainable (model_v2.py:334) but only reaches the graph through
# ceil -> cast(int64): its gradient is None and apply_gradients skips it (trainer.py:23-27).
NEVER_UPDATED = ("time_warp/time_warp_C/time_warp_window_t",)

# Variables read through tf.nn.embedding_lookup (model_v2.py:534 char_emb, :596-600 word_emb_mat via
# the axis-0 concat): their gradients are IndexedSlices, so AdadeltaOptimizer applies
# sparse_apply_adadelta [TF] -- rows no id of the batch looked up keep var and both accumulators.
SPARSE_TABLES = ("emb/var/char_emb", "emb/word/var/word_emb_mat")


def _init(shape, kind, gen):
    if kind == "zeros":
        return torch.zeros(shape)
    if kind == "window":
        return torch.tensor(3.0)
    if kind == "normal":
        return torch.randn(shape, generator=gen)
    if kind == "trunc":
        # tf.truncated_normal: resample outside 2 sigma
        t = torch.empty(shape)
        torch.nn.init.trunc_normal_(t, mean=0.0, std=0.1, a=-0.2, b=0.2, generator=gen)
        return t
    if kind == "glorot":
        fan_in, fan_out = shape[0], shape[1]
    elif kind == "glorot_conv":
        rf = shape[0] * shape[1]
        fan_in, fan_out = rf * shape[2], rf * shape[3]
    elif kind == "glorot_vec":
        fan_in = fan_out = shape[0]
    else:
        raise ValueError(kind)
    limit = math.sqrt(6.0 / (fan_in + fan_out))
    return (torch.rand(shape, generator=gen) * 2 - 1) * limit


class ParamStore:
    """All trainable variables as views into one flat buffer (+ a flat gradient buffer and
    the two flat Adadelta accumulators).  Offsets are padded to 4 floats (16 B) so every
    view is float4-aligned for the CUDA kernels."""

    def __init__(self, config, seed=1234, device="cpu", dtype=torch.float32):
        self.specs = param_specs(config)
        self.offsets = OrderedDict()
        off = 0
        for name, (shape, _) in self.specs.items():
            n = 1
            for s in shape:
                n *= s
            self.offsets[name] = (off, n, shape)
            off += (n + 3) // 4 * 4
        self.numel = off
        gen = torch.Generator().manual_seed(seed)
        flat = torch.zeros(off, dtype=torch.float32)
        for name, (shape, kind) in self.specs.items():
            o, n, _ = self.offsets[name]
            flat[o:o + n] = _init(shape, kind, gen).reshape(-1)
        self.flat = flat.to(device=device, dtype=dtype)
        # "row was looked up this step" flags of the sparse tables live in a tail of the gradient
        # buffer, so the data-parallel all-reduce of that ONE buffer also forms their union over ranks
        self.touch_rows = OrderedDict()           # table name -> (first flag, rows)
        tlen = 0
        tindex = torch.full((off,), -1, dtype=torch.int32)
        for name in SPARSE_TABLES:
            if name in self.offsets:
                o, n, shape = self.offsets[name]
                rows, dim = shape
                self.touch_rows[name] = (tlen, rows)
                tindex[o:o + n] = (tlen + torch.arange(rows, dtype=torch.int32)).repeat_interleave(dim)
                tlen += rows
        self.touch_len = (tlen + 3) // 4 * 4
        self.touch_index = tindex.to(device) if tlen else None
        self.grad_buf = torch.zeros(off + self.touch_len, dtype=dtype, device=device)
        self.grad_flat = self.grad_buf[:off]
        self.touch = self.grad_buf[off:] if tlen else None
        self.accum = torch.zeros_like(self.flat)          # Adadelta E[g^2]
        self.accum_update = torch.zeros_like(self.flat)   # Adadelta E[dx^2]
        # 1.0 where Adadelta may touch the slot, 0.0 for never-updated variables / padding
        mask = torch.zeros(off, dtype=torch.float32)
        for name, (o, n, _) in self.offsets.items():
            if name not in NEVER_UPDATED:
                mask[o:o + n] = 1.0
        self.update_mask = mask.to(device=device, dtype=dtype)

    def view(self, name, buf=None):
        o, n, shape = self.offsets[name]
        buf = self.flat if buf is None else buf
        return buf[o:o + n].view(shape)

    def grad(self, name):
        return self.view(name, self.grad_flat)

    def as_dict(self, buf=None):
        return OrderedDict((n, self.view(n, buf)) for n in self.offsets)

    def __contains__(self, name):
        return name in self.offsets

    def state_dict(self):
        return {"flat": self.flat.detach().cpu(), "accum": self.accum.cpu(),
                "accum_update": self.accum_update.cpu(),
                "names": list(self.offsets.keys())}

    def load_state_dict(self, sd):
        assert list(sd["names"]) == list(self.offsets.keys()), "variable set differs from checkpoint"
        self.flat.copy_(sd["flat"])
        self.accum.copy_(sd["accum"])
        self.accum_update.copy_(sd["accum_update"])
