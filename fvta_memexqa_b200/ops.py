"""Host-side operators over the C ABI (include/fvta_b200.h).

Each function takes CUDA torch tensors, checks layout, and passes raw device pointers to
libfvta_b200.so on torch's current stream.  torch supplies device memory, streams and autograd
plumbing only; all arithmetic of the hot path happens in the library's kernels.  There is no
fallback: CPU tensors or a missing library raise.
"""
import ctypes
import math

import torch

from . import _lib
from ._lib import FvtaAttnDesc, FvtaAttnParams, FvtaAttnGrads, FvtaLstmDesc, check


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _req(t, dtype, name):
    if t is None:
        return None
    if not t.is_cuda:
        raise _lib.FvtaError("%s must be a CUDA tensor (the FVTA path has no CPU fallback)" % name)
    if t.dtype != dtype:
        raise _lib.FvtaError("%s must be %s, got %s" % (name, dtype, t.dtype))
    return t.contiguous()


def _mask_u8(m, name):
    if m is None:
        return None
    if not m.is_cuda:
        raise _lib.FvtaError("%s must be a CUDA tensor" % name)
    if m.dtype == torch.bool:
        return m.contiguous().view(torch.uint8)
    return m.to(torch.uint8).contiguous()


class AttnConfig:
    """The static part of an attention call (the reference's keyword arguments)."""

    def __init__(self, simiMatrix=1, add_tanh=False, use_time_warp=False, warp_type=1, window_t=3.0,
                 time_warp_att=False):
        self.simiMatrix = int(simiMatrix)
        self.add_tanh = bool(add_tanh)
        self.use_time_warp = bool(use_time_warp)
        self.warp_type = int(warp_type)
        self.window = int(math.ceil(float(window_t)))  # model_v2.py:335 ceil -> int64
        self.time_warp_att = bool(time_warp_att)

    def desc(self, N, K, T, D, JQ, has_mask, path=0):
        return FvtaAttnDesc(N=N, K=K, T=T, D=D, JQ=JQ, simi=self.simiMatrix, add_tanh=int(self.add_tanh),
                            use_time_warp=int(self.use_time_warp), warp_type=self.warp_type,
                            window=self.window, time_warp_att=int(self.time_warp_att),
                            has_mask=int(has_mask), ws_from_fwd=0, path=int(path))


def _attn_params(W, b, WH_W, WH_b, WC_W, WC_b):
    return FvtaAttnParams(att_W=_ptr(W), att_b=_ptr(b), WH_W=_ptr(WH_W), WH_b=_ptr(WH_b),
                          WC_W=_ptr(WC_W), WC_b=_ptr(WC_b))


def attn_forward(cfg, hall, hall_mask, hq, hq_mask, lq, W, b, WH_W=None, WH_b=None, WC_W=None, WC_b=None,
                 want_logits=False, want_c=False, want_warp_h=False, reuse=None, path=0, folded=False):
    """fvta_attn_fwd.  hall [N,K,T,D], hq [N,JQ,D] -> dict(g1_all, a_logits, c_col, warp_h, stats, path).
    path: _lib.PATH_AUTO, or _lib.PATH_FFMA to pin the CUDA-core tile kernel (A/B parity checks).
    folded=True (with reuse=): the workspace of the earlier call still holds the folded question
    operands for the same hq / lq / parameters (FvtaAttnDesc.ws_from_fwd), only the context is new."""
    lib = _lib.load()
    f32 = torch.float32
    hall = _req(hall, f32, "hall")
    hq = _req(hq, f32, "hq")
    lq = _req(lq, f32, "lq")
    W, b = _req(W, f32, "att_W"), _req(b, f32, "att_b")
    WH_W, WH_b, WC_W, WC_b = (_req(x, f32, "time_warp param") for x in (WH_W, WH_b, WC_W, WC_b))
    N, K, T, D = hall.shape
    JQ = hq.shape[1]
    has_mask = hall_mask is not None and hq_mask is not None  # model_v2.py:232, :263
    hm = _mask_u8(hall_mask, "hall_mask") if has_mask else None
    qm = _mask_u8(hq_mask, "hq_mask") if has_mask else None
    desc = cfg.desc(N, K, T, D, JQ, has_mask, path)
    dev = hall.device
    ws_bytes = lib.fvta_attn_workspace_bytes(ctypes.byref(desc))
    if ws_bytes == 0:
        raise _lib.FvtaError("fvta_attn: unsupported shape/flags N=%d K=%d T=%d D=%d JQ=%d simi=%d"
                             % (N, K, T, D, JQ, cfg.simiMatrix))
    if reuse is not None:     # benchmark loops: reuse the buffers of an earlier call (no allocator traffic)
        ws, stats, g1_all, a_logits = reuse["ws"], reuse["stats"], reuse["g1_all"], reuse["a_logits"]
    else:
        ws = torch.empty(ws_bytes // 4, dtype=f32, device=dev)
        stats = torch.empty(lib.fvta_attn_stats_floats(ctypes.byref(desc)), dtype=f32, device=dev)
        g1_all = torch.empty(N, D, dtype=f32, device=dev)
        a_logits = torch.empty(N, K, T, JQ, dtype=f32, device=dev) if want_logits else None
    c_col = torch.empty(N, T, dtype=f32, device=dev) if (want_c and cfg.use_time_warp) else None
    warp_h = torch.empty(N, K, T, D, dtype=f32, device=dev) if (want_warp_h and cfg.use_time_warp) else None
    params = _attn_params(W, b, WH_W, WH_b, WC_W, WC_b)
    ran = lib.fvta_attn_fwd_path(ctypes.byref(desc), int(c_col is not None), int(warp_h is not None))
    if folded:
        if reuse is None:
            raise _lib.FvtaError("folded=True needs the buffers of the earlier call (reuse=)")
        desc.ws_from_fwd = 1
    rc = lib.fvta_attn_fwd(ctypes.byref(desc), _ptr(hall), _ptr(hm), _ptr(hq), _ptr(qm), _ptr(lq),
                           ctypes.byref(params), _ptr(g1_all), _ptr(a_logits), _ptr(c_col), _ptr(warp_h),
                           _ptr(stats), _ptr(ws), ws_bytes, _stream())
    check(rc, "fvta_attn_fwd")
    desc.ws_from_fwd = 0
    return dict(g1_all=g1_all, a_logits=a_logits, c_col=c_col, warp_h=warp_h, stats=stats, ws=ws,
                desc=desc, saved=(hall, hm, hq, qm, lq), path=ran)


def attn_backward(fwd, d_g1_all, W, b, WH_W=None, WH_b=None, WC_W=None, WC_b=None, sinks=None,
                  d_hall_out=None, d_hq_out=None, path=None):
    """fvta_attn_bwd given the dict attn_forward returned.  Returns dict of gradients.  `sinks`:
    optional tensors the six parameter gradients are written into (the entry point overwrites).
    d_hall_out / d_hq_out: optional contiguous tensors to receive d_hall / d_hq (overwritten)."""
    lib = _lib.load()
    f32 = torch.float32
    hall, hm, hq, qm, lq = fwd["saved"]
    desc = fwd["desc"]
    d_g1_all = _req(d_g1_all, f32, "d_g1_all")
    dev = hall.device
    for t, ref, nme in ((d_hall_out, hall, "d_hall_out"), (d_hq_out, hq, "d_hq_out")):
        if t is not None and (t.numel() != ref.numel() or not t.is_contiguous() or t.dtype != f32):
            raise _lib.FvtaError("%s must be a contiguous fp32 tensor of %d elements" % (nme, ref.numel()))
    out = dict(d_hall=d_hall_out if d_hall_out is not None else torch.empty_like(hall),
               d_hq=d_hq_out if d_hq_out is not None else torch.empty_like(hq),
               d_lq=torch.empty_like(lq) if (lq is not None and desc.use_time_warp) else None)
    if sinks is not None:
        gW, gb = sinks[0], sinks[1]
        gs = [sk if desc.use_time_warp else None for sk in sinks[2:6]]
    else:
        gW = torch.empty_like(W) if W is not None else None
        gb = torch.empty_like(b) if b is not None else None
        gs = [torch.empty_like(x) if (x is not None and desc.use_time_warp) else None
              for x in (WH_W, WH_b, WC_W, WC_b)]
    grads = FvtaAttnGrads(att_W=_ptr(gW), att_b=_ptr(gb), WH_W=_ptr(gs[0]), WH_b=_ptr(gs[1]),
                          WC_W=_ptr(gs[2]), WC_b=_ptr(gs[3]))
    params = _attn_params(W, b, WH_W, WH_b, WC_W, WC_b)
    ws = fwd["ws"]
    # the forward's workspace is kept alive in `fwd`: reuse its folded operands (2: and the tensor-core images)
    desc.ws_from_fwd = 2 if fwd.get("path") == _lib.PATH_TC else 1
    if path is not None:        # A/B: _lib.PATH_FFMA pins the CUDA-core backward tile kernel
        desc.path = int(path)
    rc = lib.fvta_attn_bwd(ctypes.byref(desc), _ptr(hall), _ptr(hm), _ptr(hq), _ptr(qm), _ptr(lq),
                           ctypes.byref(params), _ptr(fwd["stats"]), _ptr(d_g1_all), _ptr(out["d_hall"]),
                           _ptr(out["d_hq"]), _ptr(out["d_lq"]), ctypes.byref(grads), _ptr(ws),
                           ws.numel() * 4, _stream())
    check(rc, "fvta_attn_bwd")
    out.update(d_att_W=gW, d_att_b=gb, d_WH_W=gs[0], d_WH_b=gs[1], d_WC_W=gs[2], d_WC_b=gs[3])
    return out


def adadelta_step(var, grad, accum, accum_update, update_mask, lr, rho=0.95, eps=1e-8, grad_scale=1.0,
                  touch_index=None, touch=None):
    """fvta_adadelta over flat buffers (in place).  touch_index / touch: sparse rows of the embedding
    tables (tf.nn.embedding_lookup -> IndexedSlices -> sparse_apply_adadelta; see the header)."""
    lib = _lib.load()
    f32 = torch.float32
    for t, nme in ((var, "var"), (grad, "grad"), (accum, "accum"), (accum_update, "accum_update")):
        if _req(t, f32, nme) is not t:
            raise _lib.FvtaError("%s must be contiguous" % nme)
    if touch_index is not None:
        touch_index = _req(touch_index, torch.int32, "touch_index")
        touch = _req(touch, f32, "touch")
        if touch_index.numel() != var.numel():
            raise _lib.FvtaError("touch_index must have one entry per parameter")
    rc = lib.fvta_adadelta(var.numel(), _ptr(var), _ptr(grad), _ptr(accum), _ptr(accum_update),
                           _ptr(update_mask), _ptr(touch_index), _ptr(touch), lr, rho, eps, grad_scale, _stream())
    check(rc, "fvta_adadelta")


# ---------------------------------------------------------------------------------------------
# GEMM / linear
# ---------------------------------------------------------------------------------------------
ACT = {None: 0, "none": 0, "tanh": 1, "relu": 2}


def gemm(A, B, bias=None, act=None, trans_a=False, trans_b=False, out=None, accumulate=False):
    """fvta_gemm: act(op(A) @ op(B) + bias); A, B 2-D contiguous fp32 CUDA tensors."""
    lib = _lib.load()
    f32 = torch.float32
    A, B, bias = _req(A, f32, "A"), _req(B, f32, "B"), _req(bias, f32, "bias")
    M, K = (A.shape[1], A.shape[0]) if trans_a else (A.shape[0], A.shape[1])
    K2, N = (B.shape[1], B.shape[0]) if trans_b else (B.shape[0], B.shape[1])
    if K != K2:
        raise _lib.FvtaError("gemm: inner dimensions differ (%d vs %d)" % (K, K2))
    if out is None:
        out = torch.empty(M, N, dtype=f32, device=A.device)
    rc = lib.fvta_gemm(M, N, K, _ptr(A), A.shape[1], int(trans_a), _ptr(B), B.shape[1], int(trans_b),
                       _ptr(bias), ACT[act], _ptr(out), out.shape[1], int(accumulate), _stream())
    check(rc, "fvta_gemm")
    return out


class LinearFn(torch.autograd.Function):
    """`linear` of the reference (model_v2.py:75-100): flatten, x @ W + b, optional tanh."""

    @staticmethod
    def forward(ctx, x, W, b, add_tanh, gW=None, gb=None):
        """gW / gb: optional gradient sinks (views of the flat gradient buffer).  When given, the
        backward accumulates into them directly and reports no gradient to autograd, which
        saves the zero-fill + add that AccumulateGrad would launch per parameter."""
        flat = x.reshape(-1, x.shape[-1])
        y = gemm(flat, W, b, "tanh" if add_tanh else None)
        ctx.save_for_backward(flat, W, y)
        ctx.add_tanh = add_tanh
        ctx.xshape = x.shape
        ctx.sinks = (gW, gb)
        return y.reshape(*x.shape[:-1], W.shape[1])

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        flat, W, y = ctx.saved_tensors
        N = W.shape[1]
        dy = _req(dy.reshape(-1, N), torch.float32, "dy")
        gW, gb = ctx.sinks
        db = gb if gb is not None else torch.zeros(N, dtype=dy.dtype, device=dy.device)
        # d z = d y * (1 - y^2) and d b += column sums, one launch (csrc/embed.cu)
        dz = torch.empty_like(dy) if ctx.add_tanh else dy
        rc = lib.fvta_bias_act_grad(dy.shape[0], N, _ptr(dy), N, _ptr(y) if ctx.add_tanh else None, N,
                                    1 if ctx.add_tanh else 0, _ptr(dz) if ctx.add_tanh else None, N, _ptr(db),
                                    _stream())
        check(rc, "fvta_bias_act_grad")
        # the image features and other fed inputs need no gradient: skip that GEMM
        dx = gemm(dz, W, trans_b=True).reshape(ctx.xshape) if ctx.needs_input_grad[0] else None
        if gW is not None:
            gemm(flat, dz, trans_a=True, out=gW, accumulate=True)
            return dx, None, None, None, None, None
        dW = gemm(flat, dz, trans_a=True)
        return dx, dW, db, None, None, None


def linear(x, W, b, add_tanh=False, gW=None, gb=None):
    return LinearFn.apply(x, W, b, add_tanh, gW, gb)


def gemm_gather(table, rows, B, bias=None, act=None, trans_a=False, out=None, accumulate=False):
    """fvta_gemm_gather: act(op(table[rows]) @ B + bias) without materialising table[rows]."""
    lib = _lib.load()
    f32 = torch.float32
    table, B, bias = _req(table, f32, "table"), _req(B, f32, "B"), _req(bias, f32, "bias")
    rows = _req(rows, torch.int64, "rows")
    n_rows, width = rows.numel(), table.shape[1]
    M, K = (width, n_rows) if trans_a else (n_rows, width)
    if B.shape[0] != K:
        raise _lib.FvtaError("gemm_gather: inner dimensions differ (%d vs %d)" % (K, B.shape[0]))
    N = B.shape[1]
    if out is None:
        out = torch.empty(M, N, dtype=f32, device=table.device)
    rc = lib.fvta_gemm_gather(M, N, K, _ptr(table), width, _ptr(rows), int(trans_a), _ptr(B), N, 0, _ptr(bias),
                              ACT[act], _ptr(out), out.shape[1], int(accumulate), _stream())
    check(rc, "fvta_gemm_gather")
    return out


class GatherLinearFn(torch.autograd.Function):
    """tf.nn.embedding_lookup(image_emb_mat, pis) -> `linear` (+tanh) of model_v2.py:634-645 as ONE
    product whose A rows are read in place from the fed photo-feature table (fvta_gemm_gather): the
    gathered [tokens, image_feat_dim] matrix never exists, forward or backward.  The table is fed, not
    trained: no gradient flows to it."""

    @staticmethod
    def forward(ctx, table, rows, W, b, add_tanh, gW=None, gb=None):
        y = gemm_gather(table, rows, W, b, "tanh" if add_tanh else None)
        ctx.save_for_backward(table, rows, y)
        ctx.add_tanh = add_tanh
        ctx.wshape = W.shape
        ctx.sinks = (gW, gb)
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        table, rows, y = ctx.saved_tensors
        N = ctx.wshape[1]
        dy = _req(dy, torch.float32, "dy")
        gW, gb = ctx.sinks
        db = gb if gb is not None else torch.zeros(N, dtype=dy.dtype, device=dy.device)
        dz = torch.empty_like(dy) if ctx.add_tanh else dy
        rc = lib.fvta_bias_act_grad(dy.shape[0], N, _ptr(dy), N, _ptr(y) if ctx.add_tanh else None, N,
                                    1 if ctx.add_tanh else 0, _ptr(dz) if ctx.add_tanh else None, N, _ptr(db),
                                    _stream())
        check(rc, "fvta_bias_act_grad")
        if gW is not None:
            gemm_gather(table, rows, dz, trans_a=True, out=gW, accumulate=True)
            return None, None, None, None, None, None, None
        dW = gemm_gather(table, rows, dz, trans_a=True)
        return None, None, dW, db, None, None, None


def gather_linear(table, rows, W, b, add_tanh=False, gW=None, gb=None):
    return GatherLinearFn.apply(table, rows, W, b, add_tanh, gW, gb)


# ---------------------------------------------------------------------------------------------
# BiLSTM reader
# ---------------------------------------------------------------------------------------------
def _lstm_desc(x, kernel, path=0):
    B, J, In = x.shape
    d = kernel.shape[1] // 4
    if kernel.shape[0] != In + d:
        raise _lib.FvtaError("bilstm: kernel rows %d != In + d = %d" % (kernel.shape[0], In + d))
    return FvtaLstmDesc(B=B, J=J, In=In, d=d, path=path)


def bilstm_forward(x, lens, kernel_fw, bias_fw, kernel_bw=None, bias_bw=None, path=0):
    """fvta_bilstm_fwd.  x [B,J,In] fp32, lens [B] int32.  Passing no bw weights shares fw's
    (the reference hands the same cell object to both directions)."""
    lib = _lib.load()
    f32 = torch.float32
    x = _req(x, f32, "x")
    lens = _req(lens, torch.int32, "lens")
    kernel_fw, bias_fw = _req(kernel_fw, f32, "kernel_fw"), _req(bias_fw, f32, "bias_fw")
    if kernel_bw is None:
        kernel_bw, bias_bw = kernel_fw, bias_fw
    else:
        kernel_bw, bias_bw = _req(kernel_bw, f32, "kernel_bw"), _req(bias_bw, f32, "bias_bw")
    desc = _lstm_desc(x, kernel_fw, path)      # path: _lib.LSTM_AUTO / LSTM_FFMA / LSTM_TC (forward recurrence)
    B, J, d = desc.B, desc.J, desc.d
    dev = x.device
    out = torch.empty(B, J, 2 * d, dtype=f32, device=dev)
    last = torch.empty(B, 2 * d, dtype=f32, device=dev)
    if B == 0 or J == 0:
        return out.zero_(), last.zero_(), None, desc
    reserve = torch.empty(lib.fvta_bilstm_reserve_floats(ctypes.byref(desc)), dtype=f32, device=dev)
    ws_bytes = lib.fvta_bilstm_workspace_bytes(ctypes.byref(desc))
    ws = torch.empty(ws_bytes // 4, dtype=f32, device=dev)
    rc = lib.fvta_bilstm_fwd(ctypes.byref(desc), _ptr(x), _ptr(lens), _ptr(kernel_fw), _ptr(bias_fw),
                             _ptr(kernel_bw), _ptr(bias_bw), _ptr(out), _ptr(last), _ptr(reserve), _ptr(ws),
                             ws_bytes, _stream())
    check(rc, "fvta_bilstm_fwd")
    return out, last, reserve, desc


class BiLSTMFn(torch.autograd.Function):
    """bidirectional_dynamic_rnn(cell, cell, x, sequence_length) with shared fw/bw weights
    (model_v2.py:694 etc.).  Returns (outputs [B,J,2d], final h [B,2d])."""

    @staticmethod
    def forward(ctx, x, lens, kernel, bias):
        out, last, reserve, desc = bilstm_forward(x, lens, kernel, bias)
        ctx.save_for_backward(x, lens, kernel, out, reserve)
        ctx.desc = desc
        ctx.mark_non_differentiable(lens)
        return out, last

    @staticmethod
    def backward(ctx, d_out, d_last):
        lib = _lib.load()
        x, lens, kernel, out, reserve = ctx.saved_tensors
        desc = ctx.desc
        f32 = torch.float32
        d_x = torch.empty_like(x)
        d_kernel = torch.zeros_like(kernel)
        d_bias = torch.zeros(kernel.shape[1], dtype=f32, device=x.device)
        if reserve is None:
            return d_x.zero_(), None, d_kernel, d_bias
        d_out = _req(d_out, f32, "d_out") if d_out is not None else None
        d_last = _req(d_last, f32, "d_last") if d_last is not None else None
        ws_bytes = lib.fvta_bilstm_workspace_bytes(ctypes.byref(desc))
        ws = torch.empty(ws_bytes // 4, dtype=f32, device=x.device)
        xc = x.contiguous()
        rc = lib.fvta_bilstm_bwd(ctypes.byref(desc), _ptr(xc), _ptr(lens), _ptr(kernel), _ptr(kernel), _ptr(out),
                                 _ptr(reserve), _ptr(d_out), _ptr(d_last), _ptr(d_x), _ptr(d_kernel),
                                 _ptr(d_bias), _ptr(d_kernel), _ptr(d_bias), _ptr(ws), ws_bytes, _stream())
        check(rc, "fvta_bilstm_bwd")
        return d_x, None, d_kernel, d_bias


def bilstm(x, lens, kernel, bias):
    return BiLSTMFn.apply(x, lens, kernel, bias)


# ---------------------------------------------------------------------------------------------
# attention as an autograd op
# ---------------------------------------------------------------------------------------------
class ArenaGradSink:
    """Gradient buffer of the context-tensor arena [hall rows | hq rows] x D that the attention
    backwards write into directly.  Without it autograd materialises one zero-filled arena per
    slice (hall, hq) and adds them -- three arena-sized launches per step for nothing."""

    def __init__(self, arena, hall_rows, hq_row0):
        self.buf = torch.empty_like(arena)
        self.hall_rows, self.hq_row0 = hall_rows, hq_row0
        self.hq_written = False

    def hall(self):
        return self.buf[:self.hall_rows]

    def hq(self):
        return self.buf[self.hq_row0:]


class FvtaAttentionFn(torch.autograd.Function):
    """Time-warp gate + attention_3d (model_v2.py:953-1024) as one differentiable op.
    Inputs: hall [N,K,T,D], hall_mask, hq, hq_mask, lq, W, b, WH_W, WH_b, WC_W, WC_b.
    Outputs: g1_all [N,D] (differentiable), a_logits / c_col / warp_h (visualisation; no grad).

    Optional: `sinks` (gradient views for the six parameters, written in place), and
    `asink = (ArenaGradSink, role)` with `anchor` = the arena tensor when hall / hq are slices of
    the reader's arena:
      role "context":  hall = arena hall rows, hq = arena hq rows.  d_hall goes straight into the
                       sink, d_hq is written (or added, if the question attention got there first)
                       into its hq rows, and the whole sink is reported as the gradient of `anchor`;
      role "question": hall = arena hq rows (question attention, model_v2.py:1044).  d_hall is
                       written into the sink's hq rows; nothing is reported for them here (the
                       context op, which autograd necessarily runs later, reports the sink)."""

    @staticmethod
    def forward(ctx, cfg, want, hall, hall_mask, hq, hq_mask, lq, W, b, WH_W, WH_b, WC_W, WC_b, sinks=None,
                anchor=None, asink=None):
        fwd = attn_forward(cfg, hall, hall_mask, hq, hq_mask, lq, W, b, WH_W, WH_b, WC_W, WC_b,
                           want_logits=want[0], want_c=want[1], want_warp_h=want[2])
        ctx.fwd = fwd
        ctx.params = (W, b, WH_W, WH_b, WC_W, WC_b)
        ctx.has_lq = lq is not None
        ctx.sinks = sinks       # gradient views for (W, b, WH_W, WH_b, WC_W, WC_b): written in place
        ctx.asink = asink
        outs = [fwd["g1_all"]]
        for key in ("a_logits", "c_col", "warp_h"):
            t = fwd[key] if fwd[key] is not None else torch.empty(0, device=hall.device)
            ctx.mark_non_differentiable(t)
            outs.append(t)
        return tuple(outs)

    @staticmethod
    def backward(ctx, d_g1, *_unused):
        W, b, WH_W, WH_b, WC_W, WC_b = ctx.params
        d_hall_out = d_hq_out = None
        sink = role = None
        add_hq = False
        if ctx.asink is not None:
            sink, role = ctx.asink
            if role == "context":
                d_hall_out = sink.hall()
                add_hq = sink.hq_written
                d_hq_out = None if add_hq else sink.hq()
            else:
                d_hall_out = sink.hq()
        g = attn_backward(ctx.fwd, d_g1, W, b, WH_W, WH_b, WC_W, WC_b, sinks=ctx.sinks,
                          d_hall_out=d_hall_out, d_hq_out=d_hq_out)
        ctx.fwd = None
        d_hall, d_hq, d_anchor = g["d_hall"], g["d_hq"], None
        if sink is not None:
            d_hall = None
            if role == "context":
                if add_hq:
                    sink.hq().add_(d_hq.reshape(sink.hq().shape))
                d_hq = None
                d_anchor = sink.buf
            else:
                sink.hq_written = True
        d_lq = g["d_lq"] if ctx.has_lq else None
        if ctx.sinks is not None:
            return (None, None, d_hall, None, d_hq, None, d_lq, None, None, None, None, None, None, None,
                    d_anchor, None)
        return (None, None, d_hall, None, d_hq, None, d_lq,
                g["d_att_W"], g["d_att_b"], g["d_WH_W"], g["d_WH_b"], g["d_WC_W"], g["d_WC_b"], None,
                d_anchor, None)


# ---------------------------------------------------------------------------------------------
# packed reader: every text call site in one launch, outputs written in place into the arena
# ---------------------------------------------------------------------------------------------
_READER_SIDE = {}


def _reader_side_stream(device):
    """The stream the image reader runs on, next to the text reader (one per device, created once --
    before any CUDA-graph capture, by the warm-up steps)."""
    key = torch.device(device).index
    if key not in _READER_SIDE:
        _READER_SIDE[key] = torch.cuda.Stream(device=device)
    return _READER_SIDE[key]


def _reader_alloc(lib, x_tok, lens, kernel, Jmax):
    from ._lib import FvtaReaderDesc
    f32 = torch.float32
    d = kernel.shape[1] // 4
    desc = FvtaReaderDesc(B=lens.numel(), ntok=x_tok.shape[0], In=x_tok.shape[1], d=d, Jmax=Jmax, path=0)
    dev = x_tok.device
    last = torch.empty(lens.numel(), 2 * d, dtype=f32, device=dev)
    reserve = torch.empty(lib.fvta_reader_reserve_floats(ctypes.byref(desc)), dtype=f32, device=dev)
    ws_bytes = lib.fvta_reader_workspace_bytes(ctypes.byref(desc))
    ws = torch.empty(ws_bytes // 4, dtype=f32, device=dev)
    return desc, last, reserve, ws, ws_bytes


def _reader_launch(lib, bufs, x_tok, off, lens, outrow, kernel, bias, arena):
    desc, last, reserve, ws, ws_bytes = bufs
    rc = lib.fvta_reader_fwd(ctypes.byref(desc), _ptr(x_tok), _ptr(off), _ptr(lens), _ptr(outrow), _ptr(kernel),
                             _ptr(bias), _ptr(arena), _ptr(last), _ptr(reserve), _ptr(ws), ws_bytes, _stream())
    check(rc, "fvta_reader_fwd")


class ReaderFn(torch.autograd.Function):
    """The eight bidirectional_dynamic_rnn call sites of model_v2.py:694-823 as two launches
    (text cell, image cell) over packed tokens, writing into one zeroed arena
    [hall rows | hq rows] x 2d.  Returns (arena, last_text [Bt,2d], last_img [Bi,2d])."""

    @staticmethod
    def forward(ctx, x_text, x_img, t_off, t_len, t_out, i_off, i_len, i_out, kt, bt, ki, bi, arena_rows,
                Jmax_t, Jmax_i, sinks=None):
        lib = _lib.load()
        f32 = torch.float32
        x_text, x_img = _req(x_text, f32, "x_text"), _req(x_img, f32, "x_img")
        kt, bt, ki, bi = (_req(v, f32, "reader weights") for v in (kt, bt, ki, bi))
        d = kt.shape[1] // 4
        arena = torch.zeros(arena_rows, 2 * d, dtype=f32, device=x_text.device)
        # The two readers are independent (different cells, different arena rows) and both latency-bound
        # -- a recurrence is ~2 us per step on a handful of CTAs -- so the image reader runs on a second
        # stream under the text reader (inside a captured step: a parallel branch of the graph).  Every
        # buffer is allocated on the calling stream, which waits for the side stream before returning.
        bt_ = _reader_alloc(lib, x_text, t_len, kt, Jmax_t)
        bi_ = _reader_alloc(lib, x_img, i_len, ki, Jmax_i)
        main, side = torch.cuda.current_stream(), _reader_side_stream(x_text.device)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            _reader_launch(lib, bi_, x_img, i_off, i_len, i_out, ki, bi, arena)
        _reader_launch(lib, bt_, x_text, t_off, t_len, t_out, kt, bt, arena)
        main.wait_stream(side)
        (desc_t, last_t, res_t), (desc_i, last_i, res_i) = bt_[:3], bi_[:3]
        ctx.save_for_backward(x_text, x_img, t_off, t_len, t_out, i_off, i_len, i_out, kt, ki, res_t, res_i)
        ctx.descs = (desc_t, desc_i)
        ctx.sinks = sinks       # (g kernel_t, g bias_t, g kernel_i, g bias_i) views of the flat gradient
        ctx.set_materialize_grads(False)   # unused final states arrive as None, not as zero tensors
        return arena, last_t, last_i

    @staticmethod
    def backward(ctx, d_arena, d_last_t, d_last_i):
        lib = _lib.load()
        f32 = torch.float32
        x_text, x_img, t_off, t_len, t_out, i_off, i_len, i_out, kt, ki, res_t, res_i = ctx.saved_tensors
        desc_t, desc_i = ctx.descs
        d_arena = _req(d_arena, f32, "d_arena") if d_arena is not None else None
        outs = []
        sinks = ctx.sinks
        main, side = torch.cuda.current_stream(), _reader_side_stream(x_text.device)
        calls = []
        for j, (x, off, ln, orow, k, res, desc, dl) in enumerate(
                ((x_text, t_off, t_len, t_out, kt, res_t, desc_t, d_last_t),
                 (x_img, i_off, i_len, i_out, ki, res_i, desc_i, d_last_i))):
            dx = torch.empty_like(x)
            if sinks is not None:       # the C entry point accumulates (+=): hand it the flat gradient
                dk, db = sinks[2 * j], sinks[2 * j + 1]
            else:
                dk = torch.zeros_like(k)
                db = torch.zeros(k.shape[1], dtype=f32, device=x.device)
            dl = _req(dl, f32, "d_last") if dl is not None else None
            ws_bytes = lib.fvta_reader_workspace_bytes(ctypes.byref(desc))
            ws = torch.empty(ws_bytes // 4, dtype=f32, device=x.device)
            calls.append((desc, x, off, ln, orow, k, res, dl, dx, dk, db, ws, ws_bytes))
            outs.append((dx, dk, db))

        def launch(c):
            desc, x, off, ln, orow, k, res, dl, dx, dk, db, ws, ws_bytes = c
            rc = lib.fvta_reader_bwd(ctypes.byref(desc), _ptr(x), _ptr(off), _ptr(ln), _ptr(orow), _ptr(k), _ptr(res),
                                     _ptr(d_arena), _ptr(dl), _ptr(dx), _ptr(dk), _ptr(db), _ptr(ws), ws_bytes,
                                     _stream())
            check(rc, "fvta_reader_bwd")
        # image reader backward on the side stream, text reader backward on the calling one (see forward)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            launch(calls[1])
        launch(calls[0])
        main.wait_stream(side)
        (dxt, dkt, dbt), (dxi, dki, dbi) = outs
        if not ctx.needs_input_grad[1]:
            dxi = None
        if sinks is not None:
            dkt = dbt = dki = dbi = None
        return (dxt, dxi, None, None, None, None, None, None, dkt, dbt, dki, dbi, None, None, None, None)


# ---------------------------------------------------------------------------------------------
# character CNN through the per-character table
# ---------------------------------------------------------------------------------------------
class CharCnnTableFn(torch.autograd.Function):
    """y[w,:] = relu(max_p(bias + sum_h table[ids[w,p+h]][h]))  (fvta_charcnn_fwd/bwd).
    table [VC,height,cw] = char_emb @ filter[h] comes from CharTableFn (one small product), which
    carries d table back to char_emb and the filter."""

    @staticmethod
    def forward(ctx, table, bias, ids, bias_sink=None):
        ctx.bias_sink = bias_sink
        lib = _lib.load()
        f32 = torch.float32
        table, bias = _req(table, f32, "table"), _req(bias, f32, "bias")
        ids = _req(ids, torch.int32, "char ids")
        VC, H, cw = table.shape
        ntok, W = ids.shape
        y = torch.empty(ntok, cw, dtype=f32, device=table.device)
        argpos = torch.empty(ntok, cw, dtype=torch.uint8, device=table.device)
        rc = lib.fvta_charcnn_fwd(ntok, W, VC, cw, H, _ptr(ids), _ptr(table), _ptr(bias), _ptr(y), cw, _ptr(argpos),
                                  _stream())
        check(rc, "fvta_charcnn_fwd")
        ctx.save_for_backward(ids, y, argpos)
        ctx.shape = (H, VC, cw, ntok, W)
        ctx.mark_non_differentiable(argpos)
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        f32 = torch.float32
        ids, y, argpos = ctx.saved_tensors
        H, VC, cw, ntok, W = ctx.shape
        dy = _req(dy, f32, "dy")
        d_table = torch.zeros(VC, H, cw, dtype=f32, device=dy.device)
        sink = ctx.bias_sink
        d_bias = sink if sink is not None else torch.zeros(cw, dtype=f32, device=dy.device)
        rc = lib.fvta_charcnn_bwd(ntok, W, VC, cw, H, _ptr(ids), _ptr(y), cw, _ptr(argpos), _ptr(dy), cw,
                                  _ptr(d_table), _ptr(d_bias), _stream())
        check(rc, "fvta_charcnn_bwd")
        return d_table, (None if sink is not None else d_bias), None, None


class CharTableFn(torch.autograd.Function):
    """table[v,h,:] = char_emb[v,:] @ filter[0,h]  ([VC,height,cw]) as ONE product
    char_emb @ F with F[c, h*cw + o] = filter[0,h,c,o]; the backward is two products and, with
    sinks = (gradient view of char_emb, of the filter), no autograd plumbing."""

    @staticmethod
    def forward(ctx, char_emb, filt, sinks=None):
        f32 = torch.float32
        char_emb, filt = _req(char_emb, f32, "char_emb"), _req(filt, f32, "conv filter")
        _, H, cdim, cw = filt.shape
        F = filt[0].permute(1, 0, 2).reshape(cdim, H * cw).contiguous()
        T2 = gemm(char_emb, F)
        ctx.save_for_backward(char_emb, F)
        ctx.dims = (H, cdim, cw)
        ctx.sinks = sinks
        return T2.view(char_emb.shape[0], H, cw)

    @staticmethod
    def backward(ctx, dT):
        char_emb, F = ctx.saved_tensors
        H, cdim, cw = ctx.dims
        dT2 = _req(dT, torch.float32, "d table").reshape(char_emb.shape[0], H * cw)
        g_emb, g_filt = ctx.sinks if ctx.sinks is not None else (None, None)
        if g_emb is not None:
            gemm(dT2, F, trans_b=True, out=g_emb, accumulate=True)
            d_emb = None
        else:
            d_emb = gemm(dT2, F, trans_b=True)
        d_filt = gemm(char_emb, dT2, trans_a=True).view(cdim, H, cw).permute(1, 0, 2).unsqueeze(0)
        if g_filt is not None:
            g_filt.add_(d_filt)
            d_filt = None
        return d_emb, d_filt, None


class TokenEmbedFn(torch.autograd.Function):
    """The reader's token features (model_v2.py:524-620) in one [ntok, cw + wd] buffer:
    columns [0, cw) the char-CNN of the word (table form, see CharCnnTableFn; absent when
    char_table is None), columns [cw, cw + wd) the word vector (trainable table for ids < VW, the
    fed GloVe table otherwise, :596).  Two launches forward, two backward, no concat / select
    plumbing.  sinks = (gradient view of the word table, of the conv bias) or None."""

    @staticmethod
    def forward(ctx, word_table, glove, ids, char_table, char_bias, ids_c, sinks=None):
        lib = _lib.load()
        f32 = torch.float32
        word_table, glove = _req(word_table, f32, "word table"), _req(glove, f32, "fed word table")
        ids = _req(ids, torch.int64, "token ids")
        ntok, wd = ids.shape[0], word_table.shape[1]
        VW, VF = word_table.shape[0], (glove.shape[0] if glove is not None else 0)
        cw = char_table.shape[2] if char_table is not None else 0
        ld = cw + wd
        out = torch.empty(ntok, ld, dtype=f32, device=word_table.device)
        argpos = None
        if char_table is not None:
            char_table, char_bias = _req(char_table, f32, "char table"), _req(char_bias, f32, "conv bias")
            ids_c = _req(ids_c, torch.int32, "char ids")
            VC, H, _ = char_table.shape
            W = ids_c.shape[1]
            argpos = torch.empty(ntok, cw, dtype=torch.uint8, device=out.device)
            rc = lib.fvta_charcnn_fwd(ntok, W, VC, cw, H, _ptr(ids_c), _ptr(char_table), _ptr(char_bias), _ptr(out),
                                      ld, _ptr(argpos), _stream())
            check(rc, "fvta_charcnn_fwd")
            ctx.cshape = (H, VC, cw, W)
        rc = lib.fvta_embed_rows(ntok, wd, VW, VF, _ptr(ids), _ptr(word_table), _ptr(glove),
                                 ctypes.c_void_p(out.data_ptr() + 4 * cw), ld, _stream())
        check(rc, "fvta_embed_rows")
        ctx.save_for_backward(ids, ids_c if char_table is not None else None, out, argpos)
        ctx.dims = (ntok, wd, VW, cw, ld)
        ctx.sinks = sinks
        return out

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        f32 = torch.float32
        ids, ids_c, out, argpos = ctx.saved_tensors
        ntok, wd, VW, cw, ld = ctx.dims
        dy = _req(dy, f32, "dy")
        g_word, g_bias = ctx.sinks if ctx.sinks is not None else (None, None)
        d_table = d_bias = None
        if cw:
            H, VC, _, W = ctx.cshape
            d_table = torch.zeros(VC, H, cw, dtype=f32, device=dy.device)
            d_bias = g_bias if g_bias is not None else torch.zeros(cw, dtype=f32, device=dy.device)
            rc = lib.fvta_charcnn_bwd(ntok, W, VC, cw, H, _ptr(ids_c), _ptr(out), ld, _ptr(argpos), _ptr(dy), ld,
                                      _ptr(d_table), _ptr(d_bias), _stream())
            check(rc, "fvta_charcnn_bwd")
        d_word = g_word if g_word is not None else torch.zeros(VW, wd, dtype=f32, device=dy.device)
        rc = lib.fvta_embed_rows_bwd(ntok, wd, VW, _ptr(ids), ctypes.c_void_p(dy.data_ptr() + 4 * cw), ld,
                                     _ptr(d_word), _stream())
        check(rc, "fvta_embed_rows_bwd")
        return (None if g_word is not None else d_word, None, None, d_table,
                None if (g_bias is not None or not cw) else d_bias, None, None)


# ---------------------------------------------------------------------------------------------
# small plumbing ops
# ---------------------------------------------------------------------------------------------
class EmbedFn(torch.autograd.Function):
    """Row gather whose backward is ONE index_add (atomic scatter-add of the rows) instead of
    the sort + segmented reduction of F.embedding's dense backward -- the id lists here are a few
    thousand entries, where the sort's fixed cost dominates."""

    @staticmethod
    def forward(ctx, table, ids, sink=None):
        ctx.save_for_backward(ids)
        ctx.rows = table.shape[0]
        ctx.sink = sink
        return table.index_select(0, ids)

    @staticmethod
    def backward(ctx, dy):
        (ids,) = ctx.saved_tensors
        if ctx.sink is not None:
            ctx.sink.index_add_(0, ids, dy.contiguous())
            return None, None, None
        g = torch.zeros(ctx.rows, dy.shape[1], dtype=dy.dtype, device=dy.device)
        g.index_add_(0, ids, dy.contiguous())
        return g, None, None


def embed(table, ids, sink=None):
    if table.requires_grad:
        return EmbedFn.apply(table, ids, sink)
    return table.index_select(0, ids)


class OutputLossFn(torch.autograd.Function):
    """fvta_output_loss: logits / yp / mean CE and all gradients in one launch.  Returns
    (loss, logits, yp); only `loss` is differentiable.  With sinks, dW / db go straight into the
    flat gradient buffer during the forward (skipped when autograd is off)."""

    @staticmethod
    def forward(ctx, gq, g1, gc, W, b, y, sinks, want_grad):
        lib = _lib.load()
        f32 = torch.float32
        gq, g1, gc = _req(gq, f32, "gq"), _req(g1, f32, "g1"), _req(gc, f32, "gc")
        W, b = _req(W, f32, "W"), _req(b, f32, "b")
        N, C, D = gc.shape
        dev = gq.device
        y8 = _mask_u8(y, "y") if y is not None else None
        logits = torch.empty(N, C, dtype=f32, device=dev)
        yp = torch.empty(N, C, dtype=f32, device=dev)
        loss = torch.zeros((), dtype=f32, device=dev)
        want = bool(want_grad) and y is not None   # grad mode is always off inside Function.forward
        gbuf = None
        if want:
            # one buffer so that the backward scales all three with a single launch
            seg = (N * D + 3) // 4 * 4          # keep every segment 16-byte aligned
            gbuf = torch.empty(2 * seg + N * C * D, dtype=f32, device=dev)
            d_gq, d_g1 = gbuf[:N * D].view(N, D), gbuf[seg:seg + N * D].view(N, D)
            d_gc = gbuf[2 * seg:].view(N, C, D)
            dW, db = sinks if sinks is not None else (torch.zeros_like(W), torch.zeros_like(b))
        else:
            d_gq = d_g1 = d_gc = dW = db = None
        rc = lib.fvta_output_loss(N, C, D, _ptr(gq), _ptr(g1), _ptr(gc), _ptr(W), _ptr(b), _ptr(y8), _ptr(logits),
                                  _ptr(yp), _ptr(loss), _ptr(d_gq), _ptr(d_g1), _ptr(d_gc), _ptr(dW), _ptr(db),
                                  _stream())
        check(rc, "fvta_output_loss")
        ctx.grads = (gbuf, (N, C, D), None if sinks is not None else dW, None if sinks is not None else db)
        ctx.mark_non_differentiable(logits, yp)
        return loss, logits, yp

    @staticmethod
    def backward(ctx, dloss, _dlogits, _dyp):
        gbuf, (N, C, D), dW, db = ctx.grads
        if gbuf is None:
            raise _lib.FvtaError("fvta_output_loss was run without gradients")
        # upstream is 1 for loss.backward(); honour anything else
        sc = lambda t: None if t is None else t * dloss   # noqa: E731
        g = gbuf * dloss
        seg = (N * D + 3) // 4 * 4
        return (g[:N * D].view(N, D), g[seg:seg + N * D].view(N, D), g[2 * seg:].view(N, C, D), sc(dW), sc(db),
                None, None, None)
