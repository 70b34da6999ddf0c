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

    def desc(self, N, K, T, D, JQ, has_mask):
        return FvtaAttnDesc(N=N, K=K, T=T, D=D, JQ=JQ, simi=self.simiMatrix, add_tanh=int(self.add_tanh),
                            use_time_warp=int(self.use_time_warp), warp_type=self.warp_type,
                            window=self.window, time_warp_att=int(self.time_warp_att),
                            has_mask=int(has_mask))


def _attn_params(W, b, WH_W, WH_b, WC_W, WC_b):
    return FvtaAttnParams(att_W=_ptr(W), att_b=_ptr(b), WH_W=_ptr(WH_W), WH_b=_ptr(WH_b),
                          WC_W=_ptr(WC_W), WC_b=_ptr(WC_b))


def attn_forward(cfg, hall, hall_mask, hq, hq_mask, lq, W, b, WH_W=None, WH_b=None, WC_W=None, WC_b=None,
                 want_logits=False, want_c=False, want_warp_h=False):
    """fvta_attn_fwd.  hall [N,K,T,D], hq [N,JQ,D] -> dict(g1_all, a_logits, c_col, warp_h, stats)."""
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
    desc = cfg.desc(N, K, T, D, JQ, has_mask)
    dev = hall.device
    ws_bytes = lib.fvta_attn_workspace_bytes(ctypes.byref(desc))
    if ws_bytes == 0:
        raise _lib.FvtaError("fvta_attn: unsupported shape/flags N=%d K=%d T=%d D=%d JQ=%d simi=%d"
                             % (N, K, T, D, JQ, cfg.simiMatrix))
    ws = torch.empty(ws_bytes // 4, dtype=f32, device=dev)
    stats = torch.empty(lib.fvta_attn_stats_floats(ctypes.byref(desc)), dtype=f32, device=dev)
    g1_all = torch.empty(N, D, dtype=f32, device=dev)
    a_logits = torch.empty(N, K, T, JQ, dtype=f32, device=dev) if want_logits else None
    c_col = torch.empty(N, T, dtype=f32, device=dev) if (want_c and cfg.use_time_warp) else None
    warp_h = torch.empty(N, K, T, D, dtype=f32, device=dev) if (want_warp_h and cfg.use_time_warp) else None
    params = _attn_params(W, b, WH_W, WH_b, WC_W, WC_b)
    rc = lib.fvta_attn_fwd(ctypes.byref(desc), _ptr(hall), _ptr(hm), _ptr(hq), _ptr(qm), _ptr(lq),
                           ctypes.byref(params), _ptr(g1_all), _ptr(a_logits), _ptr(c_col), _ptr(warp_h),
                           _ptr(stats), _ptr(ws), ws_bytes, _stream())
    check(rc, "fvta_attn_fwd")
    return dict(g1_all=g1_all, a_logits=a_logits, c_col=c_col, warp_h=warp_h, stats=stats, ws=ws,
                desc=desc, saved=(hall, hm, hq, qm, lq))


def attn_backward(fwd, d_g1_all, W, b, WH_W=None, WH_b=None, WC_W=None, WC_b=None):
    """fvta_attn_bwd given the dict attn_forward returned.  Returns dict of gradients."""
    lib = _lib.load()
    f32 = torch.float32
    hall, hm, hq, qm, lq = fwd["saved"]
    desc = fwd["desc"]
    d_g1_all = _req(d_g1_all, f32, "d_g1_all")
    dev = hall.device
    out = dict(d_hall=torch.empty_like(hall), d_hq=torch.empty_like(hq),
               d_lq=torch.empty_like(lq) if (lq is not None and desc.use_time_warp) else None)
    gW = torch.empty_like(W) if W is not None else None
    gb = torch.empty_like(b) if b is not None else None
    gs = [torch.empty_like(x) if (x is not None and desc.use_time_warp) else None
          for x in (WH_W, WH_b, WC_W, WC_b)]
    grads = FvtaAttnGrads(att_W=_ptr(gW), att_b=_ptr(gb), WH_W=_ptr(gs[0]), WH_b=_ptr(gs[1]),
                          WC_W=_ptr(gs[2]), WC_b=_ptr(gs[3]))
    params = _attn_params(W, b, WH_W, WH_b, WC_W, WC_b)
    ws = fwd["ws"]
    rc = lib.fvta_attn_bwd(ctypes.byref(desc), _ptr(hall), _ptr(hm), _ptr(hq), _ptr(qm), _ptr(lq),
                           ctypes.byref(params), _ptr(fwd["stats"]), _ptr(d_g1_all), _ptr(out["d_hall"]),
                           _ptr(out["d_hq"]), _ptr(out["d_lq"]), ctypes.byref(grads), _ptr(ws),
                           ws.numel() * 4, _stream())
    check(rc, "fvta_attn_bwd")
    out.update(d_att_W=gW, d_att_b=gb, d_WH_W=gs[0], d_WH_b=gs[1], d_WC_W=gs[2], d_WC_b=gs[3])
    return out


def adadelta_step(var, grad, accum, accum_update, update_mask, lr, rho=0.95, eps=1e-8, grad_scale=1.0):
    """fvta_adadelta over flat buffers (in place)."""
    lib = _lib.load()
    f32 = torch.float32
    for t, nme in ((var, "var"), (grad, "grad"), (accum, "accum"), (accum_update, "accum_update")):
        if _req(t, f32, nme) is not t:
            raise _lib.FvtaError("%s must be contiguous" % nme)
    rc = lib.fvta_adadelta(var.numel(), _ptr(var), _ptr(grad), _ptr(accum), _ptr(accum_update),
                           _ptr(update_mask), lr, rho, eps, grad_scale, _stream())
    check(rc, "fvta_adadelta")
