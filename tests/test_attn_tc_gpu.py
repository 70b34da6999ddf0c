"""GPU parity of the tcgen05 / TMEM / TMA forward kernel (csrc/attn_fwd_tc.cu) in the configurations
bench.py and the roofline measure: FVTA flags (simi 2, tanh, band window), D = 100, K = 6, many tiles
per CTA.  Every case

  * asserts, through the C ABI's own path query, that the tensor-core kernel is what ran;
  * compares with the fp64 collapsed oracle (oracle/collapsed.py, itself pinned to the literal
    restatement of model_v2.py:210-341, 953-1009 in tests/test_oracle.py) under a FIXED bound;
  * runs the same inputs again on the CUDA-core kernel (FvtaAttnDesc.path = FVTA_PATH_FFMA) and
    compares the two kernels with each other.

Bounds (fp32, BASELINE.json north_star "1e-4 relative"): g1_all 1e-4 of max|g1_all|; logits 1e-4
ABSOLUTE (they are tanh / cosine values in [-1, 1], so this is also 1e-4 of the largest one, and it
is elementwise: small entries are held to the same absolute error); masked logits bit-exact -1e30.
The achieved errors are printed (pytest -s / the captured log) -- no bound adapts to the data.
"""
import pytest
import torch

from oracle import collapsed
from fvta_memexqa_b200 import ops, _lib
from tests.helpers import attn_case

pytestmark = pytest.mark.gpu
BOUND = 1e-4


def _run(cfg, params, inp, dev, path=0, **want):
    acfg = ops.AttnConfig(cfg.simiMatrix, cfg.add_tanh, cfg.use_time_warp, cfg.warp_type, 3.0,
                          cfg.use_time_warp_att)
    N, K, M, JMAX, D = inp["hall"].shape
    p = {k: v.to(dev) for k, v in params.items()}
    return ops.attn_forward(
        acfg, inp["hall"].reshape(N, K, M * JMAX, D).to(dev), inp["hall_mask"].reshape(N, K, -1).to(dev),
        inp["hq"].to(dev), inp["q_mask"].to(dev), inp["lq"].to(dev),
        p.get("attention/all/att_logits/W"), p.get("attention/all/att_logits/b"),
        p.get("time_warp/WH/W"), p.get("time_warp/WH/b"), p.get("time_warp/WC/W"), p.get("time_warp/WC/b"),
        path=path, **want), p


def _oracle64(cfg, params, inp):
    p64 = {k: v.double() for k, v in params.items()}
    return collapsed.fvta_attention(inp["hall"].double(), inp["hall_mask"], inp["hq"].double(),
                                    inp["q_mask"], inp["lq"].double(), p64, cfg)


def _errors(out, ref):
    g = out["g1_all"].double().cpu()
    e_g = ((g - ref["g1_all"]).abs().max() / (ref["g1_all"].abs().max() + 1e-30)).item()
    res = {"g1_all_rel": e_g}
    if out["a_logits"] is not None:
        a = out["a_logits"].double().cpu().reshape(ref["att_logits"].shape)
        valid = ref["att_logits"] > -1e29
        res["logit_abs"] = ((a - ref["att_logits"]).abs() * valid).max().item()
        res["masked_exact"] = bool(torch.all(out["a_logits"].cpu().reshape(valid.shape)[~valid]
                                             == torch.tensor(-1e30, dtype=torch.float32)))
    if out["c_col"] is not None and ref.get("c") is not None:
        res["c_abs"] = (out["c_col"].double().cpu() - ref["c"]).abs().max().item()
    return res


def _check(name, errs, logit_bound=BOUND):
    print("[tc parity] %-52s %s" % (name, "  ".join("%s=%.2e" % (k, v) if isinstance(v, float) else "%s=%s" % (k, v)
                                                     for k, v in errs.items())))
    assert errs["g1_all_rel"] < BOUND, (name, errs)
    if "logit_abs" in errs:
        assert errs["logit_abs"] < logit_bound, (name, errs)
        assert errs["masked_exact"], (name, errs)
    if "c_abs" in errs:
        assert errs["c_abs"] < BOUND, (name, errs)


# N, K, M, JMAX, D, JQ, simi, tanh, warp, wtype, density, q_density
TC_CASES = {
    "bench shape N=32 T=256 (C2 batch, 3 tiles per CTA)": (32, 6, 4, 64, 100, 25, 2, True, True, 5, 0.25, 0.6),
    "bench shape, dense masks": (32, 6, 4, 64, 100, 25, 2, True, True, 5, 1.0, 1.0),
    "N=4 T=4096 (7 tiles per CTA)": (4, 6, 8, 512, 100, 32, 2, True, True, 5, 0.5, 0.7),
    "C3-max slice N=1 T=32768 (>= 10 tiles per CTA, ring wraps)": (1, 6, 64, 512, 100, 32, 2, True, True, 5, 1.0, 1.0),
    "N=2 T=256 (small N: more CTAs than FFMA tiles)": (2, 6, 4, 64, 100, 25, 2, True, True, 5, 0.3, 0.6),
    "K=5 (--no_photo)": (3, 5, 4, 64, 100, 25, 2, True, True, 5, 0.4, 0.6),
    "simi 1 + tanh": (2, 6, 4, 40, 100, 25, 1, True, True, 5, 0.5, 0.6),
    "simi 3 + tanh": (2, 6, 4, 40, 100, 25, 3, True, True, 5, 0.5, 0.6),
    "cosine": (2, 6, 4, 40, 100, 32, 4, False, True, 5, 0.5, 0.6),
    "no time warp": (2, 6, 4, 64, 100, 25, 2, True, False, 1, 0.5, 0.6),
    "D=8": (2, 6, 2, 20, 8, 4, 2, True, True, 5, 0.5, 0.6),
    "D=36": (2, 6, 4, 40, 36, 13, 2, True, True, 5, 0.5, 0.6),
    "D=64": (2, 6, 4, 64, 64, 32, 2, True, True, 5, 0.5, 0.6),
    "D=128, K=8": (2, 8, 4, 64, 128, 32, 2, True, True, 5, 0.5, 0.6),
    "window type 2 (diagonal)": (2, 6, 4, 40, 100, 25, 2, True, True, 2, 0.5, 0.6),
    "T=12 (one partial tile)": (3, 6, 1, 12, 100, 25, 2, True, True, 5, 0.5, 0.6),
}


@pytest.mark.parametrize("name", list(TC_CASES))
def test_tc_kernel_vs_oracle_and_ffma(cuda_device, name):
    N, K, M, JMAX, D, JQ, simi, tanh, warp, wtype, dens, qd = TC_CASES[name]
    cfg, params, inp = attn_case(N, K, M, JMAX, D, JQ, simi, tanh, warp, wtype, False, density=dens,
                                 q_density=qd, seed=17 + N + D)
    ref = _oracle64(cfg, params, inp)
    for want_logits, want_c in ((True, True), (False, False)):
        out, _ = _run(cfg, params, inp, cuda_device, want_logits=want_logits, want_c=want_c)
        assert out["path"] == _lib.PATH_TC, "the tcgen05 kernel did not run: %s" % _lib.PATH_NAMES.get(out["path"])
        _check("%s%s" % (name, "" if want_logits else " (no optional outputs)"), _errors(out, ref))
    # A/B: the CUDA-core kernel on the same inputs
    ffma, _ = _run(cfg, params, inp, cuda_device, path=_lib.PATH_FFMA, want_logits=True, want_c=True)
    assert ffma["path"] == _lib.PATH_FFMA
    tc, _ = _run(cfg, params, inp, cuda_device, want_logits=True, want_c=True)
    d_g = ((tc["g1_all"] - ffma["g1_all"]).abs().max() / ffma["g1_all"].abs().max()).item()
    valid = (ffma["a_logits"] > -1e29)
    d_a = ((tc["a_logits"] - ffma["a_logits"]).abs() * valid).max().item()
    print("[tc vs ffma] %-51s g1_all_rel=%.2e logit_abs=%.2e" % (name, d_g, d_a))
    assert d_g < BOUND and d_a < BOUND
    assert torch.equal(tc["a_logits"][~valid], ffma["a_logits"][~valid])
    # the statistics the backward consumes: softmax normaliser convention differs (fixed reference
    # maximum 1 vs running maximum) but alpha = exp(m - M) / L and beta must agree
    K_, Dd = K, D
    st_tc = tc["stats"][:N * K_ * 8].reshape(N, K_, 8).cpu().double()
    st_ff = ffma["stats"][:N * K_ * 8].reshape(N, K_, 8).cpu().double()
    assert (st_tc[..., 4] - st_ff[..., 4]).abs().max() < 1e-4           # beta
    assert torch.equal(st_tc[..., 2] > -1e29, st_ff[..., 2] > -1e29)    # which modalities are fully masked
    r_tc = tc["stats"][N * K_ * 8:].reshape(N, K_, Dd).cpu().double()
    r_ff = ffma["stats"][N * K_ * 8:].reshape(N, K_, Dd).cpu().double()
    assert ((r_tc - r_ff).abs().max() / (r_ff.abs().max() + 1e-30)) < 1e-4


def test_tc_fully_masked_sample_and_modality(cuda_device):
    """TF semantics: a fully masked modality gets beta = 0 unless every modality
    is masked; a fully masked SAMPLE gives the uniform softmax over all T rows of every modality and
    beta = 1/K -- the 2^-100 masked weight of the tensor-core kernel must reproduce both."""
    cfg, params, inp = attn_case(4, 6, 4, 64, 100, 25, seed=5, density=0.5)
    inp["hall_mask"][1] = False                     # whole sample masked (zero-padded batch row) ...
    inp["hall"][1] = 0
    inp["hall_mask"][2, 3] = False                  # ... one modality masked
    inp["hall"][2, 3] = 0
    inp["q_mask"][3, 5:] = False
    inp["hq"][3, 5:] = 0
    out, _ = _run(cfg, params, inp, cuda_device, want_logits=True, want_c=True)
    assert out["path"] == _lib.PATH_TC
    ref = _oracle64(cfg, params, inp)
    _check("fully masked sample / modality", _errors(out, ref))
    assert torch.isfinite(out["g1_all"]).all()
    assert out["g1_all"][1].abs().max().item() == 0.0
    beta = out["stats"][:4 * 6 * 8].reshape(4, 6, 8)[..., 4].cpu()
    assert torch.allclose(beta[1], torch.full((6,), 1.0 / 6.0), atol=1e-6)
    assert beta[2, 3].item() == 0.0


def test_tc_saturated_tanh_ties(cuda_device):
    """Large pre-activations saturate tanh to exactly 1.0f: many tied maxima within and across tiles.
    g1_all and the tie bookkeeping (count, first position) must match the oracle's amax semantics.
    Logit bound 1e-3 here (stated, not derived): with the inputs scaled by 3 the pre-activation
    S = g*P + g^2*(w2.h^2) + ... has terms of O(1e4), where ONE fp32 ulp is 1e-3 -- any fp32
    evaluation of model_v2.py:244-246 (TF's included) carries that on the few unsaturated entries."""
    cfg, params, inp = attn_case(2, 6, 8, 64, 100, 25, seed=9, density=0.8)
    inp["hall"] = inp["hall"] * 3.0
    inp["hq"] = inp["hq"] * 3.0
    out, _ = _run(cfg, params, inp, cuda_device, want_logits=True)
    assert out["path"] == _lib.PATH_TC
    ref = _oracle64(cfg, params, inp)
    _check("saturated tanh (ties)", _errors(out, ref), logit_bound=1e-3)
    a = out["a_logits"].cpu()
    st = out["stats"][:2 * 6 * 8].reshape(2, 6, 8).cpu()
    for n in range(2):
        for k in range(6):
            s = a[n, k].max()
            assert st[n, k, 2].item() == s.item()
            assert st[n, k, 3].item() == float((a[n, k] == s).sum().item())      # tie count
            pos = (a[n, k] == s).nonzero()[0]
            assert (st[n, k, 5].item(), st[n, k, 6].item()) == (float(pos[0]), float(pos[1]))


def test_tc_backward_after_tc_forward(cuda_device):
    """The training step runs the tensor-core forward and the CUDA-core backward: the gradients from
    that pair must match the fp64 oracle at the bench shape."""
    N, K, M, JMAX, D, JQ = 8, 6, 4, 64, 100, 25
    cfg, params, inp = attn_case(N, K, M, JMAX, D, JQ, seed=21, density=0.3)
    out, p = _run(cfg, params, inp, cuda_device)
    assert out["path"] == _lib.PATH_TC
    g = torch.Generator().manual_seed(5)
    dg1 = torch.randn(N, D, generator=g)
    gr = ops.attn_backward(out, dg1.to(cuda_device), p.get("attention/all/att_logits/W"),
                           p.get("attention/all/att_logits/b"), p.get("time_warp/WH/W"),
                           p.get("time_warp/WH/b"), p.get("time_warp/WC/W"), p.get("time_warp/WC/b"))
    pp = {k: v.double().clone().requires_grad_(True) for k, v in params.items()}
    hall = inp["hall"].double().clone().requires_grad_(True)
    hq = inp["hq"].double().clone().requires_grad_(True)
    lq = inp["lq"].double().clone().requires_grad_(True)
    o = collapsed.fvta_attention(hall, inp["hall_mask"], hq, inp["q_mask"], lq, pp, cfg)
    (o["g1_all"] * dg1.double()).sum().backward()
    refs = {"d_hall": hall.grad.reshape(N, K, -1, D), "d_hq": hq.grad, "d_lq": lq.grad,
            "d_att_W": pp["attention/all/att_logits/W"].grad, "d_att_b": pp["attention/all/att_logits/b"].grad,
            "d_WH_W": pp["time_warp/WH/W"].grad, "d_WH_b": pp["time_warp/WH/b"].grad,
            "d_WC_W": pp["time_warp/WC/W"].grad, "d_WC_b": pp["time_warp/WC/b"].grad}
    # bounds per gradient: 1e-4 of its largest entry, except sums over all N*K*T rows with heavy
    # cancellation (d_lq, d_WC_*, d_WH_b: K * W_WC * sum_t dz), which carry the fp32 rounding of their
    # partial sums: 5e-4, stated here, not derived from the data
    bounds = {"d_lq": 5e-4, "d_WC_W": 5e-4, "d_WC_b": 5e-4, "d_WH_b": 5e-4}
    for k, r in refs.items():
        e = ((gr[k].double().cpu() - r).abs().max() / (r.abs().max() + 1e-30)).item()
        print("[tc fwd + bwd] %-8s rel=%.2e" % (k, e))
        assert e < bounds.get(k, BOUND), (k, e)


BWD_TC_CASES = {
    "bench shape N=8": (8, 6, 4, 64, 100, 25, 2, True, True, 5, 0.3, 0.6),
    "dense masks, N=3": (3, 6, 4, 64, 100, 25, 2, True, True, 5, 1.0, 1.0),
    "long T, ring wraps": (2, 6, 8, 128, 100, 32, 2, True, True, 5, 0.6, 0.8),
    "K=5": (3, 5, 4, 64, 100, 25, 2, True, True, 5, 0.4, 0.6),
    "simi 1 + tanh": (2, 6, 4, 40, 100, 25, 1, True, True, 5, 0.5, 0.6),
    "simi 3 + tanh": (2, 6, 4, 40, 100, 25, 3, True, True, 5, 0.5, 0.6),
    "cosine": (2, 6, 4, 40, 100, 32, 4, False, True, 5, 0.5, 0.6),
    "no time warp": (2, 6, 4, 64, 100, 25, 2, True, False, 1, 0.5, 0.6),
    "D=36": (2, 6, 4, 40, 36, 13, 2, True, True, 5, 0.5, 0.6),
    "D=128, K=8": (2, 8, 4, 64, 128, 32, 2, True, True, 5, 0.5, 0.6),
    "T=12": (3, 6, 1, 12, 100, 25, 2, True, True, 5, 0.5, 0.6),
}


@pytest.mark.parametrize("name", list(BWD_TC_CASES))
def test_tc_backward_vs_oracle_and_ffma(cuda_device, name):
    """The tensor-core backward (csrc/attn_bwd_tc.cu: tcgen05 recompute + sparse d-logit routing) against the
    fp64 oracle's autograd and against the CUDA-core backward on the same forward.  Fixed bounds: 1e-4 of
    the largest entry of each gradient, 5e-4 for the four gradients that are sums over every row with heavy
    cancellation (d_lq, d_WC_W, d_WC_b, d_WH_b)."""
    N, K, M, JMAX, D, JQ, simi, tanh, warp, wtype, dens, qd = BWD_TC_CASES[name]
    cfg, params, inp = attn_case(N, K, M, JMAX, D, JQ, simi, tanh, warp, wtype, False, density=dens,
                                 q_density=qd, seed=5 + N + D)
    out, p = _run(cfg, params, inp, cuda_device)
    assert out["path"] == _lib.PATH_TC
    g = torch.Generator().manual_seed(5)
    dg1 = torch.randn(N, D, generator=g)
    names = ("attention/all/att_logits/W", "attention/all/att_logits/b", "time_warp/WH/W", "time_warp/WH/b",
             "time_warp/WC/W", "time_warp/WC/b")
    args = [p.get(k) for k in names]
    gr = ops.attn_backward(out, dg1.to(cuda_device), *args)
    gr_ff = ops.attn_backward(out, dg1.to(cuda_device), *args, path=_lib.PATH_FFMA)
    pp = {k: v.double().clone().requires_grad_(True) for k, v in params.items()}
    hall = inp["hall"].double().clone().requires_grad_(True)
    hq = inp["hq"].double().clone().requires_grad_(True)
    lq = inp["lq"].double().clone().requires_grad_(True)
    o = collapsed.fvta_attention(hall, inp["hall_mask"], hq, inp["q_mask"], lq, pp, cfg)
    (o["g1_all"] * dg1.double()).sum().backward()
    refs = {"d_hall": hall.grad.reshape(N, K, -1, D), "d_hq": hq.grad, "d_lq": lq.grad}
    for key, nme in zip(("d_att_W", "d_att_b", "d_WH_W", "d_WH_b", "d_WC_W", "d_WC_b"), names):
        refs[key] = pp[nme].grad if nme in pp else None
    bounds = {"d_lq": 5e-4, "d_WC_W": 5e-4, "d_WC_b": 5e-4, "d_WH_b": 5e-4}
    for k, r in refs.items():
        if r is None or gr.get(k) is None:
            continue
        if r.abs().max() == 0:
            assert gr[k].abs().max().item() == 0.0, k
            continue
        e = ((gr[k].double().cpu() - r).abs().max() / r.abs().max()).item()
        e_ff = ((gr[k].double().cpu() - gr_ff[k].double().cpu()).abs().max() / r.abs().max()).item()
        print("[tc bwd] %-22s %-8s vs oracle %.2e  vs ffma %.2e" % (name, k, e, e_ff))
        assert e < bounds.get(k, BOUND), (k, e)
        assert e_ff < 2 * bounds.get(k, BOUND), (k, e_ff)


@pytest.mark.parametrize("warp", [False, True], ids=["no-warp", "time-warp"])
def test_tc_backward_tied_columns(cuda_device, warp):
    """Six identical question words: their logit columns tie EXACTLY in every context row, so a row whose
    maximum falls on them has six equal, unsaturated maxima (tf.reduce_max splits the gradient evenly).
    More than three entries per row is the tensor-core backward's serial path (the row's thread keeps its
    d P vector in tensor memory and applies the row's gradient itself) -- with small inputs so that the
    tanh stays away from +-1 and the entries really carry gradient."""
    N, K, M, JMAX, D, JQ = 2, 6, 4, 32, 100, 25
    cfg, params, inp = attn_case(N, K, M, JMAX, D, JQ, 2, True, warp, 5, False, density=0.7, q_density=1.0, seed=77)
    inp["hall"] = inp["hall"] * 0.15
    inp["hq"] = inp["hq"] * 0.15
    inp["hq"][:, 1:6] = inp["hq"][:, 0:1]
    inp["q_mask"][:, :6] = True
    out, p = _run(cfg, params, inp, cuda_device, want_logits=True)
    assert out["path"] == _lib.PATH_TC
    a = out["a_logits"].cpu().reshape(N, K, -1, out["a_logits"].shape[-1])
    valid = inp["hall_mask"].reshape(N, K, -1)
    rowmax = a[..., :JQ].max(-1).values
    tied6 = ((a[..., :6] == rowmax[..., None]).all(-1) & valid & (rowmax.abs() < 1.0))
    assert tied6.sum().item() >= 10, "the case must contain rows with six tied, unsaturated maxima"
    g = torch.Generator().manual_seed(11)
    dg1 = torch.randn(N, D, generator=g)
    names = ("attention/all/att_logits/W", "attention/all/att_logits/b", "time_warp/WH/W", "time_warp/WH/b",
             "time_warp/WC/W", "time_warp/WC/b")
    args = [p.get(k) for k in names]
    gr = ops.attn_backward(out, dg1.to(cuda_device), *args)
    gr_ff = ops.attn_backward(out, dg1.to(cuda_device), *args, path=_lib.PATH_FFMA)
    pp = {k: v.double().clone().requires_grad_(True) for k, v in params.items()}
    hall = inp["hall"].double().clone().requires_grad_(True)
    hq = inp["hq"].double().clone().requires_grad_(True)
    lq = inp["lq"].double().clone().requires_grad_(True)
    o = collapsed.fvta_attention(hall, inp["hall_mask"], hq, inp["q_mask"], lq, pp, cfg)
    (o["g1_all"] * dg1.double()).sum().backward()
    refs = {"d_hall": hall.grad.reshape(N, K, -1, D), "d_hq": hq.grad, "d_lq": lq.grad if warp else None}
    for key, nme in zip(("d_att_W", "d_att_b", "d_WH_W", "d_WH_b", "d_WC_W", "d_WC_b"), names):
        refs[key] = pp[nme].grad if nme in pp else None
    # the five time-warp gradients are sums of d z over every row with heavy cancellation (see
    # tests/test_attn_gpu.py::TC_CANCELLING_BOUND); with inputs this small the sums are smaller still
    bounds = {"d_lq": 1.5e-3, "d_WC_W": 1.5e-3, "d_WC_b": 1.5e-3, "d_WH_b": 1.5e-3, "d_WH_W": 1.5e-3}
    for k, r in refs.items():
        if r is None or gr.get(k) is None or r.abs().max() == 0:
            continue
        e = ((gr[k].double().cpu() - r).abs().max() / r.abs().max()).item()
        e_ff = ((gr[k].double().cpu() - gr_ff[k].double().cpu()).abs().max() / r.abs().max()).item()
        print("[tc bwd tied columns] %-8s vs oracle %.2e  vs ffma %.2e" % (k, e, e_ff))
        assert e < bounds.get(k, BOUND), (k, e)
        assert e_ff < 2 * bounds.get(k, BOUND), (k, e_ff)


def test_folded_workspace_reuse(cuda_device):
    """FvtaAttnDesc.ws_from_fwd on the forward: a second call over a NEW context with the same question
    reuses the folded operands and launches exactly one kernel."""
    cfg, params, inp = attn_case(4, 6, 4, 64, 100, 25, seed=31, density=0.5)
    out, p = _run(cfg, params, inp, cuda_device, want_logits=True)
    lib = _lib.load()
    inp2 = dict(inp, hall=inp["hall"].flip(0).contiguous(), hall_mask=inp["hall_mask"].flip(0).contiguous())
    acfg = ops.AttnConfig(cfg.simiMatrix, cfg.add_tanh, cfg.use_time_warp, cfg.warp_type, 3.0, False)
    N, K, M, JMAX, D = inp["hall"].shape
    dev = cuda_device
    n0 = lib.fvta_launch_count()
    out2 = ops.attn_forward(acfg, inp2["hall"].reshape(N, K, -1, D).to(dev), inp2["hall_mask"].reshape(N, K, -1).to(dev),
                            inp["hq"].to(dev), inp["q_mask"].to(dev), inp["lq"].to(dev),
                            p["attention/all/att_logits/W"], p["attention/all/att_logits/b"], p["time_warp/WH/W"],
                            p["time_warp/WH/b"], p["time_warp/WC/W"], p["time_warp/WC/b"], want_logits=True,
                            reuse=out, folded=True)
    assert lib.fvta_launch_count() - n0 == 1
    ref = _oracle64(cfg, params, inp2)
    _check("folded workspace reuse", _errors(out2, ref))
