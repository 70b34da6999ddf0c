#!/usr/bin/env python
"""bench.py -- FVTA QA samples/sec fwd+bwd (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one full training step of the FVTA model (README.MD:222-226 flags) at the C2 shape
(N=32 QAs per GPU, M=4 albums x JI=8 photos, JQ=25, hidden 50, fp32):
embeddings -> BiLSTM reader -> context tensor -> time-warp gate + focal attention -> question
attention -> output -> loss -> backward -> (NCCL all-reduce of the flat gradient) -> Adadelta.

One JSON line is printed by rank 0:
  value       whole-job samples/s with the batch already resident in HBM
  e2e         the same through Trainer.step (the reference-facing call) from PINNED HOST tensors,
              host->device copies and the loss read-back inside the timed region
  roofline    the focal-attention forward (fvta_attn_fwd) at the C3-max microbenchmark shape
              (64 albums x 64 photos, JQ 32; 315 MB of context > L2), algorithmic bytes / time
  cpu_baseline  the literal CPU oracle (a port: TensorFlow cannot run here) on the host cores
`--impl reference` times that CPU oracle alone, on the same config/metric.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "FVTA QA samples/sec fwd+bwd"
UNIT = "samples/s"
PER_GPU_BATCH = 32


def workload_config(n_gpus, strong=False):
    name = ("C4 (strong scaling): the C2 training step at a FIXED global batch of %d split over the GPUs"
            % (PER_GPU_BATCH * n_gpus)) if strong else "C2: FVTA training step (fwd+bwd+Adadelta)"
    return {"workload": name + ", README.MD:222-226 flags, "
                        "batch %d per GPU, M=4 JI=8 JXP=8 JD=10 JXA=8 JT=4 JG=4 JQ=25 JA=5 W=16, hidden 50, "
                        "idim 2537, K=6 T=256 D=100" % PER_GPU_BATCH,
            "global_batch": PER_GPU_BATCH * n_gpus, "per_gpu_batch": PER_GPU_BATCH,
            "parallelism": "dp%d" % n_gpus,
            "l2": "flushed between steps (256 MiB memset between the per-step CUDA-event pairs; ms_per_step_incl_l2_flush brackets the whole loop)",
            "mask_density": "MemexQA-shaped (SURVEY 8d length distributions)"}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.samples = []      # (host time, csv line)
        self.proc = None
        self.window = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.perf_counter(), line.strip()))

    def mark(self, t0, t1):
        """Host-clock window of the timed region; samples inside it are the ones reported."""
        self.window = (t0, t1)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        lines = [l for (t, l) in self.samples if self.window and self.window[0] <= t <= self.window[1]]
        scope = "timed region"
        if not lines:   # region shorter than the sampling period: use the run's samples under load
            lines = [l for (_, l) in self.samples[-25:]]
            scope = "whole run (timed region shorter than one nvidia-smi period)"
        for line in lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for nme, val in zip(names, parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx,
                "reasons": sorted(reasons), "samples": len(sm), "scope": scope}


# ------------------------------------------------------------------------------------------
# CPU oracle arm (cpu_baseline and --impl reference)
# ------------------------------------------------------------------------------------------
def cpu_reference_step(n_samples, seed=0):
    """One bounded CPU sample of the workload: the LITERAL oracle (op-for-op TF graph incl. the
    [K,T,T,4d] time-warp tensors) run as the README runs FVTA -- micro-batches of 1
    (README.MD:146-148) -- fwd+bwd per QA, gradients accumulated, one Adadelta update.
    Returns seconds."""
    import torch
    from oracle import full as oracle_full
    from oracle.tf_ops import adadelta_update
    from fvta_memexqa_b200.config import fvta_config
    from fvta_memexqa_b200.params import ParamStore, NEVER_UPDATED
    from fvta_memexqa_b200.synthetic import make_feed
    cfg = fvta_config(batch_size=1)
    ps = ParamStore(cfg, seed=1234)
    params = {k: v.clone() for k, v in ps.as_dict().items()}
    feeds = [make_feed(cfg, seed=1234 + seed * 100 + i, glove_vocab=300) for i in range(n_samples)]
    t0 = time.perf_counter()
    acc = {}
    for feed in feeds:
        leaves = {k: v.clone().requires_grad_(True) for k, v in params.items()}
        out = oracle_full.forward(feed, leaves, cfg, is_train=True, block="literal")
        names = [k for k in leaves if k not in NEVER_UPDATED]
        grads = torch.autograd.grad(out["loss"], [leaves[k] for k in names], allow_unused=True)
        for k, g in zip(names, grads):
            if g is not None:
                acc[k] = acc.get(k, 0) + g / n_samples
    for k, g in acc.items():
        z = torch.zeros_like(params[k])
        params[k], _, _ = adadelta_update(params[k], g, z, z, cfg.init_lr)
    return time.perf_counter() - t0


def run_reference(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    # bounded sample per step (~0.33 s of 16-core CPU work per QA): the whole run stays within a few
    # minutes whatever K the caller asks for
    n_samples = 4 if args.steps <= 25 else (2 if args.steps <= 60 else 1)
    for i in range(min(args.warmup, 3)):
        cpu_reference_step(1, seed=100 + i)
    times = [cpu_reference_step(n_samples, seed=i) for i in range(args.steps)]
    total = sum(times)
    value = n_samples * args.steps / total
    sample = ("%d QAs per step in micro-batches of 1 (README.MD:146-148), literal torch-CPU restatement of "
              "model_v2.py incl. the [K,T,T,4d] time-warp tensors, fwd+bwd+Adadelta; TensorFlow 1.x "
              "cannot run in this image" % n_samples)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------
# roofline leg: focal attention forward at the C3-max shape
# ------------------------------------------------------------------------------------------
def attention_roofline(dev, iters=10, with_logits=False):
    import torch
    from fvta_memexqa_b200 import ops
    from fvta_memexqa_b200.config import fvta_config
    from fvta_memexqa_b200.params import ParamStore
    N, K, M, JI, JXP, D, JQ = 4, 6, 64, 64, 8, 100, 32
    T = M * JI * JXP
    cfg = fvta_config(batch_size=N)
    ps = ParamStore(cfg, seed=7, device=dev)
    p = ps.as_dict()
    g = torch.Generator(device=dev).manual_seed(1)
    hall = torch.randn(N, K, T, D, device=dev, generator=g)
    hmask = torch.ones(N, K, T, dtype=torch.bool, device=dev)       # mask density 1.0
    hq = torch.randn(N, JQ, D, device=dev, generator=g)
    qmask = torch.ones(N, JQ, dtype=torch.bool, device=dev)
    lq = torch.randn(N, D, device=dev, generator=g)
    acfg = ops.AttnConfig(2, True, True, 5, 3.0, False)

    from fvta_memexqa_b200 import _lib
    lib = _lib.load()
    bufs = {"r": None}

    def call(folded=False):
        bufs["r"] = ops.attn_forward(acfg, hall, hmask, hq, qmask, lq, p["attention/all/att_logits/W"],
                                     p["attention/all/att_logits/b"], p["time_warp/WH/W"], p["time_warp/WH/b"],
                                     p["time_warp/WC/W"], p["time_warp/WC/b"], want_logits=with_logits,
                                     reuse=bufs["r"], folded=folded)

    graphs = {}

    def graph_for(folded):
        """The call as a CUDA graph (what the training step replays): the host cost of the ctypes
        binding (~50 us per call) would otherwise leave the GPU idle between 100 us launches."""
        if folded not in graphs:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                call(folded)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g_ = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_):
                call(folded)
            graphs[folded] = g_
        return graphs[folded]

    def timed(n, folded, burst=20):
        """n event pairs, each around `burst` back-to-back replays (no host sync inside a pair: the
        GPU stays busy, so the SM clock is the sustained one, not the idle-ramp one a lone 0.1 ms launch
        sees).  The working set (315 MB) exceeds L2, so consecutive launches do not feed each other."""
        g_ = graph_for(folded)
        ts = []
        for _ in range(n):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(burst):
                g_.replay()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e-3 / burst)
        return statistics.mean(ts), min(ts)

    def warm(seconds=0.4):
        t_end = time.time() + seconds
        while time.time() < t_end:
            for _ in range(20):
                call(True)
            torch.cuda.synchronize()
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    if bufs["r"]["path"] != _lib.PATH_TC:
        raise SystemExit("bench.py: the C3-max attention did not run on the tcgen05 kernel (path %r)" % bufs["r"]["path"])
    # parity of exactly this launch, outside the timed region: a slice of the first sample against the
    # fp64 CPU oracle (the whole C3-max context would take the oracle minutes)
    parity = attention_roofline_parity(acfg, p, hall, hmask, hq, qmask, lq)
    call()
    warm()
    sampler = ClockSampler(dev.index or 0)
    sampler.start()
    h0 = time.perf_counter()
    t_op, _ = timed(iters, False)        # the whole operator: operand folding (one small kernel) + the tile kernel
    # the dominant kernel alone: the workspace already holds the folded question operands
    # (FvtaAttnDesc.ws_from_fwd), so the call is exactly ONE launch: attn_fwd_tc incl. its fused combine
    n0 = lib.fvta_launch_count()
    call(True)
    assert lib.fvta_launch_count() - n0 == 1
    torch.cuda.synchronize()
    t, t_min = timed(iters, True)
    sampler.mark(h0, time.perf_counter())
    clocks = sampler.stop()
    # algorithmic bytes of the attention forward
    bytes_alg = 4 * N * K * T * D + N * K * T + 4 * N * JQ * D + N * JQ + 4 * N * D + 4 * N * D
    if with_logits:
        bytes_alg += 4 * N * K * T * JQ
    peak, how = measured_peaks()
    achieved = bytes_alg / t / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "attn_fwd_traffic.json")
    if os.path.exists(tp):
        with open(tp) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
    return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic, "kernel": "attn_fwd_tc (tile kernel of fvta_attn_fwd incl. the fused combine)",
            "launch_ms_min": t_min * 1e3, "parity": parity, "clocks": clocks,
            "whole_op": {"launch_ms": t_op * 1e3, "achieved": bytes_alg / t_op / 1e9,
                         "frac": bytes_alg / t_op / 1e9 / peak,
                         "what": "fvta_attn_fwd = operand folding (attn_prep_all) + attn_fwd_tc"},
            "workload": "C3-max: N=4 K=6 T=64*64*8=32768 D=100 JQ=32, mask density 1.0, working set "
                        "%.0f MB > L2" % (bytes_alg / 1e6),
            "algorithmic_bytes": bytes_alg, "launch_ms": t * 1e3, "peak_source": how}


def _traffic(name):
    """dram bytes per launch of the kernel from its committed ncu --set full capture (profiles/)."""
    tp = os.path.join(ROOT, "profiles", name)
    if os.path.exists(tp):
        with open(tp) as f:
            return json.load(f).get("dram_bytes_per_launch")
    return None


def attention_roofline_bwd(dev, iters=10, path=None):
    """Backward of the focal attention at the C3-max shape: fvta_attn_bwd as a whole (accumulator memset,
    per-modality scalars, the tensor-core tile kernel attn_bwd_tc, question / parameter gradient kernels)
    against SURVEY 8(d)'s B_bwd = 8 N K T D (hall read, d_hall written) + masks + small."""
    import torch
    from fvta_memexqa_b200 import ops, _lib
    from fvta_memexqa_b200.config import fvta_config
    from fvta_memexqa_b200.params import ParamStore
    N, K, M, JI, JXP, D, JQ = 4, 6, 64, 64, 8, 100, 32
    T = M * JI * JXP
    cfg = fvta_config(batch_size=N)
    p = ParamStore(cfg, seed=7, device=dev).as_dict()
    g = torch.Generator(device=dev).manual_seed(1)
    hall = torch.randn(N, K, T, D, device=dev, generator=g)
    hmask = torch.ones(N, K, T, dtype=torch.bool, device=dev)
    hq = torch.randn(N, JQ, D, device=dev, generator=g)
    qmask = torch.ones(N, JQ, dtype=torch.bool, device=dev)
    lq = torch.randn(N, D, device=dev, generator=g)
    dg1 = torch.randn(N, D, device=dev, generator=g)
    acfg = ops.AttnConfig(2, True, True, 5, 3.0, False)
    names = ("attention/all/att_logits/W", "attention/all/att_logits/b", "time_warp/WH/W", "time_warp/WH/b",
             "time_warp/WC/W", "time_warp/WC/b")
    args = [p[k] for k in names]
    fwd = ops.attn_forward(acfg, hall, hmask, hq, qmask, lq, *args)
    d_hall, d_hq = torch.empty_like(hall), torch.empty_like(hq)
    sinks = [torch.empty_like(a) for a in args]

    def call():
        ops.attn_backward(fwd, dg1, *args, sinks=sinks, d_hall_out=d_hall, d_hq_out=d_hq, path=path)
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        call()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    n0 = _lib.load().fvta_launch_count()
    with torch.cuda.graph(gr):
        call()
    launches = int(_lib.load().fvta_launch_count() - n0)
    t_end = time.time() + 0.3
    while time.time() < t_end:
        for _ in range(10):
            gr.replay()
        torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            gr.replay()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-4)
    t = statistics.mean(ts)
    bytes_alg = 8 * N * K * T * D + N * K * T + 8 * N * JQ * D + N * JQ + 8 * N * D
    peak, how = measured_peaks()
    return {"bound": "hbm", "achieved": bytes_alg / t / 1e9, "peak": peak, "unit": "GB/s",
            "frac": bytes_alg / t / 1e9 / peak, "traffic": _traffic("attn_bwd_traffic.json"),
            "kernel": "fvta_attn_bwd, whole operator (%d launches; the tile kernel is attn_bwd_tc)" % launches,
            "launch_ms": t * 1e3, "launch_ms_min": min(ts) * 1e3, "algorithmic_bytes": bytes_alg,
            "workload": "C3-max: N=4 K=6 T=32768 D=100 JQ=32, mask density 1.0", "peak_source": how}


def attention_roofline_parity(acfg, p, hall, hmask, hq, qmask, lq, t_slice=1024):
    """The kernel the roofline leg times vs the fp64 collapsed oracle (oracle/ is the checker here, never
    the thing measured): sample 0, first `t_slice` positions, run through the same entry point (the
    slice is a contiguous copy; many tiles per CTA like the full launch)."""
    import torch
    from fvta_memexqa_b200 import ops, _lib
    from fvta_memexqa_b200.config import fvta_config
    from oracle import collapsed
    h = hall[:1, :, :t_slice].contiguous()
    m = hmask[:1, :, :t_slice].contiguous()
    out = ops.attn_forward(acfg, h, m, hq[:1], qmask[:1], lq[:1], p["attention/all/att_logits/W"],
                           p["attention/all/att_logits/b"], p["time_warp/WH/W"], p["time_warp/WH/b"],
                           p["time_warp/WC/W"], p["time_warp/WC/b"], want_logits=True)
    cfg = fvta_config(batch_size=1)
    K, T, D = h.shape[1:]
    p64 = {k: v.detach().double().cpu() for k, v in p.items()}
    ref = collapsed.fvta_attention(h.double().cpu().reshape(1, K, 1, T, D), m.cpu().reshape(1, K, 1, T),
                                   hq[:1].double().cpu(), qmask[:1].cpu(), lq[:1].double().cpu(), p64, cfg)
    err_g = ((out["g1_all"].double().cpu() - ref["g1_all"]).abs().max() / ref["g1_all"].abs().max()).item()
    err_a = (out["a_logits"].double().cpu().reshape(ref["att_logits"].shape) - ref["att_logits"]).abs().max().item()
    ok = err_g < 1e-4 and err_a < 1e-4
    if not ok:
        raise SystemExit("bench.py: attention parity failed at the roofline shape: g1_all %.3e logits %.3e"
                         % (err_g, err_a))
    return {"path": _lib.PATH_NAMES[out["path"]], "g1_all_rel_err": err_g, "a_logits_max_abs_err": err_a,
            "bound": 1e-4, "oracle": "oracle/collapsed.py fp64", "slice": "sample 0, first %d positions" % t_slice}


# ------------------------------------------------------------------------------------------
# own arm
# ------------------------------------------------------------------------------------------
def run_own(args):
    import signal
    import torch
    # a wedged collective must not hold N GPUs until someone else's timeout: give up after 5 min
    signal.signal(signal.SIGALRM, lambda *_: os._exit(17))
    signal.alarm(300 if int(os.environ.get("WORLD_SIZE", "1")) > 1 else 900)
    from fvta_memexqa_b200 import _lib
    from fvta_memexqa_b200.config import fvta_config
    from fvta_memexqa_b200.data import SyntheticDataset
    from fvta_memexqa_b200.dp import DataParallel
    from fvta_memexqa_b200.model import get_model
    from fvta_memexqa_b200.trainer import Trainer

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the FVTA path has no CPU fallback "
                         "(use --impl reference for the CPU oracle)")
    lib = _lib.load()
    dp = DataParallel()
    rank, world = dp.rank, dp.world_size
    local = int(os.environ.get("LOCAL_RANK", "0"))
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()       # nvidia-smi takes ~100 ms to come up: start it before building the model
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dp.device = dev

    cfg = fvta_config(batch_size=PER_GPU_BATCH, seed=1234)
    model = get_model(cfg)
    dp.broadcast_params(model.params.flat)
    trainer = Trainer(model, cfg, dp=dp if world > 1 else None)
    collective = "none (1 GPU)"
    if world > 1:
        fused = (not args.nccl) and dp.enable_p2p(model.params)
        collective = ("fvta_p2p_allreduce_adadelta: gradient sum over NVLink peer memory fused with Adadelta, one launch"
                      if fused else "ncclAllReduce of the flat gradient buffer + fvta_adadelta")
    data = SyntheticDataset(cfg, num_examples=PER_GPU_BATCH * 4, seed=1234 + 1000 * rank, pin=True)
    batches = list(data.get_batches(PER_GPU_BATCH, num_steps=4, shuffle=False))
    # one packed-size bucket for every batch of every rank -> one graph, captured in lockstep
    dp.pin_caps(model, [b[1] for b in batches])
    dfeeds = [model.to_device(model.get_feed_dict(b[1], is_train=True)) for b in batches]
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    h2d = max(model.h2d_bytes(df) for df in dfeeds)
    if not args.no_graph:
        # the whole device step replays as ONE CUDA graph per packed-token bucket (Trainer.enable_graph)
        trainer.enable_graph()
        for df in dfeeds:
            trainer.graph_step(df)          # captures each bucket this data touches, outside the timed region
        device_step = trainer.graph_step
    else:
        device_step = trainer._device_step

    def sync_all():
        torch.cuda.synchronize()
        dp.barrier()
        torch.cuda.synchronize()

    def timed_steps(step_fn, n):
        """n steps, each bracketed by its own CUDA-event pair on the launching stream; the L2 flush
        runs between the pairs.  Returns (seconds inside the pairs, seconds of the whole bracket
        including the flushes, last result)."""
        pairs = []
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        out = None
        b0.record()
        for i in range(n):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = step_fn(i)
            e1.record()
            pairs.append((e0, e1))
        b1.record()
        sync_all()
        per = [a.elapsed_time(b) for a, b in pairs]
        timed_steps.last = per
        return sum(per) * 1e-3, b0.elapsed_time(b1) * 1e-3, out

    def spread(per):
        """min / median / p90 of the per-step times and the mean of each fifth of the run (rank 0)."""
        q = sorted(per)
        k = max(1, len(per) // 5)
        return {"min": q[0], "median": statistics.median(q), "p90": q[int(0.9 * (len(q) - 1))],
                "block_means": [round(statistics.mean(per[i:i + k]), 4) for i in range(0, k * 5, k) if per[i:i + k]]}

    # ---- device-resident leg -------------------------------------------------------------
    for i in range(args.warmup):
        flush.zero_()
        device_step(dfeeds[i % len(dfeeds)])
    sync_all()
    l0 = lib.fvta_launch_count()
    th0 = time.perf_counter()
    t_steps, t_bracket, loss = timed_steps(lambda i: device_step(dfeeds[i % len(dfeeds)]), args.steps)
    dev_spread = spread(timed_steps.last)
    launches = lib.fvta_launch_count() - l0
    if not args.no_graph:
        launches = trainer.launches_per_step * args.steps   # replayed graph nodes
    t_dev = dp.max_over_ranks(t_steps)
    t_dev_bracket = dp.max_over_ranks(t_bracket)
    sampler.mark(th0, time.perf_counter())

    # ---- end-to-end leg: Trainer.step from pinned host tensors, loss read back -----------
    # warm-up with the SAME pipelined pattern as the timed steps (copy stream, staging buffers, the pinned
    # loss slots and the prefetch path are all created on first use)
    use_prefetch = not args.no_prefetch and not args.no_graph
    if use_prefetch:
        trainer.prefetch(batches[0])
    hs = []
    for i in range(max(3, args.warmup)):
        hs.append(trainer.step_handle(None, batches[i % len(batches)]))
        if use_prefetch:
            trainer.prefetch(batches[(i + 1) % len(batches)])
        if len(hs) > 1:
            hs.pop(0).result()
    while hs:
        hs.pop(0).result()
    sync_all()
    pending = {}

    def e2e_step(i):
        # the reference-facing call with HOST buffers.  The input of step i+1 starts its host->device
        # copy on a copy stream while step i runs (Trainer.prefetch); step i waited for ITS copy,
        # issued one step earlier.  The loss of every step is copied back to pinned host memory right
        # behind the step (Trainer.step_handle) and READ one step later, so the device does not idle
        # while the interpreter enqueues the next step.  One H2D copy and one D2H read per step, all
        # inside the bracketed region (the last read happens before the closing synchronisation).
        h = trainer.step_handle(None, batches[i % len(batches)])
        if not args.no_prefetch and not args.no_graph:
            trainer.prefetch(batches[(i + 1) % len(batches)])      # runs under step i's kernels
        prev = pending.pop("h", None)
        pending["h"] = h
        out = (prev.result() if prev is not None else float("nan")), None, None
        if i == args.steps - 1:
            out = (pending.pop("h").result(), None, None)          # the final step's loss
        return out
    if use_prefetch:
        trainer.prefetch(batches[0])
    t_steps, t_bracket, res = timed_steps(e2e_step, args.steps)
    loss_val = res[0]
    e2e_spread = spread(timed_steps.last)
    t_e2e = dp.max_over_ranks(t_steps)
    t_e2e_bracket = dp.max_over_ranks(t_bracket)
    clocks = sampler.stop() if rank == 0 else None
    fresh = None
    if world == 1 and not args.no_graph:
        try:
            fresh = fresh_batch_leg(args, cfg, model, trainer, timed_steps, sync_all)
        except Exception as e:       # (worker processes / shared memory unavailable: the other legs stand)
            fresh = {"error": "%s: %s" % (type(e).__name__, e)}
            getattr(model, "resident", {}).clear()

    if rank != 0:
        _shutdown_dist(world)
        return
    total = PER_GPU_BATCH * world * args.steps
    line = {"metric": METRIC, "value": total / t_dev, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000.0 * t_dev / args.steps, "higher_is_better": True,
            "scaling": "strong" if args.global_batch else "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(world, bool(args.global_batch)), "clocks": clocks,
            "e2e": {"value": total / t_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": 1000.0 * t_e2e / args.steps,
                    "ms_per_step_incl_l2_flush": 1000.0 * t_e2e_bracket / args.steps,
                    "ms_per_step_spread_rank0": e2e_spread},
            "ms_per_step_spread_rank0": dev_spread,
            "ms_per_step_incl_l2_flush": 1000.0 * t_dev_bracket / args.steps,
            "gpu_launches": int(launches), "final_loss": loss_val}
    # how the step ran (not part of `config`: the workload is the same for both arms)
    if fresh is not None:
        line["e2e_fresh_batches"] = fresh
    line["execution"] = "eager launches" if args.no_graph else "one CUDA graph per step"
    line["collective"] = collective
    if world == 1:
        line["roofline"] = attention_roofline(dev)
        line["roofline_with_logits"] = attention_roofline(dev, with_logits=True)
        line["roofline_bwd"] = attention_roofline_bwd(dev)
        if not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            torch.set_num_threads(cores)
            n = 36                      # ~12 s of CPU work on 16 cores
            cpu_reference_step(1, seed=99)
            t = cpu_reference_step(n)
            line["cpu_baseline"] = {
                "value": n / t, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": "%d QAs, micro-batches of 1 (README.MD:146-148), literal torch-CPU oracle of "
                          "model_v2.py (TensorFlow 1.x cannot run here), fwd+bwd+Adadelta, %.1f s" % (n, t)}
    print(json.dumps(line), flush=True)      # the line is out before any teardown can stall
    _shutdown_dist(world)


def fresh_batch_leg(args, cfg, model, trainer, timed_steps, sync_all):
    """The end-to-end figure with the HOST PIPELINE inside the measurement: every step trains on a batch
    that did not exist before the timed region started.  A MemexQA-shaped dataset in the reference's
    on-disk format (data.make_synthetic_raw) is tokenised once (pipeline.TokenizedStore); worker processes
    then assemble, pack and hand over each batch (the vectorised get_feed_dict, model_v2.py:1099-1565)
    while the device trains on the previous ones.  The photo-feature table and the GloVe matrix stay on
    the device (Model.set_resident_tables); a step's host->device copy is its token ids and masks."""
    import torch
    from fvta_memexqa_b200.data import Dataset, make_synthetic_raw, add_vocab
    from fvta_memexqa_b200.pipeline import TokenizedStore, BatchLoader, resident_photo_table
    n_q, n_alb = 4096, 1200
    vocab = 3 * (cfg.word_vocab_size - 2)            # a third of the words are trainable: |word2idx| = word_vocab_size
    data, shared = make_synthetic_raw(n_q, num_albums=n_alb, seed=4242, feat_dim=cfg.image_feat_dim,
                                      emb_dim=cfg.word_emb_size, vocab=vocab)
    add_vocab(shared, cfg)
    assert len(shared["word2idx"]) == cfg.word_vocab_size and len(shared["char2idx"]) <= cfg.char_vocab_size
    t0 = time.perf_counter()
    store = TokenizedStore(Dataset(data, "train", shared=shared), cfg)
    t_tok = time.perf_counter() - t0
    model.set_resident_tables(image_emb_mat=resident_photo_table(store, model.device),
                              existing_emb_mat=store.existing_emb_mat)
    warm = max(3, args.warmup)
    ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 4)
    workers = max(2, min(16, ncpu - 2))
    loader = BatchLoader(store, model, PER_GPU_BATCH, num_steps=warm + args.steps + 1, shuffle=True, seed=7,
                         workers=workers, depth=2 * workers)
    loader.pin_caps()
    loader.start()
    it = iter(loader)
    cur = next(it)
    for _ in range(warm):                            # captures the step graph of this input layout
        nxt = next(it)
        trainer.prefetch(nxt)
        trainer.step(None, cur)
        cur = nxt
    sync_all()
    state = {"cur": cur}

    def step(i):
        h = trainer.step_handle(None, state["cur"])
        nxt = next(it)                               # (blocks if the workers have fallen behind)
        trainer.prefetch(nxt)                        # fetch + host->device copy run under this step's kernels
        state["cur"] = nxt
        prev = state.pop("h", None)
        state["h"] = h
        out = (prev.result() if prev is not None else float("nan")), None, None     # loss read one step late
        if i == args.steps - 1:
            out = (state.pop("h").result(), None, None)
        return out
    h2d = state["cur"][1].step_input(model, True)["blob"].numel()
    w0 = time.perf_counter()
    t_steps, t_bracket, res = timed_steps(step, args.steps)
    wall = time.perf_counter() - w0
    # the same workload with the host pipeline taken out: ONE of these batches, already packed and pinned,
    # copied and trained every step (what the `e2e` key measures on the pre-generated batches)
    last = state["cur"]
    keep = last[1].step_input(model, True)
    from fvta_memexqa_b200.pipeline import HostBatch
    fixed = [(last[0], HostBatch(dict(keep, blob=keep["blob"].clone().pin_memory()), last[0])) for _ in range(2)]
    for i in range(4):
        trainer.step(None, fixed[i % 2])
    sync_all()

    def cached_step(i):
        h = trainer.step_handle(None, fixed[i % 2])
        trainer.prefetch(fixed[(i + 1) % 2])
        prev = state.pop("hc", None)
        state["hc"] = h
        out = (prev.result() if prev is not None else float("nan")), None, None
        if i == args.steps - 1:
            out = (state.pop("hc").result(), None, None)
        return out
    trainer.prefetch(fixed[0])
    t_cached, _, _ = timed_steps(cached_step, args.steps)
    meta = keep["meta"]
    loader.close()
    model.resident.clear()
    total = PER_GPU_BATCH * args.steps
    return {"value": total / t_steps, "unit": UNIT, "ms_per_step": 1000.0 * t_steps / args.steps,
            "ms_per_step_incl_l2_flush": 1000.0 * t_bracket / args.steps,
            "samples_per_s_wall_clock_incl_l2_flush": total / wall,
            "ms_per_step_same_workload_one_cached_batch": 1000.0 * t_cached / args.steps,
            "packed_tokens": {"text_cap": meta["text_cap"], "img_cap": meta["img_cap"], "T": meta["T"]},
            "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 4, "workers": workers,
            "tokenise_once_s": round(t_tok, 3), "final_loss": res[0],
            "what": "every step: a new batch of %d QAs drawn from %d questions / %d albums, assembled + packed "
                    "+ pinned by %d worker processes, copied and trained; photo features and GloVe matrix "
                    "device-resident" % (PER_GPU_BATCH, n_q, n_alb, workers)}


def _shutdown_dist(world):
    """Tear the process group down without letting a slow NCCL teardown hold the GPUs: the
    measurement is finished, so after a short grace period the process simply exits."""
    if world <= 1:
        return
    import torch.distributed as dist
    sys.stdout.flush()
    t = threading.Thread(target=lambda: dist.destroy_process_group(), daemon=True)
    t.start()
    t.join(timeout=10.0)
    os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)   # ~0.12 s timed: enough for several clock samples
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", type=str, default="fvta_b200", choices=["fvta_b200", "reference"])
    ap.add_argument("--no-cpu-baseline", dest="no_cpu_baseline", action="store_true")
    ap.add_argument("--no-graph", dest="no_graph", action="store_true", help="launch kernels eagerly")
    ap.add_argument("--no-prefetch", dest="no_prefetch", action="store_true",
                    help="e2e leg without Trainer.prefetch (the step input is copied in front of each step)")
    ap.add_argument("--nccl", action="store_true",
                    help="data parallel with ncclAllReduce + fvta_adadelta instead of the fused peer-memory kernel (A/B)")
    ap.add_argument("--global-batch", dest="global_batch", type=int, default=0,
                    help="strong scaling (SURVEY C4): fix the GLOBAL batch and give each GPU global/N QAs "
                         "(default: weak scaling, 32 per GPU)")
    args = ap.parse_args()
    if args.global_batch:
        global PER_GPU_BATCH
        n = max(1, args.gpus, int(os.environ.get("WORLD_SIZE", "1")))
        if args.global_batch % n:
            raise SystemExit("bench.py: --global-batch must divide by the GPU count")
        PER_GPU_BATCH = args.global_batch // n
    if args.warmup < 3 and args.impl != "reference":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_own(args)


if __name__ == "__main__":
    main()
