"""CPU checks of the measurement contract (bench.py) that do not need a GPU: the reference arm's
JSON line, the own arm's loud failure without CUDA, and the committed evidence under profiles/."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config"}


def _run(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          cwd=ROOT, env=e, timeout=900)


def test_reference_arm_line():
    """`bench.py --impl reference`: one JSON line, the contract's keys, the CPU oracle as the thing
    measured (kind "port": TensorFlow 1.x cannot run here), zero GPU launches."""
    out = _run("--impl", "reference", "--steps", "1", "--warmup", "1")
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert BASE_KEYS <= set(line)
    assert line["impl"] == "reference" and line["gpu_launches"] == 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0,
                           "d2h_bytes_per_step": 0}
    assert line["value"] > 0 and line["higher_is_better"] is True and line["vs_baseline"] is None
    assert "workload" in line["config"] and "model" not in line["config"]
    with open(os.path.join(ROOT, "BASELINE.json")) as f:
        base = json.load(f)
    # BASELINE.json names two quantities; the bench line's metric is the throughput one, the
    # roofline fraction travels in the `roofline` object
    assert base["metric"].startswith(line["metric"])


def test_reference_arm_other_ranks_exit_quietly():
    """Under torchrun only rank 0 runs the reference arm; the other ranks exit 0 without output."""
    out = _run("--impl", "reference", "--steps", "1", "--warmup", "1", env={"RANK": "1", "WORLD_SIZE": "2"})
    assert out.returncode == 0 and out.stdout.strip() == ""


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_own_arm_refuses_to_run_without_cuda():
    """The product path has no CPU fallback: without a device the bench fails loudly."""
    out = _run("--steps", "1", "--warmup", "3")
    assert out.returncode != 0
    assert "no CPU fallback" in (out.stderr + out.stdout)


@pytest.mark.parametrize("name", ["bench_r1_dev.json", "bench_r2_dev.json"])
def test_committed_bench_line_is_complete(name):
    """profiles/bench_r*_dev.json are lines the driver's checks would accept."""
    with open(os.path.join(ROOT, "profiles", name)) as f:
        line = json.loads(f.read().strip().splitlines()[-1])
    assert BASE_KEYS | {"clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"} <= set(line)
    if name == "bench_r2_dev.json":
        b = line["roofline_bwd"]
        assert b["bound"] == "hbm" and abs(b["frac"] - b["achieved"] / b["peak"]) < 1e-9 and b["frac"] >= 0.3
        assert line["roofline"]["frac"] >= 0.45 and line["roofline"]["parity"]["g1_all_rel_err"] < 1e-4
        f2 = line["e2e_fresh_batches"]
        assert "error" not in f2 and f2["h2d_bytes_per_step"] > 0
        assert f2["ms_per_step"] <= 1.1 * f2["ms_per_step_same_workload_one_cached_batch"]   # the review's bar
        assert {"min", "median", "p90", "block_means"} <= set(line["ms_per_step_spread_rank0"])
    r = line["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s"
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["traffic"] is None or r["traffic"] >= 0.95 * r["algorithmic_bytes"]
    assert line["gpu_launches"] > 0 and line["e2e"]["h2d_bytes_per_step"] > 0
    assert line["e2e"]["value"] <= line["value"] * 1.02      # host copies can only cost time
    assert not set(line["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


@pytest.mark.parametrize("name", ["launches_r1_bench_eager.csv", "launches_r2_bench_eager.csv"])
def test_committed_launch_list_summarises(name):
    """profiles/launches_r*_bench_eager.csv parse into one complete step whose time is dominated by
    this library's kernels (the evidence that the step is not a torch fallback)."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "launch_summary.py"),
                          os.path.join(ROOT, "profiles", name)],
                         capture_output=True, text=True, cwd=ROOT, timeout=120)
    assert out.returncode == 0, out.stderr[-1000:]
    head = out.stdout.splitlines()[0]
    assert head.startswith("step:") and "fvta:: kernels" in head
    pct = int(head.rsplit("(", 1)[1].split("%")[0])
    assert pct >= 75, head
    assert "fvta::attn_fwd_tc" in out.stdout and "fvta::gemm_mma_kernel" in out.stdout
    if "r2" in name:
        assert "fvta::attn_bwd_tc" in out.stdout
