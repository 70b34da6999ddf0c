#!/usr/bin/env python
"""Summarise one kernel of an ``ncu --set full`` report into the small JSON kept under profiles/.

    python tools/ncu_summary.py gpurun_out/attn_fwd_tc_r1b.ncu-rep profiles/attn_fwd_tc_r1_ncu_full.json \
        [--traffic profiles/attn_fwd_traffic.json --algorithmic-bytes 315413760]
"""
import argparse
import csv
import io
import json
import subprocess

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__block_size",
    "launch__grid_size", "launch__shared_mem_per_block_dynamic", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_alu.sum",
    "sm__inst_executed_pipe_xu.sum", "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
]

UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("out")
    ap.add_argument("--traffic")
    ap.add_argument("--algorithmic-bytes", type=int, default=0)
    ap.add_argument("--all", action="store_true", help="every kernel of the report: the output is a list")
    ap.add_argument("--kernel", default=None, help="name written into the traffic file")
    a = ap.parse_args()
    txt = subprocess.run(["ncu", "-i", a.rep, "--page", "raw", "--csv"], check=True, capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    names, units = rows[0], rows[1]
    col = {n: i for i, n in enumerate(names)}

    def one(vals):
        o = {"Kernel Name": {"value": vals[col["Kernel Name"]], "unit": ""}}
        for k in KEEP:
            if k in col:
                o[k] = {"value": vals[col[k]], "unit": units[col[k]]}
        o["source"] = "ncu --set full --clock-control none, %s" % a.rep
        return o
    vals = rows[2]
    out = [one(v) for v in rows[2:] if len(v) == len(names)] if a.all else one(vals)
    with open(a.out, "w") as f:
        json.dump(out, f, indent=1)
    if a.traffic:
        def b(k):
            return int(float(vals[col[k]].replace(",", "")) * UNIT[units[col[k]]])
        rd, wr = b("dram__bytes_read.sum"), b("dram__bytes_write.sum")
        with open(a.traffic, "w") as f:
            json.dump({"kernel": a.kernel or vals[col["Kernel Name"]][:60], "dram_bytes_per_launch": rd + wr, "dram_bytes_read": rd,
                       "dram_bytes_write": wr, "algorithmic_bytes": a.algorithmic_bytes,
                       "source": "ncu --set full, %s (C3-max shape)" % a.rep}, f, indent=1)
    print(json.dumps(out if not a.all else [o["Kernel Name"]["value"][:70] for o in out], indent=1))


if __name__ == "__main__":
    main()
