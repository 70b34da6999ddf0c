#!/usr/bin/env python
"""Per-kernel summary of one step from an ncu launch list (gpu__time_duration.sum CSV).

    python tools/launch_summary.py gpurun_out/launches_r1c_bench_eager.csv [--list]

A step is delimited by consecutive adadelta_kernel launches; the last complete one is reported."""
import collections
import csv
import re
import sys


def load(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = []
    for x in csv.DictReader(lines):
        if x.get("Metric Name") == "gpu__time_duration.sum":
            v = float(x["Metric Value"].replace(",", ""))
            u = x["Metric Unit"]
            v = v / 1000 if u == "ns" else (v * 1000 if u == "ms" else v)
            rows.append((x["Kernel Name"], v, x["Grid Size"], x["Block Size"]))
    return rows


def main():
    rows = load(sys.argv[1])
    idx = [i for i, r in enumerate(rows) if "adadelta" in r[0]]
    a, b = idx[-2] + 1, idx[-1] + 1
    step = rows[a:b]
    if "--list" in sys.argv:
        for k, v, g, bl in step:
            print("%7.1f %16s %14s %s" % (v, g, bl, k[:110]))
        return
    agg, cnt = collections.Counter(), collections.Counter()
    for k, v, _, _ in step:
        k = re.sub(r"\(.*", "", re.sub(r"<.*", "", k))
        agg[k] += v
        cnt[k] += 1
    tot = sum(agg.values())
    own = sum(v for k, v in agg.items() if "fvta::" in k)
    print("step: %d launches, %.1f us (cold, serialised); fvta:: kernels %.1f us (%.0f%%)"
          % (len(step), tot, own, 100 * own / tot))
    for k, v in agg.most_common(30):
        print("%8.1f %5.1f%% x%3d %s" % (v, 100 * v / tot, cnt[k], k[:90]))


if __name__ == "__main__":
    main()
