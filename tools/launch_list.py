#!/usr/bin/env python
"""Aggregate an ncu launch list (`--metrics gpu__time_duration.sum --csv --log-file X.csv`) per kernel.

    python tools/launch_list.py gpurun_out/r01/launches.csv gpurun_out/r01/bench.json > profiles/r01_launch_list.md
"""
import csv
import json
import re
import sys

PREP = ("k_gap_columns", "k_gap_flags", "k_pair_diag_means", "k_pair_compact", "k_pair_means", "k_px_", "k_dense_out", "k_gap_band", "k_fp64_peak")


def main():
    path, bench = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else None)
    rows = [r for r in csv.reader(open(path)) if len(r) > 14 and r[0].isdigit()]
    agg = {}
    for r in rows:
        name = re.sub(r"^void ", "", r[4].split("(")[0])
        t = float(r[14]) / (1e3 if r[13] == "ns" else 1.0)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += t
    total = sum(v[1] for v in agg.values())
    path_total = sum(v[1] for k, v in agg.items() if not k.startswith(PREP))
    print("# ncu launch list\n")
    cmd = sys.argv[3] if len(sys.argv) > 3 else "python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
    print("`ncu --metrics gpu__time_duration.sum --clock-control none -c N --csv %s` (the plain command exited 0 first). "
          "The first N launches of the process: untimed preparation (matrix assembly and normalisation from the pixel "
          "table), warm-up and timed steps. Per-launch times are cold-cache and serialised: compare SHARES with the "
          "CUDA-event shares of the bench line, not absolute times.\n" % cmd)
    if bench:
        d = json.loads(open(bench).read().strip().splitlines()[-1])
        k = d["kernels"]
        tot = sum(v["ms_per_step"] for v in k.values())
        k = {n: v for n, v in k.items() if n != "score"}       # the score class is the feature + forest region, not a kernel
        tot = sum(v["ms_per_step"] for v in k.values())
        print("CUDA-event shares of the same command's bench line: " +
              ", ".join("%s %.2f ms = %.1f %%" % (n, v["ms_per_step"], 100 * v["ms_per_step"] / tot) for n, v in k.items()) +
              " of %.2f ms kernel time per step.\n" % tot)
    print("| kernel | launches | total us | share of all | share of scoring-path kernels |")
    print("|---|---:|---:|---:|---:|")
    for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        sp = "(preparation)" if name.startswith(PREP) else "%.1f%%" % (100 * t / path_total)
        print("| `%s` | %d | %.1f | %.1f%% | %s |" % (name, n, t, 100 * t / total, sp))


if __name__ == "__main__":
    main()
