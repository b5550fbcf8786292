#!/usr/bin/env python
"""Turn an ncu report (--set full) into the markdown summary kept under profiles/ and refresh
profiles/<round>_traffic.json for k_features.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_ncu_summary.md \
        --command "<the ncu command line>" [--traffic profiles/r01_traffic.json --workload "..."]

Reads the report with `ncu -i REP --page raw --csv` (works without a GPU).
"""
import argparse
import csv
import io
import json
import subprocess

METRICS = [
    ("duration", "gpu__time_duration.sum"),
    ("grid", "launch__grid_size"),
    ("block", "launch__block_size"),
    ("regs/thread", "launch__registers_per_thread"),
    ("dyn smem/block", "launch__shared_mem_per_block_dynamic"),
    ("static smem/block", "launch__shared_mem_per_block_static"),
    ("DRAM read", "dram__bytes_read.sum"),
    ("DRAM write", "dram__bytes_write.sum"),
    ("DRAM throughput % of peak", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("SM throughput %", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("issue active %", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("fp64 pipe inst % of peak", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
    ("LSU shared wavefronts % of peak", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
    ("smem bank conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
    ("smem wavefronts", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
    ("achieved occupancy %", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("active threads / inst", "smsp__thread_inst_executed_per_inst_executed.ratio"),
    ("warp instructions", "smsp__inst_executed.sum"),
    ("stall wait / issue", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"),
    ("stall short scoreboard / issue", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"),
    ("stall long scoreboard / issue", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
    ("stall barrier / issue", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"),
    ("stall math pipe throttle / issue", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"),
    ("stall not selected / issue", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"),
    ("stall branch resolving / issue", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"),
]

TO_BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def read_report(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    header, units = rows[start], rows[start + 1]
    return header, units, rows[start + 2:]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("out_md")
    ap.add_argument("--command", default="")
    ap.add_argument("--title", default="ncu --set full summaries")
    ap.add_argument("--traffic", default=None)
    ap.add_argument("--workload", default="")
    ap.add_argument("--assemblies", type=int, default=50)
    a = ap.parse_args()
    header, units, rows = read_report(a.report)
    col = {h: i for i, h in enumerate(header)}
    kname = col["Kernel Name"]
    seen = {}
    for r in rows:
        name = r[kname].split("(")[0].replace("void ", "")
        seen.setdefault(name, []).append(r)
    lines = ["# " + a.title, ""]
    if a.command:
        lines += ["Command: `%s`" % a.command, ""]
    lines += ["Numbers taken under ncu are not bench values (cold caches, serialised launches).", ""]
    traffic = None
    for name, rs in seen.items():
        r = max(rs, key=lambda x: float(x[col["gpu__time_duration.sum"]].replace(",", "")))
        lines += ["## `%s` (%d launch(es) captured, the longest shown)" % (name, len(rs)), "", "| metric | value |", "|---|---|"]
        for label, m in METRICS:
            if m in col:
                lines.append("| %s (`%s`) | %s %s |" % (label, m, r[col[m]], units[col[m]]))
        lines.append("")
        if name.startswith("k_features") and a.traffic:
            rd = float(r[col["dram__bytes_read.sum"]].replace(",", "")) * TO_BYTES[units[col["dram__bytes_read.sum"]]]
            wr = float(r[col["dram__bytes_write.sum"]].replace(",", "")) * TO_BYTES[units[col["dram__bytes_write.sum"]]]
            traffic = {"kernel": "k_features", "assemblies": a.assemblies, "workload": a.workload, "dram_bytes_read": rd, "dram_bytes_write": wr,
                       "dram_bytes_per_launch": rd + wr,
                       "source": "ncu --set full --clock-control none, one launch of k_features (%s)" % a.out_md}
    open(a.out_md, "w").write("\n".join(lines))
    if traffic:
        json.dump(traffic, open(a.traffic, "w"), indent=1)
        print("traffic", traffic)


if __name__ == "__main__":
    main()
