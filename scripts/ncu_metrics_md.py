#!/usr/bin/env python
"""Tabulate selected `ncu --page raw` metrics of several reports side by side (markdown).

    python scripts/ncu_metrics_md.py label1=path1.ncu-rep label2=path2.ncu-rep ... > profiles/x.md
"""
import csv
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "smsp__sass_average_branch_targets_threads_uniform.pct",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
]


def load(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    names, units, vals = rows[0], rows[1], rows[2]
    d = {}
    for n, u, v in zip(names, units, vals):
        d[n] = (u, v)
    return d


def fmt(v):
    try:
        x = float(v.replace(",", ""))
    except ValueError:
        return v
    return f"{x:.4g}"


def main():
    cols = [a.split("=", 1) for a in sys.argv[1:]]
    data = [(lab, load(path)) for lab, path in cols]
    print("| metric | unit | " + " | ".join(lab for lab, _ in data) + " |")
    print("|---|---|" + "---|" * len(data))
    for m in METRICS:
        units = {d[m][0] for _, d in data if m in d}
        if len(units) > 1:   # ncu scales byte units per report: keep each cell's own unit
            print(f"| `{m}` | (per cell) | " + " | ".join(f"{fmt(d[m][1])} {d[m][0]}" if m in d else "-" for _, d in data) + " |")
        else:
            print(f"| `{m}` | {next(iter(units), '')} | " + " | ".join(fmt(d[m][1]) if m in d else "-" for _, d in data) + " |")


if __name__ == "__main__":
    main()
