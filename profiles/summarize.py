"""Turns an ncu launch list (CSV) and a full capture (.ncu-rep) into the text summaries committed
under profiles/.  Run here (no GPU needed):  python profiles/summarize.py <tag> <launches.csv> <prof.ncu-rep>"""
import csv
import subprocess
import sys
from collections import defaultdict

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__warps_eligible.avg.per_cycle_active", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
]


def launches(path, out):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    h = rows[0]
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    d = defaultdict(list)
    for r in rows[1:]:
        d[r[ki]].append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in d.values())
    out.write("## launch list (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised: compare SHARES)\n\n")
    out.write("| launches | avg ms | share | kernel |\n|---:|---:|---:|---|\n")
    for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
        out.write(f"| {len(v)} | {sum(v)/len(v)/1e6:.3f} | {100*sum(v)/tot:.1f}% | `{k[:120]}` |\n")
    out.write("\n")


def full(path, out):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, units = rows[0], rows[1]
    out.write("## full capture (ncu --set full --clock-control none --import-source on)\n\n")
    for r in rows[2:]:
        out.write(f"### `{r[h.index('Kernel Name')][:140]}`\n\n| metric | value | unit |\n|---|---:|---|\n")
        for k in KEYS:
            if k in h:
                out.write(f"| {k} | {r[h.index(k)]} | {units[h.index(k)]} |\n")
        out.write("\n")


if __name__ == "__main__":
    tag, lcsv, rep = sys.argv[1:4]
    with open(f"profiles/{tag}_summary.md", "w") as f:
        f.write(f"# ncu summary — {tag}\n\n")
        if lcsv != "-":
            launches(lcsv, f)
        if rep != "-":
            full(rep, f)
    print(open(f"profiles/{tag}_summary.md").read())
