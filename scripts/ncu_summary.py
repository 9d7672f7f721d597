#!/usr/bin/env python
"""Condense ncu outputs into the small text files kept under profiles/.

    python scripts/ncu_summary.py launches gpurun_out/x_launches.csv  > profiles/x_launch_shares.txt
    python scripts/ncu_summary.py report   gpurun_out/x.ncu-rep       > profiles/x_ncu.txt
"""
import csv
import io
import re
import subprocess
import sys
from collections import OrderedDict

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_static", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.sum", "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
]


def launches(path):
    rows = [r for r in csv.reader(open(path, newline="")) if len(r) > 10]
    hdr = rows[0]
    kn, mv, mn = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    agg = OrderedDict()
    for r in rows[1:]:
        if r[mn] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r[kn])
        name = re.sub(r"<.*", "<>", name)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[mv].replace(",", ""))
    tot = sum(a[1] for a in agg.values())
    print(f"# {path}: {sum(a[0] for a in agg.values())} launches, {tot / 1e3:.1f} us total (ncu serialised, cold cache: read SHARES)")
    print(f"{'kernel':58s} {'launches':>8s} {'total us':>10s} {'avg us':>9s} {'share':>7s}")
    for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{name[:58]:58s} {c:8d} {t / 1e3:10.1f} {t / c / 1e3:9.2f} {100 * t / tot:6.1f}%")


def report(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    kn = hdr.index("Kernel Name")
    print(f"# {path} (ncu --set full --clock-control none; per launch)")
    for r in rows[2:]:
        print(f"## {r[kn][:100]}")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"{k:90s} {r[i]:>16s} {units[i]}")
        print()


if __name__ == "__main__":
    {"launches": launches, "report": report}[sys.argv[1]](sys.argv[2])
