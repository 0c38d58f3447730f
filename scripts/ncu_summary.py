#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, on the CPU box): key raw metrics per launch + hottest source lines.
Usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_<name>.txt"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio"]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    name_i = hdr.index("Kernel Name")
    for r in rows[2:]:
        print("== " + r[name_i])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print("   %-90s %s %s" % (k, r[i], units[i]))
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    if not rows:
        return
    # find header row
    for hi, r in enumerate(rows):
        if "Source" in r and any("Sampling" in c for c in r):
            break
    else:
        return
    hdr = rows[hi]
    si = hdr.index("Source")
    samp = [i for i, c in enumerate(hdr) if c.startswith("# Samples") or c == "Warp Stall Sampling (All Samples)"]
    if not samp:
        samp = [i for i, c in enumerate(hdr) if "Sampling (All" in c]
    if not samp:
        return
    col = samp[0]
    items = []
    for r in rows[hi + 1:]:
        try:
            items.append((float(r[col] or 0), r[si].strip()))
        except (ValueError, IndexError):
            continue
    total = sum(v for v, _ in items) or 1.0
    print("== hottest source lines by warp-stall samples (first profiled launch)")
    for v, s in sorted(items, reverse=True)[:18]:
        print("   %5.1f%%  %s" % (100.0 * v / total, s[:150]))


if __name__ == "__main__":
    main(sys.argv[1])
