#!/usr/bin/env python
"""Summarise ncu outputs brought back from the GPU box into small text files for profiles/.

  python tools/summarize_ncu.py launches gpurun_out/launches_r1.csv  > profiles/r1_launches.md
  python tools/summarize_ncu.py full     gpurun_out/prof_r1.ncu-rep  > profiles/r1_kernels.md
"""
import collections
import csv
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__occupancy_limit_warps",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_shared_atom.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, mi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if r[mi] != "gpu__time_duration.sum":
            continue
        agg.setdefault(r[ki].split("(")[0], []).append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    print("| kernel | launches | total ms | share | mean us |")
    print("|---|---|---|---|---|")
    for k, v in agg.items():
        print(f"| `{k}` | {len(v)} | {sum(v)/1e6:.3f} | {100*sum(v)/tot:.1f}% | {sum(v)/len(v)/1e3:.1f} |")
    print(f"\n(source: `{path}`; ncu --metrics gpu__time_duration.sum --clock-control none; per-launch times are"
          " cold-cache and serialised -- compare shares, not absolutes)")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0]
        print(f"### `{name}` (launch id {r[hdr.index('ID')]})\n")
        print("| metric | value | unit |\n|---|---|---|")
        for m in KEEP:
            if m in hdr:
                i = hdr.index(m)
                print(f"| {m} | {r[i]} | {units[i]} |")
        print()
    print(f"(source: `{path}`; ncu --set full --clock-control none --import-source on)")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
