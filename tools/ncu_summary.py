#!/usr/bin/env python
"""Summarise ncu outputs into small text files for profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches.csv          > profiles/rNN_launches.txt
    python tools/ncu_summary.py kernel   gpurun_out/prof.ncu-rep [regex]  > profiles/rNN_kernel.txt
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max", "sass__inst_executed_local_loads",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__pipe_tensor_cycles_active",
    "smsp__average_warps_issue_stalled", "smsp__warp_issue_stalled", "smsp__pcsamp_warps_issue_stalled",
]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        n = r[ki].split("(")[0]
        v = float(r[vi].replace(",", ""))
        v = v / 1000 if r[ui] == "ns" else v * 1000 if r[ui] == "ms" else v
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)")
    for n, (c, t) in agg.items():
        print("%-58s n=%4d total=%10.1f us avg=%9.2f us share=%5.1f%%" % (n[:58], c, t, t / c, 100 * t / tot))
    print("total %.1f us over %d launches" % (tot, sum(a[0] for a in agg.values())))


def kernel(path, pattern=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    body = [r for r in rows[2:] if pattern is None or re.search(pattern, r[ki])]
    print("# ncu --set full --clock-control none: %s" % path)
    for r in body:
        print("## " + r[ki][:110])
        for i, h in enumerate(hdr):
            if any(h.startswith(k) for k in KEYS) and r[i] not in ("", "0", "n/a"):
                print("   %-85s %-14s %s" % (h, units[i], r[i]))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        kernel(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
