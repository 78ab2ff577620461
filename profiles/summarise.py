#!/usr/bin/env python
"""Turn the raw ncu outputs that gpurun brings back (gpurun_out/, scratch) into the small text
summaries committed under profiles/.

    python profiles/summarise.py launches gpurun_out/launches_X.csv  > profiles/X_launches.txt
    python profiles/summarise.py kernel   gpurun_out/prof_X.ncu-rep  > profiles/X_kernel.txt
"""
import collections
import csv
import io
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size",
        "launch__block_size", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed_op_shared_atom.sum",
        "l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum", "lts__t_bytes.sum", "sm__cycles_elapsed.avg",
        "smsp__cycles_active.avg", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__average_warp_latency_issue_stalled_barrier.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1.0)
        a = agg.setdefault(r[ki], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)")
    print("# source: %s   launches: %d   total: %.1f us" % (path, sum(a[0] for a in agg.values()), tot))
    print("%-72s %7s %14s %7s" % ("kernel", "count", "time_us", "share"))
    for n, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("%-72s %7d %14.1f %6.1f%%" % (n[:72], a[0], a[1], 100 * a[1] / tot))


def kernel(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    ni = hdr.index("Kernel Name")
    print("# ncu --set full --clock-control none --import-source on; source: %s" % path)
    for r in rows[2:]:
        print("## launch id %s: %s  grid %s block %s" % (r[0], r[ni][:90], r[hdr.index("Grid Size")], r[hdr.index("Block Size")]))
        for k in KEEP:
            if k in hdr:
                i = hdr.index(k)
                print("  %-84s %14s %s" % (k, r[i], units[i]))


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2])
