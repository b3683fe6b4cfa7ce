#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into profiles/ (tracked).
usage: tools/ncu_summary.py <launches.csv> <report.ncu-rep> <out_prefix>"""
import collections
import csv
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.avg", "sm__cycles_active.avg",
        "sm__inst_executed.sum", "smsp__inst_executed.sum", "sm__sass_inst_executed_op_shared_ld.sum",
        "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_not_selected_per_warp_active.pct",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct"]


def launches(path, out):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    h = rows[0]
    iname, ival, iunit = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    scale = {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9, "second": 1e9, "nsecond": 1.0, "usecond": 1e3, "msecond": 1e6}
    for r in rows[1:]:
        v = float(r[ival].replace(",", "")) * scale.get(r[iunit], 1.0)
        agg[r[iname]][0] += 1
        agg[r[iname]][1] += v
    tot = sum(v[1] for v in agg.values())
    with open(out, "w") as f:
        f.write("# per-kernel device time from `ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: compare SHARES)\n")
        f.write("kernel\tlaunches\ttotal_us\tshare_pct\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%s\t%d\t%.1f\t%.2f\n" % (k, v[0], v[1] / 1e3, 100 * v[1] / tot))


def full(rep, out):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    h, u = rows[0], rows[1]
    with open(out, "w") as f:
        f.write("# selected metrics of `ncu --set full --clock-control none` (%s)\n" % rep.split("/")[-1])
        for row in rows[2:]:
            f.write("## kernel: %s\n" % row[h.index("Kernel Name")])
            for i, n in enumerate(h):
                if n in KEEP:
                    f.write("%s\t%s\t%s\n" % (n, row[i], u[i]))


if __name__ == "__main__":
    launches(sys.argv[1], sys.argv[3] + "_launches.tsv")
    full(sys.argv[2], sys.argv[3] + "_ncu_full.tsv")
