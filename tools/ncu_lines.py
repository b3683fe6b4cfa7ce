#!/usr/bin/env python
"""Instruction and stall-sample share per source line of an ncu report taken with --import-source on (-lineinfo build).
   python tools/ncu_lines.py report.ncu-rep [min_permille]"""
import collections
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    thr = float(sys.argv[2]) if len(sys.argv) > 2 else 5.0
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    cur, hdr = None, None
    agg = collections.defaultdict(lambda: [0, 0, ""])
    for r in csv.reader(out.splitlines()):
        if len(r) == 2 and r[0] in ("File Path", "File Name"):
            cur = r[1].split("/")[-1]
            continue
        if r and r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) < 8 or "Instructions Executed" not in hdr:
            continue
        try:
            ln, inst, samp = int(r[0]), int(r[hdr.index("Instructions Executed")]), int(r[hdr.index("# Samples")])
        except ValueError:
            continue
        a = agg[(cur, ln)]
        a[0] += inst
        a[1] += samp
        a[2] = r[1][:100].strip()
    tot = sum(v[0] for v in agg.values()) or 1
    ts = sum(v[1] for v in agg.values()) or 1
    for k, v in sorted(agg.items()):
        if v[0] * 1000 > thr * tot or v[1] * 1000 > thr * ts:
            print("%-22s %5d inst %5.2f%% samp %5.2f%%  %s" % (k[0], k[1], 100 * v[0] / tot, 100 * v[1] / ts, v[2]))


if __name__ == "__main__":
    main()
