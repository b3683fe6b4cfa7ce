#!/usr/bin/env python
"""The north star's criterion (<= 4 mismatches / <= 1 gap) at the north star's gRNA count on ONE 135 Mb genome: the
seed-and-join screen against the bit-sliced DP path.  Prints one JSON line per gRNA count: seconds of the seed-and-join screen
over the whole genome, of the DP path on a 1/64 window range (extrapolated), and whether the two agree count for count on that
range.   python tools/bench_seedjoin.py [n_grna ...]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench as B  # noqa: E402
from minorg_b200 import _lib  # noqa: E402
from minorg_b200.ot_regex import make_criteria  # noqa: E402


def main():
    frac = 1.0                                   # --frac F: time the seed-and-join screen on that fraction of the genome and scale
    argv = sys.argv[1:]
    if "--frac" in argv:
        i = argv.index("--frac")
        frac = float(argv[i + 1])
        del argv[i:i + 2]
    ungapped = "--ungapped" in argv                  # Hamming <= 4 instead of <= 4 mismatches / <= 1 gap (the reference scan is then the pair-word kernel)
    if ungapped:
        argv.remove("--ungapped")
    counts_n = [int(x) for x in argv] or [1_000_000]
    name, sms = _lib.init(0)
    dev = torch.device("cuda", 0)
    stream = torch.cuda.current_stream().cuda_stream
    clens = B.genome_layout(135_000_000)
    targets = B.make_targets()
    ascii_dev = B.synth_genome_device(torch, 0, clens, targets[0], dev)
    offsets = np.concatenate([[0], np.cumsum(clens)]).astype(np.int64)
    genome = _lib.Genome.from_device_ascii(ascii_dev.data_ptr(), offsets, ["c%d" % i for i in range(len(clens))], stream=stream)
    torch.cuda.synchronize()
    found = genome.find_exact([s for _, s in targets[0]], stream=stream)
    genome.set_masks(sorted({(c, s, e) for _, c, s, e, _ in found}), stream=stream)
    crit = make_criteria(ot_mismatch=5, ot_gap=0 if ungapped else 2, grna_length=20)
    rng = np.random.Generator(np.random.PCG64(1000))
    total = sum(clens)
    for n in counts_n:
        base = B.choose_grnas(_lib, targets, min(n, 10000))
        extra = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=(max(0, n - len(base)), 20))]
        grnas = base + [bytes(r).decode() for r in extra]
        q = _lib.Queries(grnas, stream=stream)
        c = torch.zeros(2 * n, dtype=torch.int32, device=dev)

        def run(mode, wb=0, we=-1):
            os.environ["MG_GAPPED"] = mode
            c.zero_()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.screen(genome, q, crit, None, c.data_ptr(), win_begin=wb, win_end=we, stream=stream)
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / 1e3, c.clone()
        run("seedjoin", 0, 1 << 20)                                  # builds the index, warms up (the two launches an ncu capture takes)
        warm_stats = _lib.last_seedjoin_stats()
        glen = genome.global_length
        if frac >= 1.0:
            t_sj, c_sj = run("seedjoin")
        else:
            t_part, c_sj = run("seedjoin", glen // 4, glen // 4 + int(glen * frac))
            t_sj = t_part / frac
        cand, surv = _lib.last_seedjoin_stats()
        a, b = glen // 2, glen // 2 + glen // 64
        t_sj_part, c_sj_part = run("seedjoin", a, b)
        t_dp_part, c_dp_part = run("dp", a, b)
        equal = bool(torch.equal(c_sj_part, c_dp_part))
        print(json.dumps({"n_grna": n, "genome_bp": total, "criterion": "<=4 mismatches, no gap, both strands, PAM-less (reference scan: pair-word kernel)" if ungapped else "<=4 mismatches / <=1 gap, both strands, no pattern, PAM-less",
                          "seedjoin_seconds": t_sj, "timed_fraction_of_genome": frac, "seedjoin_tcell_per_s": 2.0 * n * total / t_sj / 1e12,
                          "anchors_total": int(c_sj.sum().item()), "candidates": cand, "stage2_candidates": surv,
                          "candidates_per_cell": cand / (2.0 * n * total * min(frac, 1.0)),
                          "warmup_range_candidates": warm_stats[0], "warmup_range_stage2_candidates": warm_stats[1],
                          "range_windows": b - a, "seedjoin_seconds_on_range": t_sj_part, "dp_seconds_on_range": t_dp_part,
                          "dp_seconds_whole_genome_extrapolated": t_dp_part * glen / (b - a),
                          "speedup_on_range": t_dp_part / t_sj_part, "counts_equal_on_range": equal,
                          "north_star_seconds_on_8_gpus_if_1M": t_sj * (1_000_000 / n) if n >= 1_000_000 else None, "device": name}))
        sys.stdout.flush()
        q.free()
    os.environ.pop("MG_GAPPED", None)


if __name__ == "__main__":
    main()
