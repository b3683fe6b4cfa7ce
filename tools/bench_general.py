#!/usr/bin/env python
"""Throughput of the non-default screen modes (gapped / PAM proximity / patterns) on one GPU.
Not the headline bench (bench.py); numbers are quoted in DESIGN.md.   python tools/bench_general.py"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from minorg_b200 import _lib  # noqa: E402
from minorg_b200.ot_regex import make_criteria  # noqa: E402
from minorg_b200.pam import PAM  # noqa: E402


def main():
    _lib.init(0)
    dev = torch.device("cuda", 0)
    stream = torch.cuda.current_stream().cuda_stream
    gen = torch.Generator(device=dev)
    gen.manual_seed(7)
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
    cases = [
        ("ungapped <=4 mm, PAM-less (fast path)", 135e6, 10000, 20, dict(ot_mismatch=5), ".NGG"),
        ("reference defaults: ot_mismatch=1 (exact matches), PAM-less", 135e6, 10000, 20, dict(ot_mismatch=1), ".NGG"),
        ("ungapped <=2 mm, PAM-less", 135e6, 10000, 20, dict(ot_mismatch=3), ".NGG"),
        ("ungapped <=4 mm + PAM proximity (.NGG)", 135e6, 10000, 20, dict(ot_mismatch=5, ot_pamless=False), ".NGG"),
        ("ungapped <=4 mm + PAM proximity (TTTV. L=23)", 135e6, 10000, 23, dict(ot_mismatch=5, ot_pamless=False), "TTTV."),
        ("pattern (0mg5,1mg6-)|(0mg6,1m7-) ungapped", 135e6, 10000, 20, dict(ot_pattern="(0mg5,1mg6-)|(0mg6,1m7-)"), ".NGG"),
        ("gapped <=4 mm / <=1 gap (edit distance K=5)", 135e6, 4096, 20, dict(ot_mismatch=5, ot_gap=2), ".NGG"),
        ("gapped <=1 mm / <=1 gap (edit distance K=2)", 135e6, 4096, 20, dict(ot_mismatch=2, ot_gap=2), ".NGG"),
        ("pattern 0mg-10,1mg-11- with <=1 gap (config 3)", 135e6, 10000, 20, dict(ot_pattern="0mg-10,1mg-11-", ot_gap=2), ".NGG"),
        ("two gaps: <=2 mm / <=2 gaps (edit distance K=4, banded-DP verifier)", 135e6, 2048, 20, dict(ot_mismatch=3, ot_gap=3), ".NGG"),
        ("two gaps: <=4 mm / <=2 gaps (edit distance K=6, banded-DP verifier)", 135e6, 1024, 20, dict(ot_mismatch=5, ot_gap=3), ".NGG"),
        ("pattern 5mg1- with <=1 gap (edit distance K=5; bound reject + DFS verifier)", 135e6, 1024, 20, dict(ot_pattern="5mg1-", ot_gap=2), ".NGG"),
        ("Cas12a 23-nt gapped <=4 mm / <=1 gap + PAM proximity TTTV.", 135e6, 2048, 23, dict(ot_mismatch=5, ot_gap=2, ot_pamless=False), "TTTV."),
    ]
    only = sys.argv[1:]
    scale = float(os.environ.get("MG_BENCH_SCALE", "1"))          # genome size factor (profiling under ncu)
    for name, nb, n, L, kw, pam in cases:
        if only and not any(o in name for o in only):
            continue
        nb = int(nb * scale)
        codes = torch.randint(0, 4, (nb,), dtype=torch.uint8, device=dev, generator=gen)
        ascii_ = torch.take(lut, codes.to(torch.int64))
        del codes
        g = _lib.Genome.from_device_ascii(ascii_.data_ptr(), np.array([0, nb // 2, nb], dtype=np.int64), ["a", "b"], stream=stream)
        torch.cuda.synchronize()
        rng = np.random.Generator(np.random.PCG64(3))
        grnas = [bytes(b"ACGT"[i] for i in rng.integers(0, 4, size=L)).decode() for _ in range(n)]
        q = _lib.Queries(grnas, stream=stream)
        crit = make_criteria(grna_length=L, **kw)
        prog = PAM(pam, L).program()
        counts = torch.zeros(2 * n, dtype=torch.int32, device=dev)
        _lib.screen(g, q, crit, prog, counts.data_ptr(), stream=stream)     # warm-up
        torch.cuda.synchronize()
        counts.zero_()
        t0 = time.perf_counter()
        _lib.screen(g, q, crit, prog, counts.data_ptr(), stream=stream)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        c = counts.cpu().numpy()
        print(json.dumps({"case": name, "genome_bp": nb, "n_grna": n, "L": L, "seconds": dt,
                          "grna_gbp_per_s": n * nb / 1e9 / dt, "tcell_per_s": 2.0 * n * nb / dt / 1e12,
                          "prefilter_edits": crit.prefilter_edits, "grna_failed": int(((c[:n] + c[n:]) > 0).sum())}))
        sys.stdout.flush()
        q.free()
        g.free()
        del ascii_


if __name__ == "__main__":
    main()
