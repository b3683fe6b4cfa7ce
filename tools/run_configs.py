#!/usr/bin/env python
"""Runs the synthetic configurations BASELINE.json names (configs[1..4]) on ONE GPU, full size unless --scale says
otherwise, and prints one JSON line each (kept in profiles/).  Device-timed with genomes resident, like bench.py's `value`."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from minorg_b200 import _lib  # noqa: E402
from minorg_b200.ot_regex import make_criteria  # noqa: E402
from minorg_b200.pam import PAM  # noqa: E402


def synth(dev, seed, n_bases):
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
    out = torch.empty(n_bases, dtype=torch.uint8, device=dev)
    step = 1 << 28
    for a in range(0, n_bases, step):                       # chunked: the int64 index temp stays small
        b = min(n_bases, a + step)
        out[a:b] = torch.take(lut, torch.randint(0, 4, (b - a,), dtype=torch.int64, device=dev, generator=gen))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0, help="fraction of the named gRNA count to run")
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    _lib.init(0)
    dev = torch.device("cuda", 0)
    stream = torch.cuda.current_stream().cuda_stream
    rng = np.random.Generator(np.random.PCG64(1000))
    configs = [
        ("configs[1]: 8x135 Mb, 10k SpCas9 20-nt gRNA, ungapped <=4 mm", [(g, 135_000_000, 5) for g in range(8)], 10_000, 20, dict(ot_mismatch=5), ".NGG"),
        ("configs[2]: 8x135 Mb, 100k gRNA, <=1 gap + --ot-pattern 0mg-10,1mg-11-", [(g, 135_000_000, 5) for g in range(8)], 100_000, 20, dict(ot_gap=2, ot_pattern="0mg-10,1mg-11-"), ".NGG"),
        ("configs[3]: 2x390 Mb + 50 Mb background, Cas12a TTTV. 23-nt, 100k gRNA, <=4 mm", [(20, 390_000_000, 12), (21, 390_000_000, 12), (22, 50_000_000, 4)], 100_000, 23, dict(ot_mismatch=5), "TTTV."),
        ("configs[3] with PAM proximity (ot_pamless=False)", [(20, 390_000_000, 12), (21, 390_000_000, 12), (22, 50_000_000, 4)], 100_000, 23, dict(ot_mismatch=5, ot_pamless=False), "TTTV."),
        ("configs[4]: 3.1 Gb human-scale genome (24 contigs), 1M gRNA, <=4 mm both strands", [(30, 3_100_000_000, 24)], 1_000_000, 20, dict(ot_mismatch=5), ".NGG"),
        ("north-star target: 8x135 Mb, 1M gRNA, <=4 mm / <=1 gap (run at 1/10 of --scale)", [(g, 135_000_000, 5) for g in range(8)], 100_000, 20, dict(ot_mismatch=5, ot_gap=2), ".NGG"),
    ]
    for name, genomes, n_grna, L, kw, pam in configs:
        if args.only and args.only not in name:
            continue
        n = max(32, int(n_grna * args.scale))
        parts, offsets, pos = [], [0], 0
        for seed, nb, n_ctg in genomes:
            parts.append(synth(dev, seed, nb))
            for c in range(n_ctg):
                pos += nb // n_ctg if c < n_ctg - 1 else nb - (nb // n_ctg) * (n_ctg - 1)
                offsets.append(pos)
        ascii_ = torch.cat(parts) if len(parts) > 1 else parts[0]
        del parts
        total = int(ascii_.numel())
        genome = _lib.Genome.from_device_ascii(ascii_.data_ptr(), np.array(offsets, dtype=np.int64), ["c%d" % i for i in range(len(offsets) - 1)], stream=stream)
        torch.cuda.synchronize()
        del ascii_
        torch.cuda.empty_cache()
        grnas = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=(n, L))]
        grnas = [bytes(r).decode() for r in grnas]
        q = _lib.Queries(grnas, stream=stream)
        crit = make_criteria(grna_length=L, **kw)
        prog = PAM(pam, L).program()
        counts = torch.zeros(2 * n, dtype=torch.int32, device=dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.screen(genome, q, crit, prog, counts.data_ptr(), stream=stream)
        e1.record()
        torch.cuda.synchronize()
        sec = e0.elapsed_time(e1) / 1e3
        c = counts.cpu().numpy()
        print(json.dumps({"config": name, "genome_bp": total, "n_grna": n, "n_grna_named": n_grna, "L": L, "seconds": sec,
                          "grna_gbp_per_s": n * total / 1e9 / sec, "tcell_per_s": 2.0 * n * total / sec / 1e12,
                          "seconds_at_named_size_8_gpus": sec * ((1_000_000 if name.startswith("north-star") else n_grna) / n) / 8,
                          "grna_failed": int(((c[:n] + c[n:]) > 0).sum()), "device_bytes": genome.device_bytes}))
        sys.stdout.flush()
        q.free()
        genome.free()


if __name__ == "__main__":
    main()
