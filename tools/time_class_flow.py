#!/usr/bin/env python
"""User-level end-to-end timing of the drop-in class on FILES: synthetic genomes written as FASTA, targets as FASTA,
MINORg.grna() + MINORg.filter_background() (mask finding, native FASTA load + pack, screen, writers).
python tools/time_class_flow.py [n_genomes] [genome_mb] [n_targets] [target_bp] [ot_gap]
(1 135 64 125000 2: ~10^6 gRNA from 8 Mb of targets against one 135 Mb genome under the north star's criterion)"""
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402


def write_fasta(path, seqs, width=60):
    with open(path, "wb") as f:
        for name, arr in seqs:
            f.write(b">" + name.encode() + b"\n")
            n = arr.size // width * width
            body = np.concatenate([arr[:n].reshape(-1, width), np.full((n // width, 1), 10, dtype=np.uint8)], axis=1)
            f.write(body.tobytes())
            if n < arr.size:
                f.write(arr[n:].tobytes() + b"\n")


def main():
    n_genomes = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    genome_mb = float(sys.argv[2]) if len(sys.argv) > 2 else 135.0
    n_targets = int(sys.argv[3]) if len(sys.argv) > 3 else 64
    tlen = int(sys.argv[4]) if len(sys.argv) > 4 else 2000
    ot_gap = int(sys.argv[5]) if len(sys.argv) > 5 else 0
    from minorg_b200.MINORg import MINORg
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    with tempfile.TemporaryDirectory() as td:
        rng = np.random.default_rng(1)
        targets = [("t%03d" % i, lut[rng.integers(0, 4, tlen)]) for i in range(n_targets)]
        write_fasta(os.path.join(td, "targets.fasta"), targets)
        files = []
        for g in range(n_genomes):
            n = int(genome_mb * 1e6) // 5
            contigs = []
            for c in range(5):
                arr = lut[rng.integers(0, 4, n)]
                for k in range(c, n_targets, 5 * n_genomes):          # plant some targets (masked on-targets)
                    p = 1000 + (k * 7919 * 1000) % (n - tlen - 3000)
                    arr[p:p + tlen] = targets[(k + g) % n_targets][1]
                contigs.append(("g%d_chr%d" % (g, c), arr))
            path = os.path.join(td, "genome%d.fasta" % g)
            write_fasta(path, contigs)
            files.append(path)
        t0 = time.perf_counter()
        m = MINORg(directory=os.path.join(td, "out"), prefix="flow")
        m.target = os.path.join(td, "targets.fasta")
        m.ot_mismatch = 5
        m.ot_gap = ot_gap
        for i, f in enumerate(files):
            m.add_background(f, alias="bg%d" % i)
        m.screen_reference = False
        t1 = time.perf_counter()
        m.grna()
        t2 = time.perf_counter()
        m.filter_background()
        t3 = time.perf_counter()
        n = len(m.grna_hits.seqs)
        npass = sum(1 for s in m.grna_hits.seqs if m.grna_hits.get_gRNAseq_by_seq(s).check("background"))
        total_bp = n_genomes * int(genome_mb * 1e6)
        print({"genomes": n_genomes, "genome_mb": genome_mb, "targets": n_targets, "target_bp": tlen, "ot_gap": ot_gap, "n_grna": n, "pass": npass,
               "init_s": round(t1 - t0, 3), "grna_s": round(t2 - t1, 3), "filter_background_s": round(t3 - t2, 3),
               "grna_gbp_per_s_files_to_flags": round(n * total_bp / 1e9 / (t3 - t2), 1)})
        m.close()


if __name__ == "__main__":
    main()
