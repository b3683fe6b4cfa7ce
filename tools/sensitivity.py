#!/usr/bin/env python
"""How much of the pair-word scan's speed depends on uniform random sequence?  

The stage-1 test of screen_pairs_kernel ends a (warp, 32 gRNA) step unless one of its 1024 cells can still be within
budget; on i.i.d. uniform sequence 0.95 % of the steps go on to the exact stage 2.  This runs the headline criterion
(10 000 SpCas9 20-nt gRNA, <= 4 mismatches, both strands, one 135 Mb genome) on
  uniform   the bench's genome (i.i.d. uniform ACGT), gRNA cut from planted random targets;
  at_rich   a 36 %-GC genome, 5 % of it tandem and dispersed repeats, gRNA enumerated (NGG) from a planted 10-member gene
            family at ~90 % identity to a common ancestor (MINORg's use case: homologous genes), family members masked;
  tiled     the bundled config-1 Arabidopsis subset (3 FASTA files, 1.01 Mb) tiled to 135 Mb, gRNA enumerated from its
            own target genes - every hit occurs ~134 times: the worst case for the early-out and the hit path;
and prints second_stage_fraction and Tcell/s for each.  The floor of the kernel whatever the data is the ceiling without
any early-out: 20 shared loads per step (bench.py roofline.ceilings_tcell_s.lsu_without_early_out).

    python tools/sensitivity.py > profiles/r02_sensitivity_n1.jsonl"""
import gzip
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from minorg_b200 import _lib  # noqa: E402
from minorg_b200.fasta import read_fasta, reverse_complement  # noqa: E402
from minorg_b200.ot_regex import make_criteria  # noqa: E402
from minorg_b200.pam import PAM  # noqa: E402

GENOME_BP, CONTIGS, N_GRNA, L = 135_000_000, 5, 10_000, 20
LUT = np.frombuffer(b"ACGT", dtype=np.uint8)


def enumerate_grnas(seqs, n):
    prog = PAM(".NGG", L).program()
    seen, out = set(), []
    for seq in seqs:
        plus, minus = _lib.enumerate_sites(seq, prog, L)
        rc = reverse_complement(seq)
        for text, starts in ((seq, plus), (rc, minus)):
            for s in starts.tolist():
                k = text[s:s + L].upper()
                if k not in seen and set(k) <= set("ACGT"):
                    seen.add(k)
                    out.append(k)
    return out[:n]


def mutate(rng, codes, rate):
    out = codes.copy()
    hit = rng.random(len(out)) < rate
    out[hit] = (out[hit] + rng.integers(1, 4, size=int(hit.sum()))) % 4
    return out


def genome_uniform(rng):
    g = rng.integers(0, 4, size=GENOME_BP, dtype=np.uint8)
    targets = [rng.integers(0, 4, size=2000, dtype=np.uint8) for _ in range(40)]
    return g, targets


def genome_at_rich(rng):
    g = rng.choice(4, size=GENOME_BP, p=[0.32, 0.18, 0.18, 0.32]).astype(np.uint8)
    budget, covered = int(0.05 * GENOME_BP), 0
    library = [rng.choice(4, size=int(rng.integers(300, 3000)), p=[0.32, 0.18, 0.18, 0.32]).astype(np.uint8) for _ in range(50)]
    while covered < budget:
        at = int(rng.integers(0, GENOME_BP - 6000))
        if rng.random() < 0.5:                                  # tandem repeat: unit of 2..50 bp over 200..5000 bp
            unit = rng.choice(4, size=int(rng.integers(2, 51)), p=[0.32, 0.18, 0.18, 0.32]).astype(np.uint8)
            span = int(rng.integers(200, 5000))
            rep = mutate(rng, np.tile(unit, span // len(unit) + 1)[:span], 0.02)
        else:                                                   # dispersed repeat: a library element, 5..15 % diverged
            rep = mutate(rng, library[int(rng.integers(0, len(library)))], float(rng.uniform(0.05, 0.15)))
        g[at:at + len(rep)] = rep
        covered += len(rep)
    ancestor = rng.choice(4, size=3000, p=[0.27, 0.23, 0.23, 0.27]).astype(np.uint8)
    family = [mutate(rng, ancestor, 0.10) for _ in range(10)]
    return g, family


def genome_tiled():
    cfg = os.path.join(ROOT, "tests", "golden", "config1")
    parts = []
    for f in ("subset_ref_TAIR10.fasta.gz", "subset_9654.fasta.gz", "subset_9655.fasta.gz"):
        for _, seq in read_fasta(os.path.join(cfg, f)):
            parts.append(np.frombuffer(seq.encode("latin-1"), dtype=np.uint8))
    unit = np.concatenate(parts)
    reps = GENOME_BP // len(unit) + 1
    return np.tile(unit, reps)[:GENOME_BP], [s for _, s in read_fasta(os.path.join(cfg, "targets.fasta"))], len(unit)


def run(label, ascii_host, grnas, mask_seqs, note):
    dev = torch.device("cuda", 0)
    stream = torch.cuda.current_stream().cuda_stream
    clen = GENOME_BP // CONTIGS
    offsets = np.arange(CONTIGS + 1, dtype=np.int64) * clen
    d = torch.from_numpy(np.ascontiguousarray(ascii_host[:clen * CONTIGS])).to(dev)
    genome = _lib.Genome.from_device_ascii(d.data_ptr(), offsets, ["c%d" % i for i in range(CONTIGS)], stream=stream)
    torch.cuda.synchronize()
    del d
    masks = sorted({(c, s, e) for _, c, s, e, _ in genome.find_exact(mask_seqs, stream=stream)}) if mask_seqs else []
    genome.set_masks(masks, stream=stream)
    q = _lib.Queries(grnas, stream=stream)
    n = len(grnas)
    crit = make_criteria(ot_mismatch=5, ot_gap=0, grna_length=L)
    counts = torch.zeros(2 * n, dtype=torch.int32, device=dev)
    _lib.screen(genome, q, crit, None, counts.data_ptr(), stream=stream)          # warm-up
    torch.cuda.synchronize()
    ms = []
    for _ in range(3):
        counts.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.screen(genome, q, crit, None, counts.data_ptr(), stream=stream)
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    steps, stage2 = _lib.last_screen_stats()
    c = counts.cpu().numpy().astype(np.int64)
    sec = float(np.median(ms)) / 1e3
    print(json.dumps({"genome": label, "note": note, "genome_bp": clen * CONTIGS, "n_grna": n, "masks": len(masks), "ms": sec * 1e3,
                      "tcell_per_s": 2.0 * n * clen * CONTIGS / sec / 1e12, "second_stage_fraction": stage2 / max(1, steps),
                      "hits_total": int(c.sum()), "grna_failed": int(((c[:n] + c[n:]) > 0).sum())}))
    sys.stdout.flush()
    q.free()
    genome.free()


def run_seedjoin(label, ascii_host, grnas, mask_seqs, note):
    """The north-star criterion (<= 4 mismatches / <= 1 gap) through the seed-and-join screen on the same kinds of genome, with
    a gRNA set large enough for it (skewed sets fill some index buckets much more than a random set does); the DP path on 1/64
    of the windows for comparison, per-gRNA counts compared on that range."""
    dev = torch.device("cuda", 0)
    stream = torch.cuda.current_stream().cuda_stream
    clen = GENOME_BP // CONTIGS
    offsets = np.arange(CONTIGS + 1, dtype=np.int64) * clen
    d = torch.from_numpy(np.ascontiguousarray(ascii_host[:clen * CONTIGS])).to(dev)
    genome = _lib.Genome.from_device_ascii(d.data_ptr(), offsets, ["c%d" % i for i in range(CONTIGS)], stream=stream)
    torch.cuda.synchronize()
    del d
    masks = sorted({(c, s, e) for _, c, s, e, _ in genome.find_exact(mask_seqs, stream=stream)}) if mask_seqs else []
    genome.set_masks(masks, stream=stream)
    q = _lib.Queries(grnas, stream=stream)
    n = len(grnas)
    crit = make_criteria(ot_mismatch=5, ot_gap=2, grna_length=L)
    counts = torch.zeros(2 * n, dtype=torch.int32, device=dev)
    glen = genome.global_length

    def timed(mode, wb=0, we=-1):
        os.environ["MG_SEEDJOIN"] = mode
        counts.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.screen(genome, q, crit, None, counts.data_ptr(), win_begin=wb, win_end=we, stream=stream)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 1e3, counts.clone()
    timed("on", 0, 1 << 20)                                        # index build, warm-up
    sec, c_all = timed("on")
    cand, surv = _lib.last_seedjoin_stats()
    a, b = glen // 2, glen // 2 + glen // 64
    t_sj, c_sj = timed("on", a, b)
    t_dp, c_dp = timed("off", a, b)
    os.environ.pop("MG_SEEDJOIN", None)
    c = c_all.cpu().numpy().astype(np.int64)
    print(json.dumps({"genome": label, "criterion": "<=4 mismatches / <=1 gap, seed-and-join screen", "note": note, "genome_bp": clen * CONTIGS,
                      "n_grna": n, "masks": len(masks), "seconds": sec, "tcell_per_s": 2.0 * n * clen * CONTIGS / sec / 1e12,
                      "seedjoin_path_ran": cand > 0, "candidates_per_cell": cand / (2.0 * n * clen * CONTIGS), "stage2_fraction": surv / max(1, cand),
                      "anchors_total": int(c.sum()), "grna_failed": int(((c[:n] + c[n:]) > 0).sum()),
                      "range_windows": b - a, "seedjoin_seconds_on_range": t_sj, "dp_seconds_on_range": t_dp,
                      "counts_equal_on_range": bool(torch.equal(c_sj, c_dp))}))
    sys.stdout.flush()
    q.free()
    genome.free()


def top_up(rng, grnas, n, p=None):
    """Fill a gRNA list to n with point variants of its members (half of the rest) and random 20-mers of base composition p."""
    seen, out = set(grnas), list(grnas)
    base_n = len(out)
    while len(out) < n:
        if base_n and len(out) < base_n + (n - base_n) // 2:
            k = list(out[int(rng.integers(0, base_n))])
            for _ in range(int(rng.integers(1, 4))):
                k[int(rng.integers(0, L))] = "ACGT"[int(rng.integers(0, 4))]
            k = "".join(k)
        else:
            k = LUT[rng.choice(4, size=L, p=p)].tobytes().decode()
        if k not in seen:
            seen.add(k)
            out.append(k)
    return out


def main_seedjoin(n_grna):
    _lib.init(0)
    rng = np.random.Generator(np.random.PCG64(43))
    g, targets = genome_uniform(rng)
    seqs = [LUT[t].tobytes().decode() for t in targets]
    for i, t in enumerate(targets):
        g[1_000_000 * (i + 1):1_000_000 * (i + 1) + len(t)] = t
    run_seedjoin("uniform", LUT[g], top_up(rng, enumerate_grnas(seqs, n_grna), n_grna), seqs,
                 "i.i.d. uniform ACGT; gRNA from 40 planted random 2-kb targets (masked), their point variants and random 20-mers")
    g, family = genome_at_rich(rng)
    seqs = [LUT[t].tobytes().decode() for t in family]
    for i, t in enumerate(family):
        g[10_000_000 * (i + 1):10_000_000 * (i + 1) + len(t)] = t
    run_seedjoin("at_rich_repeats_gene_family", LUT[g], top_up(rng, enumerate_grnas(seqs, n_grna), n_grna, p=[0.32, 0.18, 0.18, 0.32]), seqs,
                 "36 % GC, 5 % tandem + dispersed repeats; gRNA from a planted 10-member gene family at ~90 % identity (members masked), 1-3-base "
                 "variants of them (half of the set: thousands of gRNA share their halves) and random 20-mers of the genome's composition")
    g, tseqs, unit = genome_tiled()
    parts = [bytes(g[i:i + 200_000]).decode("latin-1") for i in range(0, unit, 200_000)]
    run_seedjoin("config1_tiled", g, top_up(rng, enumerate_grnas(tseqs + parts, n_grna), n_grna), tseqs,
                 "bundled config-1 genomes (%d bp) tiled %.0fx to 135 Mb; gRNA = NGG sites of the target genes and of the genomes themselves "
                 "(every hit repeats ~134x), variants, random" % (unit, GENOME_BP / unit))


def main():
    if "--seedjoin" in sys.argv:
        return main_seedjoin(int(sys.argv[sys.argv.index("--seedjoin") + 1]) if len(sys.argv) > sys.argv.index("--seedjoin") + 1 else 200_000)
    _lib.init(0)
    rng = np.random.Generator(np.random.PCG64(42))
    g, targets = genome_uniform(rng)
    seqs = [LUT[t].tobytes().decode() for t in targets]
    for i, t in enumerate(targets):
        g[1_000_000 * (i + 1):1_000_000 * (i + 1) + len(t)] = t
    run("uniform", LUT[g], enumerate_grnas(seqs, N_GRNA), seqs, "i.i.d. uniform ACGT; gRNA from 40 planted random 2-kb targets (masked)")
    g, family = genome_at_rich(rng)
    seqs = [LUT[t].tobytes().decode() for t in family]
    for i, t in enumerate(family):
        g[10_000_000 * (i + 1):10_000_000 * (i + 1) + len(t)] = t
    grnas = enumerate_grnas(seqs, N_GRNA)
    while len(grnas) < N_GRNA:                       # a 10 x 3 kb family yields ~4 000 distinct NGG 20-mers: top up with 1-2 mismatch variants
        base = grnas[int(rng.integers(0, len(grnas)))]
        k = list(base)
        k[int(rng.integers(0, L))] = "ACGT"[int(rng.integers(0, 4))]
        grnas.append("".join(k))
    run("at_rich_repeats_gene_family", LUT[g], grnas[:N_GRNA], seqs,
        "36 % GC, 5 % tandem + dispersed repeats; gRNA from a planted 10-member gene family at ~90 % identity (members masked), topped up with point variants")
    g, tseqs, unit = genome_tiled()
    grnas = enumerate_grnas(tseqs, N_GRNA)
    run("config1_tiled", g, grnas, tseqs, "bundled config-1 genomes (%d bp) tiled %.0fx to 135 Mb; %d gRNA from their target genes (every hit repeats)" % (unit, GENOME_BP / unit, len(grnas)))


if __name__ == "__main__":
    main()
