"""Seed-and-join screen (csrc/screen_seedjoin.cu: half-gRNA hash index + neighbourhood probes + exact inline check, used for large
gRNA sets under the one-gap no-pattern criterion) against the bit-sliced DP path + verifier (itself count-for-count equal to the
exhaustive C oracle, tests/test_gpu_screen.py) and against the C oracle directly: per-oriented-query ANCHOR counts, one to one.
MG_GAPPED=seedjoin / dp forces either path whatever the set size."""
import os
import random

import numpy as np
import pytest

from oracle import c_oracle, minorg_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L():
    from minorg_b200 import _lib
    assert _lib.device_count() > 0, "no CUDA device"
    _lib.init(0)
    return _lib


def _contig(rng, n, alphabet="ACGT"):
    return "".join(rng.choice(alphabet) for _ in range(n))


def _counts(L, genome, q, crit, mode, wb=0, we=-1):
    import torch
    os.environ["MG_GAPPED"] = mode
    try:
        c = torch.zeros(2 * q.n, dtype=torch.int32, device="cuda")
        L.screen(genome, q, crit, None, c.data_ptr(), win_begin=wb, win_end=we, stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        return c.cpu().numpy().astype(np.int64)
    finally:
        os.environ.pop("MG_GAPPED", None)


def _world(seed, n_grna, sizes, n_runs=True):
    rng = random.Random(seed)
    contigs = []
    for i, n in enumerate(sizes):
        s = list(_contig(rng, n, "ACGT" * 10 + "acgt"))
        if n_runs and n > 5000:
            for _ in range(3):                                   # N runs and stray IUPAC codes: dirty windows for the DP path
                a = rng.randrange(100, n - 200)
                for j in range(a, a + rng.choice([1, 7, 60])):
                    s[j] = "N"
            s[rng.randrange(n)] = "R"
        contigs.append(("c%d" % i, "".join(s)))
    grnas = []
    big = [c for c in contigs if len(c[1]) > 100]
    for i in range(n_grna):
        _, seq = big[i % len(big)]
        p = rng.randrange(0, len(seq) - 24)
        g = list(seq[p:p + 20].upper().replace("N", "T").replace("R", "A"))
        for _ in range(rng.choice([0, 1, 2, 3, 4, 5, 6])):
            g[rng.randrange(20)] = rng.choice("ACGT")
        r = rng.random()
        if r < 0.3:                                              # the text has an extra base (deletion in the gRNA)
            k = rng.randrange(1, 19)
            g = g[:k] + g[k + 1:] + [rng.choice("ACGT")]
        elif r < 0.6:                                            # the gRNA has an extra base (insertion)
            k = rng.randrange(1, 19)
            g = g[:k] + [rng.choice("ACGT")] + g[k:19]
        elif r < 0.65:
            g = [rng.choice("ACGT") for _ in range(20)]
        g = "".join(g)
        grnas.append(O.reverse_complement(g) if i % 3 == 1 else g)
    return contigs, grnas


@pytest.mark.parametrize("M", [4, 3, 2, 1, 0])
def test_seedjoin_equals_dp_path_medium(L, M):
    from minorg_b200.ot_regex import make_criteria
    contigs, grnas = _world(100 + M, 3000, [400000, 150000, 37, 0, 90000, 3000])
    names = [c[0] for c in contigs]
    # (two long overlapping intervals: their interiors are screened by neither path, only their edges are delegated)
    masks = [(0, 1000, 3000), (0, 2500, 2600), (1, 0, 500), (4, 89000, 90000), (4, 40000, 40010), (0, 200000, 230000), (0, 225000, 260000),
             (1, 100000, 120000)]
    genome = L.Genome.from_contigs(contigs)
    genome.set_masks(masks)
    q = L.Queries(grnas)
    crit = make_criteria(ot_mismatch=M + 1, ot_gap=2, grna_length=20)
    try:
        want = _counts(L, genome, q, crit, "dp")
        got = _counts(L, genome, q, crit, "seedjoin")
        bad = np.nonzero(got != want)[0]
        assert bad.size == 0, (M, bad[:10], got[bad[:10]], want[bad[:10]])
        assert want.sum() > 100, M
        # window shards add up (anchors are owned by the window they end at; dirty ranges are cut with the shard)
        glen = genome.global_length
        cuts = [0, glen // 3 + 5, glen // 3 + 6, 2 * glen // 3, glen]
        parts = sum(_counts(L, genome, q, crit, "seedjoin", a, b) for a, b in zip(cuts[:-1], cuts[1:]))
        assert np.array_equal(parts, want), M
    finally:
        q.free()
        genome.free()
    assert names


def test_seedjoin_equals_c_oracle_small(L):
    """Directly against the exhaustive C oracle (2 x 40 kb, 400 gRNA), masks included."""
    from minorg_b200.ot_regex import make_criteria
    contigs, grnas = _world(7, 400, [40000, 41000, 300])
    masks = [(0, 100, 900), (1, 20000, 20500)]
    genome = L.Genome.from_contigs(contigs)
    genome.set_masks(masks)
    q = L.Queries(grnas)
    try:
        for M in (4, 2):
            crit = make_criteria(ot_mismatch=M + 1, ot_gap=2, grna_length=20)
            got = _counts(L, genome, q, crit, "seedjoin")
            want = c_oracle.screen_nopattern(contigs, masks, grnas, M, 1, threads=16).astype(np.int64)
            bad = np.nonzero(got != want)[0]
            assert bad.size == 0, (M, bad[:10], got[bad[:10]], want[bad[:10]])
            assert want.sum() > 50
    finally:
        q.free()
        genome.free()


def test_seedjoin_chunks_and_non_acgt_grna(L):
    """Chunk + halo genomes with the seed-and-join path; a gRNA with a non-ACGT base sends the whole set to the DP path."""
    from minorg_b200.ot_regex import make_criteria
    contigs, grnas = _world(9, 800, [120000, 60000])
    names = [c[0] for c in contigs]
    blob = np.frombuffer("".join(s for _, s in contigs).encode(), dtype=np.uint8)
    offsets = np.array([0, 120000, 180000], dtype=np.int64)
    whole = L.Genome.from_contigs(contigs)
    q = L.Queries(grnas)
    crit = make_criteria(ot_mismatch=5, ot_gap=2, grna_length=20)
    try:
        want = _counts(L, whole, q, crit, "dp")
        glen = whole.global_length
        got = np.zeros_like(want)
        for a, b in ((0, 50001), (50001, 130000), (130000, glen)):
            g = L.Genome.from_host_array_chunk(blob, offsets, names, a, b, halo=256)
            got += _counts(L, g, q, crit, "seedjoin")
            g.free()
        assert np.array_equal(got, want)
        q2 = L.Queries(grnas[:50] + ["ACGTNACGTACGTACGTACG"])
        assert np.array_equal(_counts(L, whole, q2, crit, "seedjoin"), _counts(L, whole, q2, crit, "dp"))
        q2.free()
    finally:
        q.free()
        whole.free()


def test_seedjoin_default_selection_stats_and_degenerate_sets(L):
    """Without MG_GAPPED the path is chosen by the set size (MG_SEEDJOIN_MIN); the kernel's own counters say which ran.
    A set in which thousands of gRNA share a 10-base half overflows a bucket's count field: the index is unusable and the
    DP path takes the whole set.  Duplicate gRNA are separate queries with equal counts."""
    import torch
    from minorg_b200.ot_regex import make_criteria
    contigs, grnas = _world(21, 1500, [60000, 30000], n_runs=False)
    genome = L.Genome.from_contigs(contigs)
    crit = make_criteria(ot_mismatch=5, ot_gap=2, grna_length=20)
    stream = torch.cuda.current_stream().cuda_stream

    def plain(q):
        c = torch.zeros(2 * q.n, dtype=torch.int32, device="cuda")
        L.screen(genome, q, crit, None, c.data_ptr(), stream=stream)
        torch.cuda.synchronize()
        return c.cpu().numpy().astype(np.int64)

    os.environ.pop("MG_GAPPED", None)
    q = L.Queries(grnas + grnas[:7])
    try:
        want = _counts(L, genome, q, crit, "dp")
        os.environ["MG_SEEDJOIN_MIN"] = "1000"
        _counts(L, genome, q, crit, "dp")                     # resets nothing: the counters belong to the seed-and-join launches
        got = plain(q)
        cand, surv = L.last_seedjoin_stats()
        assert np.array_equal(got, want)
        assert cand > 0 and 0 < surv < cand
        assert np.array_equal(got[:7], got[1500:1507]) and np.array_equal(got[q.n:q.n + 7], got[q.n + 1500:q.n + 1507])
        os.environ["MG_SEEDJOIN_MIN"] = "100000"              # above the set size: the DP path, same counts
        assert np.array_equal(plain(q), want)
    finally:
        os.environ.pop("MG_SEEDJOIN_MIN", None)
        q.free()
    # 5 000 gRNA with the same first half: bucket of table 0 holds 5 000 > 4 095 entries
    rng = random.Random(5)
    half = contigs[0][1][1000:1010].upper()
    shared = [half + _contig(rng, 10) for _ in range(5000)]
    q = L.Queries(shared)
    try:
        assert np.array_equal(_counts(L, genome, q, crit, "seedjoin"), _counts(L, genome, q, crit, "dp"))
    finally:
        q.free()
        genome.free()


@pytest.mark.parametrize("M", [5, 4, 3, 2, 1, 0])
def test_seedjoin_ungapped_equals_pair_word_scan_and_oracle(L, M):
    """ot_gap = 0 (Hamming <= M): routes R1 / R2 only, t = M // 2; against the pair-word scan (masks, N runs, contig ends, window
    shards, chunk + halo genomes) and, for a subset of the gRNA, against the C oracle."""
    from minorg_b200.ot_regex import make_criteria
    contigs, grnas = _world(300 + M, 2500, [300000, 120000, 25, 0, 50000])
    names = [c[0] for c in contigs]
    masks = [(0, 1000, 3000), (0, 2500, 2600), (1, 0, 500), (4, 49000, 50000), (4, 20000, 20010)]
    genome = L.Genome.from_contigs(contigs)
    genome.set_masks(masks)
    q = L.Queries(grnas)
    crit = make_criteria(ot_mismatch=M + 1, ot_gap=0, grna_length=20)
    try:
        want = _counts(L, genome, q, crit, "dp")
        got = _counts(L, genome, q, crit, "seedjoin")
        bad = np.nonzero(got != want)[0]
        assert bad.size == 0, (M, bad[:10], got[bad[:10]], want[bad[:10]])
        assert L.last_seedjoin_stats()[0] > 0
        assert want.sum() > (20 if M else 0), M
        glen = genome.global_length
        cuts = [0, glen // 3 + 5, glen // 3 + 6, 2 * glen // 3, glen]
        parts = sum(_counts(L, genome, q, crit, "seedjoin", a, b) for a, b in zip(cuts[:-1], cuts[1:]))
        assert np.array_equal(parts, want), M
        if M in (4, 1):
            sub = grnas[:300]
            q2 = L.Queries(sub)
            ref = c_oracle.screen_hamming(contigs, masks, sub, M, threads=16).astype(np.int64)
            c2 = _counts(L, genome, q2, crit, "seedjoin")
            assert np.array_equal(c2[:len(sub)] + c2[len(sub):], ref), M      # the Hamming oracle reports both orientations together
            q2.free()
    finally:
        q.free()
        genome.free()
    if M == 4:
        blob = np.frombuffer("".join(s for _, s in contigs).encode(), dtype=np.uint8)
        offsets = np.concatenate([[0], np.cumsum([len(s) for _, s in contigs])]).astype(np.int64)
        q = L.Queries(grnas)
        got = np.zeros_like(want)
        for a, b in ((0, 200001), (200001, 300030), (300030, glen)):
            g = L.Genome.from_host_array_chunk(blob, offsets, names, a, b, halo=256)
            got += _counts(L, g, q, crit, "seedjoin")
            g.free()
        q.free()
        # (chunk genomes carry no masks here: compare with the unmasked whole genome)
        whole = L.Genome.from_contigs(contigs)
        q = L.Queries(grnas)
        assert np.array_equal(got, _counts(L, whole, q, crit, "dp"))
        q.free()
        whole.free()


def test_seedjoin_ragged_subjects_fall_back_or_delegate(L):
    """Subjects that are mostly N runs, cut into thousands of short contigs, or carry thousands of masked intervals: the
    delegated window ranges exceed their share / count limits and the exhaustive kernels take the whole call (the seed-and-join
    counters read zero), or - below the limits - the two sides split the windows.  Counts are the DP path's."""
    import torch
    from minorg_b200.ot_regex import make_criteria
    rng = random.Random(31)
    crit = make_criteria(ot_mismatch=5, ot_gap=2, grna_length=20)
    base = _contig(rng, 120000)
    grnas = []
    for i in range(600):
        p = rng.randrange(0, len(base) - 30)
        g = list(base[p:p + 20])
        for _ in range(rng.choice([0, 1, 2, 3, 4])):
            g[rng.randrange(20)] = rng.choice("ACGT")
        grnas.append("".join(g))
    q = L.Queries(grnas)
    try:
        # (1) an N every ~40 bases: every window is near an invalid block
        s = list(base)
        for j in range(0, len(s), 41):
            s[j] = "N"
        # (2) 3 000 contigs of 40 bases; (3) 6 000 masked intervals of 5 bases, 20 apart
        cases = [([("a", "".join(s))], []),
                 ([("c%d" % i, base[40 * i:40 * i + 40]) for i in range(3000)], []),
                 ([("a", base)], [(0, 20 * i, 20 * i + 5) for i in range(6000)]),
                 ([("a", base)], [(0, 3000 * i, 3000 * i + 400) for i in range(40)])]
        for k, (contigs, masks) in enumerate(cases):
            genome = L.Genome.from_contigs(contigs)
            genome.set_masks(masks)
            want = _counts(L, genome, q, crit, "dp")
            got = _counts(L, genome, q, crit, "seedjoin")
            assert np.array_equal(got, want), k
            if k < 3:
                assert L.last_seedjoin_stats() == (0, 0), k        # the scan kernel did not run
            else:
                assert L.last_seedjoin_stats()[0] > 0
            genome.free()
        assert want.sum() > 100
    finally:
        q.free()
