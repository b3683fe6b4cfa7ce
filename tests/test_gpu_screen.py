"""GPU parity tests (run on the B200 box: pytest -m gpu).  Every call goes through the C ABI of
libminorg_b200.so (ctypes); results are compared with the CPU oracle on the same inputs."""
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
    name, sms = _lib.init(0)
    print("device:", name, sms, "SMs")
    return _lib


def _random_contig(rng, n, alphabet="ACGT"):
    return "".join(rng.choice(alphabet) for _ in range(n))


def test_pack_roundtrip(L):
    rng = random.Random(1)
    contigs = [("a", _random_contig(rng, 1000, "ACGTacgtNRY")), ("b", _random_contig(rng, 31)), ("c", ""),
               ("d", _random_contig(rng, 5)), ("e", _random_contig(rng, 4097, "ACGTn"))]
    g = L.Genome.from_contigs(contigs)
    assert g.n_contigs == 5 and g.total_bases == sum(len(s) for _, s in contigs)
    for i, (_, s) in enumerate(contigs):
        assert g.contig_length(i) == len(s)
        want = "".join(c.upper() if c in "ACGTacgt" else "N" for c in s)
        assert g.decode(i, 0, len(s)) == want
    g.free()


def _hamming_via_abi(L, contigs, masks, grnas, M):
    from minorg_b200.ot_regex import make_criteria
    blob = np.frombuffer("".join(s for _, s in contigs).encode() or b"\0", dtype=np.uint8)
    offsets = np.zeros(len(contigs) + 1, dtype=np.int64)
    np.cumsum([len(s) for _, s in contigs], out=offsets[1:])
    gb = np.frombuffer("".join(grnas).encode(), dtype=np.uint8)
    crit = make_criteria(ot_mismatch=M + 1, ot_gap=0)
    return L.screen_host(blob, offsets, masks, gb, len(grnas), len(grnas[0]), crit, None)


def test_hamming_golden_cases(L, golden):
    n = 0
    for c in golden("screen_cases.json")["cases"]:
        k = c["criteria"]
        if k.get("pattern") is not None or not k["pamless"] or k["ot_gap"] > 1:
            continue
        M = max(1, k["ot_mismatch"]) - 1
        contigs = [tuple(x) for x in c["contigs"]]
        names = [x[0] for x in contigs]
        masks = [(names.index(m[1]), m[2], m[3]) for m in c["masks"]]
        got = _hamming_via_abi(L, contigs, masks, c["grnas"], M)
        want = O.screen_hamming(c["grnas"], contigs, c["masks"], M)
        assert got.tolist() == want, c["name"]
        assert [x > 0 for x in got.tolist()] == c["fail"], c["name"]     # verdicts of the reference's predicates
        n += 1
    assert n >= 7


@pytest.mark.parametrize("glen,M,seed", [(20, 0, 1), (20, 1, 2), (20, 2, 3), (20, 4, 4), (23, 3, 5), (12, 1, 6),
                                         (32, 5, 7), (17, 6, 8), (1, 0, 9), (20, 7, 10)])
def test_hamming_random_vs_c_oracle(L, glen, M, seed):
    rng = random.Random(seed)
    contigs = [("chr%d" % i, _random_contig(rng, rng.choice([70000, 3000, 17, 40000]), "ACGT" * 8 + "Nacgt"))
               for i in range(4)]
    grnas = []
    for i in range(300):
        _, seq = contigs[rng.randrange(len(contigs))]
        if len(seq) > glen + 4 and i % 3:
            p = rng.randrange(0, len(seq) - glen)
            g = list(seq[p:p + glen].upper().replace("N", "A"))
            for _ in range(rng.randrange(0, M + 3)):
                g[rng.randrange(glen)] = rng.choice("ACGT")
            g = "".join(g)
            grnas.append(O.reverse_complement(g) if rng.random() < 0.5 else g)
        else:
            grnas.append(_random_contig(rng, glen, "ACGT" * 6 + "N"))
    masks = [(0, 100, 400), (0, 350, 900), (1, 0, 3000), (3, 39990, 40000), (0, 69990, 70000)]
    masks = [m for m in masks if m[2] <= len(contigs[m[0]][1])]
    if M >= glen:
        M = glen - 1
    got = _hamming_via_abi(L, contigs, masks, grnas, M)
    want = c_oracle.screen_hamming(contigs, masks, grnas, M, threads=8)
    assert np.array_equal(got, want)
    assert (got > 0).any() and (got == 0).any() or glen <= 12 or M >= 5


def test_resident_api_and_sharding(L):
    """mg_screen on resident handles; shards of the window range add up to the whole (multi-GPU split)."""
    import torch
    from minorg_b200.ot_regex import make_criteria
    rng = random.Random(11)
    contigs = [("x", _random_contig(rng, 50000)), ("y", _random_contig(rng, 20000))]
    grnas = [contigs[0][1][i * 97:i * 97 + 20] for i in range(64)] + [_random_contig(rng, 20) for _ in range(37)]
    g = L.Genome.from_contigs(contigs)
    g.set_masks([(0, 0, 2000)])
    q = L.Queries(grnas)
    crit = make_criteria(ot_mismatch=3, ot_gap=0)
    whole = torch.zeros(2 * len(grnas), dtype=torch.int32, device="cuda")
    L.screen(g, q, crit, None, whole.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
    parts = torch.zeros_like(whole)
    glen = g.global_length
    cuts = [0, glen // 3, glen // 3 + 7, 2 * glen // 3, glen]
    for a, b in zip(cuts[:-1], cuts[1:]):
        L.screen(g, q, crit, None, parts.data_ptr(), win_begin=a, win_end=b, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert torch.equal(whole, parts)
    n = len(grnas)
    got = (whole[:n] + whole[n:]).cpu().numpy().astype(np.uint32)
    want = c_oracle.screen_hamming(contigs, [(0, 0, 2000)], grnas, 2, threads=4)
    assert np.array_equal(got, want)
    q.free()
    g.free()


def test_many_queries_multiple_chunks(L):
    """More query groups than fit one shared-memory chunk (exercises the chunk loop)."""
    rng = random.Random(12)
    contigs = [("x", _random_contig(rng, 3000))]
    grnas = [_random_contig(rng, 20) for _ in range(13000)]
    for i in range(0, 13000, 7):
        p = rng.randrange(0, 2980)
        grnas[i] = contigs[0][1][p:p + 20]
    got = _hamming_via_abi(L, contigs, [], grnas, 1)
    want = c_oracle.screen_hamming(contigs, [], grnas, 1, threads=8)
    assert np.array_equal(got, want)


def test_enumeration_matches_reference_vectors(L, golden, tmp_path):
    from minorg_b200.generate_grna import find_multi_gRNA
    gdir = os.path.join(os.path.dirname(__file__), "golden")
    for r in golden("enumerate.json")["rows"]:
        h = find_multi_gRNA(os.path.join(gdir, r["fasta"]), pam=r["pam"], gRNA_len=r["L"])
        assert len(h) == r["n_seqs"], (r["fasta"], r["pam"], r["L"])
        assert len(h.flatten_hits()) == r["n_hits"]
        if "hits" in r:
            got = [[s, [[x.target_id, x.range[0], x.range[1], x.strand, x.hit_id, x.target_len] for x in h.get_hits(s)]]
                   for s in h.seqs]
            assert got == r["hits"], (r["fasta"], r["pam"], r["L"])
        if "fasta_text" in r:
            fa, mp = str(tmp_path / "g.fasta"), str(tmp_path / "g.map")
            h.write_fasta(fa, write_all=True)
            h.write_mapping(mp, write_all=True, write_checks=True)
            assert open(fa).read() == r["fasta_text"]
            assert open(mp).read() == r["map_text"]


def test_mask_find_vs_oracle(L):
    rng = random.Random(13)
    t1, t2, t3 = _random_contig(rng, 700), _random_contig(rng, 45), "ACGTACGT"
    contigs = [("c0", _random_contig(rng, 5000) + t1 + _random_contig(rng, 300) + O.reverse_complement(t1).lower()),
               ("c1", t2 + _random_contig(rng, 2000) + t2 + "N" + t1[:699]),
               ("c2", "ACGTACGTACGT" + t3)]
    seqs = [("t1", t1), ("t2", t2), ("t3", t3), ("absent", _random_contig(rng, 60))]
    g = L.Genome.from_contigs(contigs)
    got = g.find_exact([s for _, s in seqs])
    names = [c[0] for c in contigs]
    got_set = {(seqs[a][0], names[b], c, d) for a, b, c, d, _ in got}
    assert got_set == O.find_masks(seqs, contigs)
    assert len(got_set) >= 6
    g.free()


def test_microbench_runs(L):
    assert L.microbench(0) > 1000 and L.microbench(1) > 500


def _general_via_abi(L, contigs, masks, grnas, k, glen):
    from minorg_b200.ot_regex import make_criteria
    from minorg_b200.pam import PAM
    blob = np.frombuffer("".join(s for _, s in contigs).encode() or b"\0", dtype=np.uint8)
    offsets = np.zeros(len(contigs) + 1, dtype=np.int64)
    np.cumsum([len(s) for _, s in contigs], out=offsets[1:])
    gb = np.frombuffer("".join(grnas).encode(), dtype=np.uint8)
    crit = make_criteria(ot_mismatch=k["ot_mismatch"], ot_gap=k["ot_gap"], ot_pattern=k.get("pattern"),
                         ot_unaligned_as_mismatch=k.get("uam", True), ot_unaligned_as_gap=k.get("uag", False),
                         ot_pamless=k["pamless"], grna_length=glen)
    pam = PAM(k["pam"], glen).program()
    return L.screen_host(blob, offsets, masks, gb, len(grnas), glen, crit, pam)


def test_general_golden_cases(L, golden, capsys):
    """PAM proximity, patterns, gaps: device verdicts vs the verdicts of the reference's own predicates over
    the exhaustive alignment universe (tests/golden/screen_cases.json)."""
    bad = []
    n = 0
    for c in golden("screen_cases.json")["cases"]:
        k = c["criteria"]
        contigs = [tuple(x) for x in c["contigs"]]
        names = [x[0] for x in contigs]
        masks = [(names.index(m[1]), m[2], m[3]) for m in c["masks"]]
        got = _general_via_abi(L, contigs, masks, c["grnas"], k, c["L"])
        flags = [x > 0 for x in got.tolist()]
        if flags != c["fail"]:
            bad.append((c["name"], flags, c["fail"]))
        n += 1
    assert not bad, bad
    assert n >= 30


@pytest.mark.parametrize("seed", range(16))
def test_general_random_vs_python_oracle(L, seed):
    rng = random.Random(100 + seed)
    glen = rng.choice([10, 12])
    contigs = [("a", _random_contig(rng, rng.randrange(20, 60), "ACGT" * 5 + "N")), ("b", _random_contig(rng, rng.randrange(3, 30)))]
    grnas = []
    for i in range(6):
        _, seq = contigs[0]
        p = rng.randrange(0, len(seq) - glen)
        g = list(seq[p:p + glen].replace("N", "C"))
        for _ in range(rng.randrange(0, 3)):
            g[rng.randrange(glen)] = rng.choice("ACGT")
        if rng.random() < 0.4:
            del g[rng.randrange(1, glen - 1)]
            g.append(rng.choice("ACGT"))
        g = "".join(g)
        grnas.append(O.reverse_complement(g) if rng.random() < 0.5 else g)
    k = dict(ot_mismatch=rng.choice([1, 2, 3]), ot_gap=rng.choice([0, 1, 2, 2, 3]),
             pattern=rng.choice([None, None, "1mg1-", "0mg-5,2mg1-", "1m1-|0g6", "1mi1-,0d-4"]),
             uam=rng.random() < 0.8, uag=rng.random() < 0.3, pamless=rng.random() < 0.5,
             pam=rng.choice([".NGG", "TTV.", ".NG", "R{1,2}T"]))
    masks = [("m", "a", 2, 2 + glen + 3)]
    crit = O.Criteria(k["ot_mismatch"], k["ot_gap"], k["pattern"], k["uam"], k["uag"], k["pamless"], k["pam"], glen)
    want, _ = O.screen_exhaustive(grnas, contigs, masks, crit)
    got = _general_via_abi(L, contigs, [(0, 2, 2 + glen + 3)], grnas, k, glen)
    assert [x > 0 for x in got.tolist()] == want, (k, grnas, contigs)


def test_three_hundred_thousand_queries(L):
    """Many query chunks (300 000 gRNA = 18 750 groups) against a small genome, every budget class of the scan."""
    rng = random.Random(21)
    contigs = [("x", _random_contig(rng, 6000, "ACGT" * 10 + "N")), ("y", _random_contig(rng, 4000))]
    grnas = [_random_contig(rng, 20) for _ in range(300000)]
    for i in range(0, 300000, 97):
        cid, seq = contigs[i % 2]
        p = rng.randrange(0, len(seq) - 20)
        g = list(seq[p:p + 20].replace("N", "G"))
        for _ in range(i % 4):
            g[rng.randrange(20)] = rng.choice("ACGT")
        grnas[i] = "".join(g)
    for M in (0, 3):
        got = _hamming_via_abi(L, contigs, [(0, 100, 300)], grnas, M)
        want = c_oracle.screen_hamming(contigs, [(0, 100, 300)], grnas, M, threads=16)
        assert np.array_equal(got, want), M
        assert (got > 0).sum() > 300


def _resident_counts(L, genome, q, crit, pam):
    import torch
    counts = torch.zeros(2 * q.n, dtype=torch.int32, device="cuda")
    L.screen(genome, q, crit, pam, counts.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return counts.cpu().numpy().astype(np.int64)


def _medium_world(L, seed, n_grna, glen=20):
    rng = random.Random(seed)
    contigs = [("a", _random_contig(rng, 180000, "ACGT" * 12 + "N")), ("b", _random_contig(rng, 90000)), ("c", _random_contig(rng, 700))]
    grnas = []
    for i in range(n_grna):
        _, seq = contigs[i % 2]
        p = rng.randrange(0, len(seq) - glen - 2)
        g = list(seq[p:p + glen].replace("N", "T"))
        for _ in range(i % 5):
            g[rng.randrange(glen)] = rng.choice("ACGT")
        if i % 3 == 0:                       # one gap
            k = rng.randrange(2, glen - 2)
            g = g[:k] + g[k + 1:] + [rng.choice("ACGT")] if i % 2 else g[:k] + [rng.choice("ACGT")] + g[k:glen - 1]
        g = "".join(g)
        grnas.append(O.reverse_complement(g) if i % 4 == 1 else g)
    genome = L.Genome.from_contigs(contigs)
    genome.set_masks([(0, 1000, 1500), (1, 0, 300)])
    return genome, L.Queries(grnas)


def test_results_do_not_depend_on_the_survivor_queue(L):
    """The verifier queue is sliced by a probe and re-cut on overflow; a tiny queue must give the same counts."""
    from minorg_b200.ot_regex import make_criteria
    from minorg_b200.pam import PAM
    genome, q = _medium_world(L, 31, 96)
    pam = PAM(".NGG", 20).program()
    cases = [make_criteria(ot_mismatch=5, ot_pamless=False, grna_length=20),
             make_criteria(ot_mismatch=3, ot_gap=2, grna_length=20),
             make_criteria(ot_pattern="0mg-10,1mg-11-", ot_gap=2, grna_length=20),
             make_criteria(ot_pattern="1mg1-", ot_gap=2, grna_length=20)]
    try:
        for crit in cases:
            os.environ.pop("MG_QUEUE_CAP", None)
            os.environ.pop("MG_PROBE_MIN", None)
            want = _resident_counts(L, genome, q, crit, pam)
            assert want.sum() > 0
            for cap, probe in (("61", "500"), ("9", "64"), ("4000", "3")):
                os.environ["MG_QUEUE_CAP"], os.environ["MG_PROBE_MIN"] = cap, probe
                assert np.array_equal(_resident_counts(L, genome, q, crit, pam), want), (cap, probe)
    finally:
        os.environ.pop("MG_QUEUE_CAP", None)
        os.environ.pop("MG_PROBE_MIN", None)
        q.free()
        genome.free()


def test_gapped_prefilters_agree(L):
    """Two independent edit-distance prefilters (bit-sliced rows vs one-pattern-per-thread Myers) must hand the
    verifier the same anchors: identical counts over ~10^8 cells."""
    from minorg_b200.ot_regex import make_criteria
    genome, q = _medium_world(L, 32, 160)
    try:
        for kw in (dict(ot_mismatch=2, ot_gap=2), dict(ot_mismatch=5, ot_gap=2), dict(ot_mismatch=3, ot_gap=3),
                   dict(ot_pattern="2mg1-", ot_gap=2)):
            crit = make_criteria(grna_length=20, **kw)
            os.environ.pop("MG_GAPPED_PREFILTER", None)
            rows = _resident_counts(L, genome, q, crit, None)
            os.environ["MG_GAPPED_PREFILTER"] = "myers"
            myers = _resident_counts(L, genome, q, crit, None)
            assert np.array_equal(rows, myers), kw
            assert rows.sum() > 0
    finally:
        os.environ.pop("MG_GAPPED_PREFILTER", None)
        q.free()
        genome.free()


def test_scan_kernels_agree(L):
    """Pair-word scan (default), per-position early-out scan and the per-cell XOR/POPC kernel: same counts, for
    budgets with and without compile-time specialisation, two gRNA lengths."""
    from minorg_b200.ot_regex import make_criteria
    try:
        for glen in (20, 23, 17):
            genome, q = _medium_world(L, 40 + glen, 128, glen)
            for M in (0, 1, 2, 4, 5, 7):
                crit = make_criteria(ot_mismatch=M + 1)
                got = {}
                for mode in ("pairs", "singles", "popc"):
                    os.environ["MG_SCAN"] = mode
                    got[mode] = _resident_counts(L, genome, q, crit, None)
                assert np.array_equal(got["pairs"], got["popc"]) and np.array_equal(got["singles"], got["popc"]), (glen, M)
                assert got["popc"].sum() > 0
            q.free()
            genome.free()
    finally:
        os.environ.pop("MG_SCAN", None)


def test_gapped_counts_vs_c_oracle_medium(L):
    """Gapped / trimmed no-pattern screen at medium size (270 kb, 96 gRNA, ~10^8 cells): the device's per-query ANCHOR
    counts (prefilter + verifier) against the exhaustive C oracle, one to one for both orientations."""
    from minorg_b200.ot_regex import make_criteria
    rng = random.Random(77)
    contigs = [("a", _random_contig(rng, 180000, "ACGT" * 12 + "N")), ("b", _random_contig(rng, 90000)), ("c", _random_contig(rng, 700)),
               ("d", _random_contig(rng, 9))]
    grnas = []
    for i in range(96):
        _, seq = contigs[i % 2]
        p = rng.randrange(0, len(seq) - 24)
        g = list(seq[p:p + 20].replace("N", "T"))
        for _ in range(i % 5):
            g[rng.randrange(20)] = rng.choice("ACGT")
        if i % 3 == 0:
            k = rng.randrange(2, 18)
            g = (g[:k] + g[k + 1:] + [rng.choice("ACGT")]) if i % 2 else (g[:k] + [rng.choice("ACGT")] + g[k:19])
        if i % 11 == 0:                      # hanging off a contig end
            g = list(contigs[1][1][-16:]) + [rng.choice("ACGT") for _ in range(4)]
        g = "".join(g)
        grnas.append(O.reverse_complement(g) if i % 4 == 1 else g)
    masks = [(0, 1000, 1500), (1, 0, 300), (1, 89000, 90000)]
    genome = L.Genome.from_contigs(contigs)
    genome.set_masks(masks)
    q = L.Queries(grnas)
    try:
        for M, Gp in ((2, 1), (4, 1), (1, 2), (0, 1), (3, 0), (2, 3)):
            crit = make_criteria(ot_mismatch=M + 1, ot_gap=Gp + 1, grna_length=20)
            got = _resident_counts(L, genome, q, crit, None)
            want = c_oracle.screen_nopattern(contigs, masks, grnas, M, Gp, threads=16).astype(np.int64)
            assert np.array_equal(got, want), (M, Gp, np.nonzero(got != want)[0][:10], got[got != want][:10], want[got != want][:10])
            assert want.sum() > 5
        # PAM proximity (ot_pamless=False): windows widen with trimming and unused gap allowance (MINORg.py:2063-2087)
        from minorg_b200.pam import PAM
        for pam, M, Gp in ((".NGG", 3, 0), (".NGG", 4, 1), ("TTTV.", 3, 1), ("R{1,2}T", 2, 0), (".", 3, 0), ("TTN", 2, 2)):
            crit = make_criteria(ot_mismatch=M + 1, ot_gap=Gp + 1, ot_pamless=False, grna_length=20)
            got = _resident_counts(L, genome, q, crit, PAM(pam, 20).program())
            want = c_oracle.screen_nopattern(contigs, masks, grnas, M, Gp, threads=16, pam=O.OraclePAM(pam, 20)).astype(np.int64)
            assert np.array_equal(got, want), (pam, M, Gp, np.nonzero(got != want)[0][:10], got[got != want][:10], want[got != want][:10])
            assert want.sum() > 0 or pam == "."
    finally:
        q.free()
        genome.free()


def test_pattern_prefilters_are_sound_at_medium_size(L):
    """--ot-pattern modes: the prefiltered paths (Hamming budget / exact pieces / edit distance, all derived by the host
    from the pattern) against the same verifier run on EVERY anchor (MG_PREFILTER=none), 10^7 cells each."""
    from minorg_b200.ot_regex import make_criteria
    from minorg_b200.pam import PAM
    genome, q = _medium_world(L, 55, 48)
    n = q.n
    try:
        for pat, gap, pamless, exact in (("0mg-10,1mg-11-", 2, True, False), ("1mg1-", 2, True, False), ("2mg1-", 2, True, True),
                                         ("(0mg5,1mg6-)|(0mg6,1m7-)", 0, True, True), ("1m-8,2m9-", 0, True, True),
                                         ("1mg1-", 2, False, False), ("0m-6,3mg1-", 2, True, True)):
            crit = make_criteria(ot_pattern=pat, ot_gap=gap, ot_pamless=pamless, grna_length=20)
            pam = PAM(".NGG", 20).program()
            os.environ.pop("MG_PREFILTER", None)
            fast = _resident_counts(L, genome, q, crit, pam)
            os.environ["MG_PREFILTER"] = "none"
            full = _resident_counts(L, genome, q, crit, pam)
            # every prefilter hands each anchor to the verifier exactly once (the exact-piece prefilter reaches an anchor
            # from several (piece, window) pairs; the first matching pair owns it): counts must agree, not just verdicts
            assert np.array_equal(fast, full), (pat, np.nonzero(fast != full)[0][:10], fast[fast != full][:10], full[fast != full][:10])
            assert full.sum() > 0 or not pamless, pat
    finally:
        os.environ.pop("MG_PREFILTER", None)
        q.free()
        genome.free()


@pytest.mark.parametrize("glen", [7, 17, 23, 27, 31, 32])
def test_gapped_other_lengths_vs_c_oracle(L, glen):
    """The edit-distance rows kernel is instantiated per gRNA length (padded 128-bit table rows, 256 threads and no
    packed verifier window beyond 24 / 30 nt): anchor counts against the exhaustive C oracle for odd and long lengths."""
    from minorg_b200.ot_regex import make_criteria
    rng = random.Random(1000 + glen)
    contigs = [("a", _random_contig(rng, 40000, "ACGT" * 12 + "N")), ("b", _random_contig(rng, 20000)), ("c", _random_contig(rng, glen + 3))]
    grnas = []
    for i in range(40):
        _, seq = contigs[i % 2]
        p = rng.randrange(0, len(seq) - glen - 2)
        g = list(seq[p:p + glen].replace("N", "T"))
        for _ in range(i % 4):
            g[rng.randrange(glen)] = rng.choice("ACGT")
        if i % 3 == 0 and glen > 6:
            k = rng.randrange(2, glen - 2)
            g = g[:k] + g[k + 1:] + [rng.choice("ACGT")] if i % 2 else g[:k] + [rng.choice("ACGT")] + g[k:glen - 1]
        if i % 13 == 0:                      # hanging off a contig end
            g = list(contigs[1][1][-(glen - 3):]) + [rng.choice("ACGT") for _ in range(3)]
        g = "".join(g)
        grnas.append(O.reverse_complement(g) if i % 4 == 1 else g)
    masks = [(0, 1000, 1500), (1, 0, 300)]
    genome = L.Genome.from_contigs(contigs)
    genome.set_masks(masks)
    q = L.Queries(grnas)
    try:
        for M, Gp in ((1, 1), (3, 1), (0, 2)) if glen > 7 else ((0, 1), (1, 1)):
            crit = make_criteria(ot_mismatch=M + 1, ot_gap=Gp + 1, grna_length=glen)
            got = _resident_counts(L, genome, q, crit, None)
            want = c_oracle.screen_nopattern(contigs, masks, grnas, M, Gp, threads=16).astype(np.int64)
            assert np.array_equal(got, want), (glen, M, Gp, np.nonzero(got != want)[0][:10], got[got != want][:10], want[got != want][:10])
            assert want.sum() > 0
    finally:
        q.free()
        genome.free()


@pytest.mark.parametrize("glen", [19, 23])
def test_piece_prefilter_odd_lengths(L, glen):
    """Exact-piece prefilter on an odd gRNA length: the mirrored range of the '-' queries starts at an odd position, so
    single-position words join the pair words.  Counts against the verifier run on every anchor."""
    from minorg_b200.ot_regex import make_criteria
    genome, q = _medium_world(L, 90 + glen, 40, glen)
    n = q.n
    try:
        for pat, gap in (("0mg-10,1mg-11-", 2), ("0mg-12,3mg-13-", 2)):
            crit = make_criteria(ot_pattern=pat, ot_gap=gap, grna_length=glen)
            assert crit.n_pieces > 0, pat
            os.environ.pop("MG_PREFILTER", None)
            fast = _resident_counts(L, genome, q, crit, None)
            os.environ["MG_PREFILTER"] = "none"
            full = _resident_counts(L, genome, q, crit, None)
            assert np.array_equal(fast, full), (glen, pat, np.nonzero(fast != full)[0][:10])
            assert full.sum() > 0, pat
    finally:
        os.environ.pop("MG_PREFILTER", None)
        q.free()
        genome.free()


def test_pattern_counts_vs_independent_c_oracle_medium(L):
    """--ot-pattern at medium size (270 kb x 48 gRNA x 2 strands = 2.6e7 cells per pattern): the device's per-query ANCHOR counts
    (host-derived prefilter + DFS verifier) against oracle_screen_pattern - an independent C enumeration of E1 that parses the
    pattern string itself, builds the reference's per-position tuples as strings (blast.py:84-120) and walks alignments from
    the other end than nothing in the product does; it is pinned on the golden vectors executed from the reference
    (tests/test_oracle_c.py).  Covers BASELINE configs[2]'s pattern, unaligned_as_mismatch=False / unaligned_as_gap=True and
    PAM proximity."""
    from minorg_b200.ot_regex import make_criteria
    from minorg_b200.pam import PAM
    rng = random.Random(55)
    contigs = [("a", _random_contig(rng, 180000, "ACGT" * 12 + "N")), ("b", _random_contig(rng, 90000)), ("c", _random_contig(rng, 700))]
    grnas = []
    for i in range(48):
        _, seq = contigs[i % 2]
        p = rng.randrange(0, len(seq) - 22)
        g = list(seq[p:p + 20].replace("N", "T"))
        for _ in range(i % 5):
            g[rng.randrange(20)] = rng.choice("ACGT")
        if i % 3 == 0:
            k = rng.randrange(2, 18)
            g = g[:k] + g[k + 1:] + [rng.choice("ACGT")] if i % 2 else g[:k] + [rng.choice("ACGT")] + g[k:19]
        g = "".join(g)
        grnas.append(O.reverse_complement(g) if i % 4 == 1 else g)
    masks = [(0, 1000, 1500), (1, 0, 300)]
    genome = L.Genome.from_contigs(contigs)
    genome.set_masks(masks)
    q = L.Queries(grnas)
    cases = [("0mg-10,1mg-11-", 2, True, False, None), ("1mg1-", 2, True, False, None),
             ("(0mg5,1mg6-)|(0mg6,1m7-)", 0, True, False, None), ("1m-8,2m9-", 0, True, False, None),
             ("0m-6,3mg1-", 2, True, False, None), ("1m-8,2m9-", 0, False, False, None), ("2mg1-", 2, False, True, None),
             ("0mg-10,1mg-11-", 2, False, False, None), ("1mg1-", 2, True, True, ".NGG"), ("1mi-8,0d1-", 2, True, False, None)]
    try:
        for pat, gap, uam, uag, pam in cases:
            crit = make_criteria(ot_pattern=pat, ot_gap=gap, ot_unaligned_as_mismatch=uam, ot_unaligned_as_gap=uag,
                                 ot_pamless=pam is None, grna_length=20)
            got = _resident_counts(L, genome, q, crit, PAM(pam, 20).program() if pam else None)
            want = c_oracle.screen_pattern(contigs, masks, grnas, pat, gap, uam, uag, threads=16,
                                           pam=O.OraclePAM(pam, 20) if pam else None).astype(np.int64)
            assert np.array_equal(got, want), (pat, gap, uam, uag, pam, np.nonzero(got != want)[0][:10], got[got != want][:10], want[got != want][:10])
            assert want.sum() > 0 or pam is not None, pat
    finally:
        q.free()
        genome.free()
