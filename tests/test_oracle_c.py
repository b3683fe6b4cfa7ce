"""The C oracle (oracle/screen_oracle.c) against the Python oracle and the golden screen cases."""
import random

from oracle import c_oracle, minorg_oracle as O


def test_c_oracle_matches_golden_cases(golden):
    n = 0
    for c in golden("screen_cases.json")["cases"]:
        k = c["criteria"]
        if k.get("pattern") is not None or not k["pamless"] or k["ot_gap"] > 1:
            continue
        M = max(1, k["ot_mismatch"]) - 1
        names = [x[0] for x in c["contigs"]]
        masks = [(names.index(m[1]), m[2], m[3]) for m in c["masks"]]
        got = c_oracle.screen_hamming([tuple(x) for x in c["contigs"]], masks, c["grnas"], M, threads=2)
        want = O.screen_hamming(c["grnas"], [tuple(x) for x in c["contigs"]], c["masks"], M)
        assert got.tolist() == want, c["name"]
        assert [x > 0 for x in got.tolist()] == c["fail"], c["name"]
        n += 1
    assert n >= 7


def test_c_oracle_random_vs_python():
    rng = random.Random(5)
    for trial in range(6):
        L = rng.choice([8, 20, 23, 32])
        contigs = [("c%d" % i, "".join(rng.choice("ACGTacgtN") for _ in range(rng.randrange(1, 300)))) for i in range(3)]
        grnas = []
        for _ in range(10):
            cid, seq = rng.choice(contigs)
            if len(seq) > L and rng.random() < 0.7:
                p = rng.randrange(0, len(seq) - L)
                g = list(seq[p:p + L].upper())
                for _m in range(rng.randrange(0, 3)):
                    g[rng.randrange(L)] = rng.choice("ACGT")
                g = "".join(g)
                grnas.append(O.reverse_complement(g) if rng.random() < 0.5 else g)
            else:
                grnas.append("".join(rng.choice("ACGT") for _ in range(L)))
        masks = [("m", "c0", 3, 3 + L + 2), ("m", "c1", 0, 10)]
        M = rng.choice([0, 1, 2, 4])
        want = O.screen_hamming(grnas, contigs, masks, M)
        got = c_oracle.screen_hamming(contigs, [(0, 3, 3 + L + 2), (1, 0, 10)], grnas, M, threads=3)
        assert got.tolist() == want, trial


def test_c_general_oracle_matches_golden_and_python(golden):
    """oracle_screen_nopattern (gaps + trimming, PAM-less): fail flags vs the verdicts of the reference's predicates on
    the golden cases, and anchor counts vs the Hamming oracle where the two criteria coincide."""
    n = 0
    for c in golden("screen_cases.json")["cases"]:
        k = c["criteria"]
        if k.get("pattern") is not None:
            continue
        M, Gp = max(1, k["ot_mismatch"]) - 1, max(1, k["ot_gap"]) - 1
        contigs = [tuple(x) for x in c["contigs"]]
        names = [x[0] for x in contigs]
        masks = [(names.index(m[1]), m[2], m[3]) for m in c["masks"]]
        pam = None if k["pamless"] else O.OraclePAM(k["pam"], c["L"])
        got = c_oracle.screen_nopattern(contigs, masks, c["grnas"], M, Gp, threads=2, pam=pam)
        ng = len(c["grnas"])
        assert [(got[i] + got[ng + i]) > 0 for i in range(ng)] == c["fail"], c["name"]
        if Gp == 0 and k["pamless"]:
            ham = c_oracle.screen_hamming(contigs, masks, c["grnas"], M, threads=2)
            assert (got[:ng] + got[ng:]).tolist() == ham.tolist(), c["name"]
        n += 1
    assert n >= 17


def _kinds(row):
    if row["kinds"] is not None:
        return row["kinds"]
    return O.Alignment.from_btop(row["btop"], row["qstart"], row["qend"], row["qlen"]).kinds


def test_c_pattern_evaluator_matches_reference_vectors(golden):
    """oracle_pattern_eval (the C restatement of ot_regex.OffTargetExpression + blast.BlastHSP tuples, parsing the pattern
    string itself) on every golden evaluation produced by executing the reference's own classes, parse errors included."""
    d = golden("ot_pattern.json")
    n = 0
    for row in d["rows"]:
        for key, want in row["results"].items():
            pat, flags = key.rsplit("|", 1)
            got = c_oracle.pattern_eval(pat, _kinds(row), row["qstart"], row["qend"], row["qlen"], flags[0] == "1", flags[1] == "1")
            assert (got is None) if isinstance(want, str) else (got == want), (pat, flags, row)
            n += 1
    assert n > 10000
    for pat, want in d["odd_patterns"].items():
        assert c_oracle.pattern_parse_ok(pat) == (not isinstance(want, str)), pat


def test_c_pattern_screen_matches_golden_cases_and_python(golden):
    """oracle_screen_pattern: fail flags against the golden screens (verdicts of the reference's predicates over E1) and the
    Python oracle; the pruned enumeration equals the unpruned one anchor for anchor."""
    n = 0
    for c in golden("screen_cases.json")["cases"]:
        k = c["criteria"]
        if k.get("pattern") is None:
            continue
        contigs = [tuple(x) for x in c["contigs"]]
        names = [x[0] for x in contigs]
        masks = [(names.index(m[1]), m[2], m[3]) for m in c["masks"]]
        pam = None if k["pamless"] else O.OraclePAM(k["pam"], c["L"])
        ng = len(c["grnas"])
        got = c_oracle.screen_pattern(contigs, masks, c["grnas"], k["pattern"], k["ot_gap"], k["uam"], k["uag"], threads=2, pam=pam)
        assert [(got[i] + got[ng + i]) > 0 for i in range(ng)] == c["fail"], c["name"]
        slow = c_oracle.screen_pattern(contigs, masks, c["grnas"], k["pattern"], k["ot_gap"], k["uam"], k["uag"], threads=2, pam=pam,
                                       prune=False)
        assert got.tolist() == slow.tolist(), c["name"]
        n += 1
    assert n >= 12
    rng = random.Random(17)
    for trial, (pat, gap, uam, uag) in enumerate([("0mg-4,1mg-5-", 2, True, False), ("1mg1-", 2, True, True), ("1m3,1g4-", 2, False, False),
                                                  ("(0mg3,1mg4-)|(0m-3,1g1-)", 2, True, False), ("0d3,1mi4-", 3, True, False),
                                                  ("1m-4--2|0g2-5", 2, False, True)]):
        L = 8
        contigs = [("c%d" % i, "".join(rng.choice("ACGTacgtN") for _ in range(rng.randrange(5, 40)))) for i in range(2)]
        grnas = []
        for _ in range(6):
            _, seq = rng.choice(contigs)
            if len(seq) > L + 1 and rng.random() < 0.8:
                p = rng.randrange(0, len(seq) - L)
                g = list(seq[p:p + L].upper().replace("N", "A"))
                if rng.random() < 0.5:
                    g[rng.randrange(L)] = rng.choice("ACGT")
                if rng.random() < 0.4:
                    k = rng.randrange(1, L - 1)
                    g = g[:k] + g[k + 1:] + [rng.choice("ACGT")]
                g = "".join(g)
                grnas.append(O.reverse_complement(g) if rng.random() < 0.5 else g)
            else:
                grnas.append("".join(rng.choice("ACGT") for _ in range(L)))
        masks = [("m", "c0", 2, 2 + L)]
        crit = O.Criteria(ot_gap=gap, pattern=pat, uam=uam, uag=uag, length=L)
        want, _ = O.screen_exhaustive(grnas, contigs, masks, crit)
        got = c_oracle.screen_pattern(contigs, [(0, 2, 2 + L)], grnas, pat, gap, uam, uag, threads=2)
        ng = len(grnas)
        assert [(got[i] + got[ng + i]) > 0 for i in range(ng)] == want, (trial, pat)
        slow = c_oracle.screen_pattern(contigs, [(0, 2, 2 + L)], grnas, pat, gap, uam, uag, threads=2, prune=False)
        assert got.tolist() == slow.tolist(), (trial, pat)


def test_bitsliced_cpu_port_equals_the_scalar_oracle():
    """oracle_screen_hamming_bitsliced (the CPU baseline bench.py times) returns the counts of oracle_screen_hamming_win on
    ragged inputs: N runs, lower case, short contigs, masks, window ranges, odd lengths, every budget class."""
    import numpy as np
    rng = random.Random(31)
    for L, M in ((20, 4), (20, 0), (23, 3), (8, 1), (32, 5), (1, 0), (21, 6), (20, 7)):
        contigs = [("a", "".join(rng.choice("ACGTacgtN" + "ACGT" * 6) for _ in range(9000))), ("b", "".join(rng.choice("ACGT") for _ in range(700))),
                   ("c", "ACG"), ("d", "")]
        grnas = []
        for i in range(150):
            if i % 2:
                p = rng.randrange(0, 9000 - L)
                s = list(contigs[0][1][p:p + L].upper().replace("N", "A"))
                for _ in range(rng.randrange(0, M + 2)):
                    s[rng.randrange(L)] = rng.choice("ACGT")
                grnas.append("".join(s))
            else:
                grnas.append("".join(rng.choice("ACGTN" + "ACGT" * 5) for _ in range(L)))
        blob = np.frombuffer("".join(s for _, s in contigs).encode(), dtype=np.uint8)
        off = np.array([0, 9000, 9700, 9703, 9703], dtype=np.int64)
        masks = [(0, 100, 400), (1, 0, 50)]
        for ranges in (None, [(-5, 4000), (10, 600), (0, 1), (0, 0)]):
            a = c_oracle.screen_hamming_arrays(blob, off, masks, grnas, M, 3, ranges)
            b = c_oracle.screen_hamming_arrays(blob, off, masks, grnas, M, 3, ranges, bitsliced=True)
            assert a.tolist() == b.tolist(), (L, M, ranges)
        assert a.sum() > 0


def test_gapped_and_pattern_oracles_window_ranges_partition():
    """win_ranges of oracle_screen_nopattern_win / oracle_screen_pattern (the C side of mg_screen's window shards for the
    general path): an anchor belongs to the window it ends at, so the counts of a partition of the window starts add up to
    the whole, on both strands, across masks and contig ends."""
    rng = random.Random(4)
    contigs = [("a", "".join(rng.choice("ACGT" * 8 + "N") for _ in range(3000))), ("b", "".join(rng.choice("ACGT") for _ in range(900)))]
    grnas = []
    for i in range(24):
        _, seq = contigs[i % 2]
        p = rng.randrange(0, len(seq) - 22)
        s = list(seq[p:p + 20].replace("N", "T"))
        for _ in range(i % 3):
            s[rng.randrange(20)] = rng.choice("ACGT")
        if i % 3 == 0:
            k = rng.randrange(2, 18)
            s = s[:k] + s[k + 1:] + [rng.choice("ACGT")]
        s = "".join(s)
        grnas.append(O.reverse_complement(s) if i % 4 == 1 else s)
    masks = [(0, 100, 160)]
    cuts = [-40, 500, 501, 1777, 3100]
    parts = [[(a, b), (a - 200, b - 200)] for a, b in zip(cuts[:-1], cuts[1:])]
    full = c_oracle.screen_nopattern(contigs, masks, grnas, 2, 1, threads=2)
    assert full.sum() > 10
    assert sum(c_oracle.screen_nopattern(contigs, masks, grnas, 2, 1, threads=2, win_ranges=r) for r in parts).tolist() == full.tolist()
    for pat, gap in (("1mg1-", 2), ("0mg-10,1mg-11-", 2), ("1m-8,2m9-", 0)):
        full = c_oracle.screen_pattern(contigs, masks, grnas, pat, gap, threads=2)
        assert full.sum() > 0
        assert sum(c_oracle.screen_pattern(contigs, masks, grnas, pat, gap, threads=2, win_ranges=r) for r in parts).tolist() == full.tolist(), pat
