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
