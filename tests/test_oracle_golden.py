"""The CPU oracle against golden vectors produced by executing the reference's own code
(tests/golden/make_golden.py) and against the reference's documented known answers."""
import os

import pytest

from oracle import minorg_oracle as O

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_pam_vectors(golden):
    d = golden("pam.json")
    n = 0
    for r in d["rows"]:
        if "error" in r:
            with pytest.raises(Exception):
                O.OraclePAM(r["pam"], r["L"]).regex_string()
            continue
        p = O.OraclePAM(r["pam"], r["L"])
        assert p.pam == r["parsed"], r
        assert p.regex_string() == r["regex"], r
        assert p.five_prime() == r["five"] and p.three_prime() == r["three"], r
        assert p.five_prime_maxlen() == r["five_max"], r
        assert p.three_prime_maxlen() == r["three_max"], r
        n += 1
    assert n > 90


def test_ot_range_vectors(golden):
    for r in golden("ot_range.json")["rows"]:
        if r.get("error"):
            with pytest.raises(O.OracleError):
                O.ot_range(r["range"])
        else:
            assert list(O.ot_range(r["range"])) == r["out"], r


def test_alignment_strings_and_patterns(golden):
    d = golden("ot_pattern.json")
    checked = 0
    for a in d["rows"]:
        aln = O.Alignment.from_btop(a["btop"], a["qstart"], a["qend"], a["qlen"])
        assert aln.tuple_fwd == a["tbtop"], a["btop"]
        assert aln.tuple_rvs == a["tbtop_rvs"], a["btop"]
        assert " " * a["qstart"] + aln.kinds + " " * (a["qlen"] - a["qend"]) == a["sbtop"]
        if a["kinds"] is not None:
            assert aln.kinds == a["kinds"]
        for key, want in a["results"].items():
            pat, flags = key.rsplit("|", 1)
            uam, uag = flags[0] == "1", flags[1] == "1"
            if isinstance(want, str):
                with pytest.raises(Exception):
                    O.OffTargetPattern(pat, uam, uag)(aln)
            else:
                assert O.OffTargetPattern(pat, uam, uag)(aln) == want, (a["btop"], a["qstart"], a["qend"], key)
                checked += 1
    assert checked > 5000
    aln = O.Alignment.from_btop("20", 0, 20, 20)
    for pat, want in d["odd_patterns"].items():
        if isinstance(want, str):
            with pytest.raises(Exception):
                O.OffTargetPattern(pat)(aln)
        else:
            assert O.OffTargetPattern(pat)(aln) == want, pat


def test_reference_commented_dummy_vectors():
    # minorg/ot_regex.py:37-48 (expected values in the reference's own comments)
    aln = O.Alignment.from_btop("1AG-T12-TAG1", 1, 21, 20)
    assert O.OffTargetPattern("0g4|(2m4,0g5)")(aln) is False
    assert O.OffTargetPattern("0g4|(2m4,1g5)")(aln) is True


def test_nopattern_and_within(golden):
    d = golden("criteria.json")
    for r in d["nopattern"]:
        got = O.nopattern_problematic(r["mm"], r["gaps"], r["qstart"], r["qend"], r["L"], r["ot_mismatch"], r["ot_gap"])
        assert got == r["problematic"], r
    for r in d["within"]:
        assert O.within_mask(tuple(r["hit"]), tuple(r["mask"])) == r["within"], r


def test_enumeration_vectors(golden):
    d = golden("enumerate.json")
    for r in d["rows"]:
        path = os.path.join(GOLDEN, r["fasta"])
        got = O.find_multi_grna(path, r["pam"], r["L"])
        assert len(got) == r["n_seqs"], (r["fasta"], r["pam"], r["L"])
        assert sum(len(h) for _, h in got) == r["n_hits"]
        if "hits" in r:
            assert [[s, [list(x) for x in h]] for s, h in got] == r["hits"]
        if "fasta_text" in r:
            assert O.grna_fasta_text(got) == r["fasta_text"]
            assert O.grna_map_text(got) == r["map_text"]


def test_documented_known_answers():
    # docs/source/tutorial_py.rst:517-531
    path = os.path.join(GOLDEN, "sample_CDS.fasta")
    assert len(O.find_multi_grna(path, ".NGG", 20)) == 201
    assert len(O.find_multi_grna(path, "TTTV.", 20)) == 95
    assert len(O.find_multi_grna(path, "ATV.", 20)) == 267
    assert O.OraclePAM(".NGG", 20).regex_string() == ".{20}(?=[GATC]GG)"
    assert O.OraclePAM("TTTV.", 20).regex_string() == "(?<=TTT[ACG]).{20}"
    assert O.OraclePAM("ATV.", 20).regex_string() == "(?<=AT[ACG]).{20}"


def test_screen_cases_against_reference_predicates(golden):
    """Tiny genomes: every alignment of the E1 universe was judged by the reference's own
    _is_offtarget_pos/_is_offtarget_aln_*/_has_pam when the fixture was made."""
    cases = golden("screen_cases.json")["cases"]
    assert len(cases) >= 25
    for c in cases:
        k = c["criteria"]
        crit = O.Criteria(k["ot_mismatch"], k["ot_gap"], k.get("pattern"), k.get("uam", True), k.get("uag", False),
                          k["pamless"], k["pam"], c["L"])
        fails, counts = O.screen_exhaustive(c["grnas"], [tuple(x) for x in c["contigs"]], c["masks"], crit)
        assert fails == c["fail"], c["name"]
        assert counts == c["n_problematic_alignments"], c["name"]
        if k.get("pattern") is None and k["pamless"] and k["ot_gap"] <= 1:
            M = max(1, k["ot_mismatch"]) - 1
            hc = O.screen_hamming(c["grnas"], [tuple(x) for x in c["contigs"]], c["masks"], M)
            assert [x > 0 for x in hc] == c["fail"], c["name"]
