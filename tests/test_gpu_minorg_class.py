"""The MINORg-shaped class end to end on BASELINE.json configs[0] (the reference's bundled Arabidopsis subset):
grna() + filter_background() on the GPU vs the CPU oracle applying the reference's criteria."""
import gzip
import os
import shutil

import numpy as np
import pytest

from oracle import c_oracle, minorg_oracle as O

pytestmark = pytest.mark.gpu
CFG1 = os.path.join(os.path.dirname(__file__), "golden", "config1")


@pytest.fixture(scope="module")
def cfg1(tmp_path_factory):
    d = tmp_path_factory.mktemp("cfg1")
    for f in ("subset_ref_TAIR10.fasta", "subset_9654.fasta", "subset_9655.fasta"):
        with gzip.open(os.path.join(CFG1, f + ".gz"), "rb") as i, open(d / f, "wb") as o:
            shutil.copyfileobj(i, o)
    shutil.copy(os.path.join(CFG1, "targets.fasta"), d / "targets.fasta")
    return d


def _oracle_background(d, grna_entries, subjects, ot_mismatch):
    targets = O.read_fasta(str(d / "targets.fasta"))
    seqs = [s for s, _ in grna_entries]
    fail = np.zeros(len(seqs), dtype=bool)
    all_masks = {}
    for f in subjects:
        contigs = O.read_fasta(str(d / f))
        names = [c[0] for c in contigs]
        masks = O.find_masks(targets, contigs)
        all_masks[f] = masks
        counts = c_oracle.screen_hamming(contigs, [(names.index(c), s, e) for _, c, s, e in masks], seqs,
                                         max(1, ot_mismatch) - 1, threads=8)
        fail |= counts > 0
    return fail, all_masks


@pytest.mark.parametrize("ot_mismatch", [1, 2, 3])
def test_config1_grna_and_filter_background(cfg1, tmp_path, ot_mismatch, capsys):
    from minorg_b200.MINORg import MINORg
    m = MINORg(directory=str(tmp_path), prefix="cfg1")
    m.target = str(cfg1 / "targets.fasta")
    m.reference = {"TAIR10": str(cfg1 / "subset_ref_TAIR10.fasta")}
    m.add_query(str(cfg1 / "subset_9654.fasta"), alias="9654")
    m.add_query(str(cfg1 / "subset_9655.fasta"), alias="9655")
    m.ot_mismatch = ot_mismatch
    m.grna()
    want = O.find_multi_grna(m.target, ".NGG", 20)
    assert m.grna_hits.seqs == [s for s, _ in want] and len(want) == 733
    assert open(m.grna_fasta).read() == O.grna_fasta_text(want)
    assert open(m.grna_map).read() == O.grna_map_text(want)
    m.filter_background()
    assert "Masking on-targets" in capsys.readouterr().out
    subjects = ["subset_ref_TAIR10.fasta", "subset_9654.fasta", "subset_9655.fasta"]
    fail, masks = _oracle_background(cfg1, want, subjects, ot_mismatch)
    got_fail = np.array([m.grna_hits.get_gRNAseq_by_seq(s).check("background") is False for s, _ in want])
    assert np.array_equal(got_fail, fail)
    assert 0 < fail.sum() < len(fail)                        # both outcomes occur
    for f in subjects:
        got = {(x.masked, x.molecule, x.start, x.end) for x in m.masked[str(cfg1 / f)]}
        assert got == masks[f]
    assert ("AT5G45050", "chromA", 13, 4905) in masks["subset_ref_TAIR10.fasta"]
    # files: .map carries the background column, pass files hold exactly the passing gRNA, mask report format
    status = {(s, "background"): ("fail" if fl else "pass") for (s, _), fl in zip(want, fail)}
    assert open(m.grna_map).read() == O.grna_map_text(want, status=status)
    passed = [l for l in open(m.pass_fasta).read().splitlines() if not l.startswith(">")]
    assert passed == [s for (s, _), fl in zip(want, fail) if not fl]
    rep = open(os.path.join(str(tmp_path), "cfg1", "cfg1_masked.txt")).read().splitlines()
    assert rep[0] == "#alias\tfname" and "source\tmolecule\tstart\tend\tmasked" in rep
    assert any(l.split("\t")[:4] == ["TAIR10", "chromA", "13", "4905"] for l in rep)
    # the steps after the path: GC filter, then mutually exclusive minimum sets of passing gRNA
    m.filter_gc()
    gc_ok = {s for s, _ in want if 0.3 <= (s.count("G") + s.count("C")) / 20 <= 0.7}
    assert {s for s, _ in want if m.grna_hits.get_gRNAseq_by_seq(s).check("GC")} == gc_ok
    sets = m.minimumset(sets=2)
    passing = {s for (s, _), fl in zip(want, fail) if not fl} & gc_ok
    assert len(sets) == 2 and not set(sets[0]) & set(sets[1])
    for group in sets:
        assert set(group) <= passing
        assert {h.target_id for s in group for h in m.grna_hits.get_hits(s)} == {"AT5G45050", "AT5G45060"}
    final = [l.split("\t") for l in open(m.final_map).read().splitlines()[1:]]
    assert {r[1] for r in final} == set(sets[0]) | set(sets[1]) and {r[8] for r in final} == {"1", "2"}
    m.close()


def test_config1_pattern_and_pam_modes(cfg1, tmp_path):
    """Same data through the general screen: --ot-pattern, gaps, PAM proximity (flags vs the Python oracle on a
    small slice so the exhaustive enumeration stays cheap)."""
    from minorg_b200.MINORg import MINORg
    ref = O.read_fasta(str(cfg1 / "subset_ref_TAIR10.fasta"))
    small = tmp_path / "small.fasta"
    with open(small, "w") as f:
        f.write(">chromA\n%s\n>chromB\n%s\n" % (ref[0][1][4800:5150], ref[1][1][:90]))
    tgt = tmp_path / "t.fasta"
    with open(tgt, "w") as f:
        f.write(">t1\n%s\n" % ref[0][1][4850:4990])
    for kw in (dict(ot_mismatch=2, ot_gap=2), dict(ot_mismatch=2, ot_pamless=False), dict(ot_pattern="0mg-8,2mg1-", ot_gap=2),
               dict(ot_mismatch=1, ot_gap=2, ot_pamless=False, pam="TTV.")):
        m = MINORg(directory=str(tmp_path), prefix="p")
        m.length = 14
        for k, v in kw.items():
            setattr(m, k, v)
        m.target = str(tgt)
        m.add_background(str(small), alias="bg")
        m.screen_reference = False
        m.grna()
        seqs = m.grna_hits.seqs[:6]
        m.grna_hits._gRNAseqs = {s: m.grna_hits._gRNAseqs[s] for s in seqs}
        m.grna_hits._hits = {s: m.grna_hits._hits[s] for s in seqs}
        m.write_all_grna_fasta()
        m.filter_background()
        contigs = O.read_fasta(str(small))
        masks = O.find_masks(O.read_fasta(str(tgt)), contigs)
        crit = O.Criteria(kw.get("ot_mismatch", 1), kw.get("ot_gap", 0), kw.get("ot_pattern"), True, False,
                          kw.get("ot_pamless", True), kw.get("pam", ".NGG"), 14)
        want, _ = O.screen_exhaustive(seqs, contigs, masks, crit)
        got = [m.grna_hits.get_gRNAseq_by_seq(s).check("background") is False for s in seqs]
        assert got == want, kw
        m.close()


def test_verdicts_only_mode_gives_the_same_verdicts(cfg1, tmp_path):
    """screen_verdicts_only (not a reference attribute; off by default): window slices with the failed gRNA dropped from
    the query set give the same pass/fail as the exhaustive counting screen, for the ungapped and the gapped criterion."""
    from minorg_b200.MINORg import MINORg
    for kw in (dict(ot_mismatch=2), dict(ot_mismatch=2, ot_gap=2)):
        verdicts = []
        for fast in (False, True):
            m = MINORg(directory=str(tmp_path / ("v%d%d" % (fast, len(kw)))), prefix="v")
            for k, v in kw.items():
                setattr(m, k, v)
            m.screen_verdicts_only = fast
            m.target = str(cfg1 / "targets.fasta")
            m.reference = {"TAIR10": str(cfg1 / "subset_ref_TAIR10.fasta")}
            m.add_query(str(cfg1 / "subset_9654.fasta"), alias="9654")
            m.grna()
            m.filter_background()
            verdicts.append([m.grna_hits.get_gRNAseq_by_seq(s).check("background") for s in m.grna_hits.seqs])
            m.close()
        assert verdicts[0] == verdicts[1], kw
        assert True in verdicts[0] and False in verdicts[0], kw


def test_cli_full_then_resume_from_map(cfg1, tmp_path):
    """`python -m minorg_b200 full ...` and a second `filter` + `minimumset` run resumed from the .map file."""
    from minorg_b200.__main__ import main
    out = tmp_path / "cli"
    common = ["-d", str(out), "--prefix", "run", "-t", str(cfg1 / "targets.fasta")]
    subj = ["-q", str(cfg1 / "subset_9654.fasta"), "-q", str(cfg1 / "subset_9655.fasta"),
            "--assembly", str(cfg1 / "subset_ref_TAIR10.fasta"), "--screen-ref", "--ot-mismatch", "2"]
    assert main(["full"] + common + subj + ["-s", "2"]) == 0
    all_map = out / "run" / "run_gRNA_all.map"
    rows = [l.split("\t") for l in open(all_map).read().splitlines()]
    assert rows[0][-3:] == ["background", "GC", "feature"]
    want = O.find_multi_grna(str(cfg1 / "targets.fasta"), ".NGG", 20)
    fail, _ = _oracle_background(cfg1, want, ["subset_ref_TAIR10.fasta", "subset_9654.fasta", "subset_9655.fasta"], 2)
    bg = {r[1]: r[-3] for r in rows[1:]}
    assert bg == {s: ("fail" if fl else "pass") for (s, _), fl in zip(want, fail)}
    final1 = open(out / "run" / "run_gRNA_final.map").read()
    assert len({l.split("\t")[8] for l in final1.splitlines()[1:]}) == 2
    # resume: re-screen from the .map with a stricter budget, then pick sets again
    assert main(["filter", "-m", str(all_map)] + common + subj[:-1] + ["3"]) == 0
    rows2 = [l.split("\t") for l in open(all_map).read().splitlines()]
    fail3, _ = _oracle_background(cfg1, want, ["subset_ref_TAIR10.fasta", "subset_9654.fasta", "subset_9655.fasta"], 3)
    assert {r[1]: r[-3] for r in rows2[1:]} == {s: ("fail" if fl else "pass") for (s, _), fl in zip(want, fail3)}
    assert main(["minimumset", "-m", str(all_map), "-d", str(out), "--prefix", "run", "-s", "1"]) == 0


def test_class_rejects_device_limits_by_name(tmp_path):
    """gRNA longer than MG_MAX_GRNA_LEN and ot_gap beyond MG_MAX_GAPS fail in the MINORg-shaped class with a MINORgError
    that names the limit - before any device call (the reference itself has no such limits, generate_grna.py:23)."""
    import pytest
    from minorg_b200.MINORg import MINORg, MINORgError
    t = tmp_path / "t.fasta"
    t.write_text(">t1\n" + "ACGTTGCAAGGCTTAGCGG" * 12 + "\n")
    m = MINORg(directory=str(tmp_path / "out"), prefix="x")
    m.target = str(t)
    m.length = 33
    with pytest.raises(MINORgError, match="MG_MAX_GRNA_LEN"):
        m.grna()
    m.length = 20
    m.ot_gap = 6
    with pytest.raises(MINORgError, match="MG_MAX_GAPS"):
        m._criteria(20)
    m.ot_gap = 4
    m._criteria(20)


def test_large_grna_set_takes_the_seed_and_join_screen_with_the_same_verdicts(tmp_path):
    """grna() on a 700-kb target gives > 40 000 gRNA: filter_background with the one-gap criterion then runs the seed-and-join
    screen by itself (mg_last_seedjoin_stats > 0); the verdicts equal those of the exhaustive DP path (MG_SEEDJOIN=off)."""
    import random
    from minorg_b200 import _lib
    from minorg_b200.MINORg import MINORg
    rng = random.Random(77)
    target = "".join(rng.choice("ACGT") for _ in range(700000))
    chrom = [rng.choice("ACGT") for _ in range(1500000)]
    for _ in range(300):                                  # diverged copies of target pieces: off-targets with mismatches and gaps
        a = rng.randrange(0, len(target) - 400)
        piece = list(target[a:a + 400])
        for _ in range(rng.randrange(5, 40)):
            j = rng.randrange(len(piece))
            r = rng.random()
            if r < 0.6:
                piece[j] = rng.choice("ACGT")
            elif r < 0.8:
                del piece[j]
            else:
                piece.insert(j, rng.choice("ACGT"))
        b = rng.randrange(0, len(chrom) - 500)
        chrom[b:b + len(piece)] = piece
    (tmp_path / "t.fasta").write_text(">gene\n%s\n" % target)
    (tmp_path / "bg.fasta").write_text(">chr1\n%s\n>chr2\n%s\n" % ("".join(chrom), target[:300000]))
    verdicts, ran = [], []
    for mode in ("", "off"):
        if mode:
            os.environ["MG_SEEDJOIN"] = mode
        try:
            m = MINORg(directory=str(tmp_path / ("run" + mode)), prefix="big")
            m.target = str(tmp_path / "t.fasta")
            m.add_background(str(tmp_path / "bg.fasta"), alias="bg")
            m.ot_mismatch, m.ot_gap = 3, 2
            m.grna()
            assert len(m.grna_hits.seqs) > 40000
            m.filter_background()
            ran.append(_lib.last_seedjoin_stats()[0] > 0)
            verdicts.append([m.grna_hits.get_gRNAseq_by_seq(s).check("background") for s in m.grna_hits.seqs])
            m.close()
        finally:
            os.environ.pop("MG_SEEDJOIN", None)
    assert ran[0]
    assert verdicts[0] == verdicts[1]
    assert 100 < sum(v is False for v in verdicts[0]) < len(verdicts[0])


def test_minimumset_prioritise_nr_through_the_class_and_cli(cfg1, tmp_path):
    """self.prioritise_nr / --prioritise-nr: mutually exclusive sets that each cover every target, equal to what
    minimum_set_nr.minimum_sets_nr picks from the passing gRNA (itself pinned to the reference on the CPU)."""
    from minorg_b200.MINORg import MINORg
    from minorg_b200.__main__ import main
    from minorg_b200.minimum_set_nr import minimum_sets_nr
    m = MINORg(directory=str(tmp_path / "nr"), prefix="nr")
    m.target = str(cfg1 / "targets.fasta")
    m.add_query(str(cfg1 / "subset_9654.fasta"), alias="9654")
    m.ot_mismatch = 2
    m.prioritise_nr = True
    m.grna()
    m.filter_background()
    sets = m.minimumset(sets=3)
    targets = {h.target_id for h in m.grna_hits.flatten_hits()}
    assert len(sets) == 3
    used = [s for g in sets for s in g]
    assert len(used) == len(set(used))
    for g in sets:
        assert {h.target_id for s in g for h in m.grna_hits.get_hits(s)} >= targets
    assert sets == minimum_sets_nr(m.grna_hits, list(m.valid_grna()), 3, targets)
    rows = [l.split("\t") for l in open(m.final_map).read().splitlines()[1:]]
    assert {r[1] for r in rows} == set(used) and {r[8] for r in rows} == {"1", "2", "3"}
    # manual mode through the class: scripted terminal answers (reject the first listed gRNA of set 1, then accept everything)
    import builtins
    answers = iter([open(m.final_map).read().splitlines()[1].split("\t")[1], "x", "x", "x", "x"])
    real_input = builtins.input
    builtins.input = lambda *_a: next(answers)
    try:
        m2 = MINORg(directory=str(tmp_path / "manual"), prefix="man")
        m2.target = str(cfg1 / "targets.fasta")
        m2.add_query(str(cfg1 / "subset_9654.fasta"), alias="9654")
        m2.ot_mismatch = 2
        m2.grna()
        m2.filter_background()
        auto = m2.minimumset(sets=1)
        rejected = sorted((m2.grna_hits.get_gRNAseq_by_seq(s).id, s) for s in auto[0])[0][1]
        answers = iter([rejected, "x"])
        manual = m2.minimumset(sets=1, manual=True)
        assert rejected in auto[0] and rejected not in manual[0]
        assert {h.target_id for s in manual[0] for h in m2.grna_hits.get_hits(s)} >= targets
        m2.close()
    finally:
        builtins.input = real_input
    m.close()
    out = tmp_path / "cli_nr"
    assert main(["full", "-d", str(out), "--prefix", "run", "-t", str(cfg1 / "targets.fasta"), "-q", str(cfg1 / "subset_9654.fasta"),
                 "--ot-mismatch", "2", "-s", "3", "--prioritise-nr"]) == 0
    rows2 = [l.split("\t") for l in open(out / "run" / "run_gRNA_final.map").read().splitlines()[1:]]
    # (the CLI also applies the GC filter, so its candidate pool differs: three exclusive covers of every target all the same)
    by_set = {}
    for r in rows2:
        by_set.setdefault(r[8], []).append(r)
    assert set(by_set) == {"1", "2", "3"}
    assert all({r[2] for r in rs} >= targets for rs in by_set.values())
    seqs_of = [{r[1] for r in rs} for rs in by_set.values()]
    assert not (seqs_of[0] & seqs_of[1]) and not (seqs_of[0] & seqs_of[2]) and not (seqs_of[1] & seqs_of[2])
