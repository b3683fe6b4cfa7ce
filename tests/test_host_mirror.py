"""Host-side mirrors of the reference interface (minorg_b200.pam / ot_regex) against the golden vectors
produced by executing the reference classes.  No GPU and no compute calls into the shared library."""
import pytest

from minorg_b200 import ot_regex as R
from minorg_b200.pam import PAM, PAM_PATTERNS
from oracle import minorg_oracle as O


def test_pam_mirror(golden, capsys):
    d = golden("pam.json")
    for r in d["rows"]:
        if "error" in r:
            with pytest.raises(Exception):
                PAM(r["pam"], r["L"]).regex()
            continue
        p = PAM(r["pam"], r["L"])
        assert p.pam == r["parsed"] and repr(p) == r["parsed"]
        assert p.regex().pattern == r["regex"]
        assert (p.five_prime(), p.three_prime()) == (r["five"], r["three"])
        assert (p.five_prime_maxlen(), p.three_prime_maxlen()) == (r["five_max"], r["three_max"])
    assert "PAM pattern: .{20}(?=[GATC]GG)" in capsys.readouterr().out      # pam.py:124 prints
    src = PAM("TTTV.", 23)
    for c in d["copy_ctor"]:
        assert PAM(src, c["arg_L"]).gRNA_length == c["got_L"]
    assert PAM_PATTERNS["Cas12a"] == "TTTV." and PAM_PATTERNS["spcas9"] == ".NGG"


def test_pam_program_lowering():
    p = PAM("R{1,2}T", 20).program()
    assert p.five.n_alts == 0 and p.three.n_alts == 2 and p.three_maxlen == 9
    assert sorted(p.three.len[i] for i in range(2)) == [3, 4]
    q = PAM("TTTV.", 23).program()
    assert q.five.n_alts == 1 and q.five.len[0] == 4 and q.three.n_alts == 0 and q.five_maxlen == 4
    assert q.five.allow4[0][3] == 0b0111 and q.five.allow4[0][0] == 0b1000
    assert (q.five.allow[0][0][ord("t") >> 5] >> (ord("t") & 31)) & 1


def test_ot_range_mirror(golden):
    for r in golden("ot_range.json")["rows"]:
        if r.get("error"):
            with pytest.raises(R.MINORgError):
                R.ot_range(r["range"])
        else:
            assert list(R.ot_range(r["range"])) == r["out"], r


class _HSP:
    """What the reference's BlastHSP offers to OffTargetExpression (blast.py:151-172)."""

    def __init__(self, aln):
        self.aln = aln

    def _num_char_in_range(self, *chars, start=None, end=None, rvs_index=False):
        return self.aln.count(chars, start, end, rvs_index)


def test_pattern_mirror(golden):
    d = golden("ot_pattern.json")
    n = 0
    for a in d["rows"]:
        hsp = _HSP(O.Alignment.from_btop(a["btop"], a["qstart"], a["qend"], a["qlen"]))
        for key, want in a["results"].items():
            pat, flags = key.rsplit("|", 1)
            uam, uag = flags[0] == "1", flags[1] == "1"
            if isinstance(want, str):
                with pytest.raises(Exception):
                    R.OffTargetExpression(pat, uam, uag)(hsp)
            else:
                assert R.OffTargetExpression(pat, uam, uag)(hsp) == want, key
                n += 1
    assert n > 5000
    hsp = _HSP(O.Alignment.from_btop("20", 0, 20, 20))
    for pat, want in d["odd_patterns"].items():
        if isinstance(want, str):
            with pytest.raises(Exception):
                R.OffTargetExpression(pat)(hsp)
        else:
            assert R.OffTargetExpression(pat)(hsp) == want, pat


def test_criteria_lowering():
    c = R.make_criteria(ot_mismatch=5, ot_gap=2, ot_pattern="(0mg5,1mg6-)|2m-4--2", ot_pamless=False)
    assert (c.max_mismatch, c.max_gap, c.pamless, c.n_units, c.n_ops) == (4, 1, 0, 3, 5)
    assert [c.ops[i] for i in range(5)] == [0, 1, -1, 2, -2]
    assert (c.units[2].start, c.units[2].end, c.units[2].rvs, c.units[2].n) == (-2, -3, 1, 2)
    assert c.units[1].start == 5 and c.units[1].end == -(2 ** 31)
    assert R.make_criteria(0, 0).max_mismatch == 0 and R.make_criteria(1, 1).max_gap == 0


def _random_alignment(rng, L, max_gap):
    """kinds string + (qs, qe) of a random member of the alignment universe."""
    qs = rng.choice([0, 0, 0, 1, 2, 3])
    qe = L - rng.choice([0, 0, 0, 1, 2, 3])
    kinds, gaps, j = [], 0, qs
    while j < qe:
        first, last = j == qs, j == qe - 1
        r = rng.random()
        if not first and gaps < max_gap and r < 0.05:
            kinds.append("d"); gaps += 1
            continue
        if not first and not last and gaps < max_gap and r < 0.10:
            kinds.append("i"); gaps += 1; j += 1
            continue
        kinds.append("m" if r > 0.85 else "."); j += 1
    return "".join(kinds), qs, qe


def test_prefilter_bounds_are_sound():
    """edit_bound / exact_pieces must be NECESSARY conditions of the pattern (the device prefilters rely on it):
    checked by brute force on random alignments evaluated with the oracle's tuple semantics."""
    import random
    rng = random.Random(99)
    L = 20
    pats = ["0mg-10,1mg-11-", "1mg1-", "2mg1-", "0mg5,1mg6-", "(0mg5,1mg6-)|(0mg6,1m7-)", "1m10,1m11-", "1m1-,1g1-",
            "0mg-12|0mg12", "1mg-14", "0m-9,3mg1-", "2m1-", "0mg8,0mg-8"]
    checked = 0
    for pat in pats:
        for gcap in (0, 1):
            e = R.OffTargetExpression(pat)
            bound = e.edit_bound(L, gcap)
            pieces = e.exact_pieces(L, gcap)
            for _ in range(1500):
                kinds, qs, qe = _random_alignment(rng, L, gcap)
                aln = O.Alignment(kinds, qs, qe, L)
                if not e(_HSP(aln)):
                    continue
                checked += 1
                cost = aln.mismatches + aln.gaps + (L - (qe - qs))
                if bound >= 0:
                    assert cost <= bound, (pat, gcap, kinds, qs, qe)
                if pieces:
                    # per query position: '.', 'm', 'i' or ' ' and whether a deletion touches it
                    per, dele, j = [" "] * L, [False] * (L + 1), qs
                    for ch in kinds:
                        if ch == "d":
                            dele[j] = True
                        else:
                            per[j] = ch
                            j += 1
                    ok = any(all(per[x] == "." for x in range(a, a + n)) and not any(dele[x] for x in range(a + 1, a + n))
                             for a, n in pieces)
                    assert ok, (pat, gcap, kinds, qs, qe, pieces)
    assert checked > 1000


def _hits_from_oracle(entries):
    from minorg_b200.grna import Target, gRNAHit, gRNAHits
    h, targets = gRNAHits(), {}
    for seq, hs in entries:
        for (tid, a, b, strand, hid, tlen) in hs:
            t = targets.setdefault((tid, tlen), Target("N" * tlen, id=tid, strand="+"))
            h.add_hit(seq, gRNAHit(t, a, b, strand, hid))
    return h


def test_grna_writers_match_reference_text(golden, tmp_path):
    """write_fasta / write_mapping (v5) byte-for-byte against files written by the reference's gRNAHits."""
    import os
    gdir = os.path.join(os.path.dirname(__file__), "golden")
    n = 0
    for r in golden("enumerate.json")["rows"]:
        if "map_text" not in r:
            continue
        h = _hits_from_oracle(O.find_multi_grna(os.path.join(gdir, r["fasta"]), r["pam"], r["L"]))
        fa, mp = str(tmp_path / "g.fa"), str(tmp_path / "g.map")
        h.write_fasta(fa, write_all=True)
        h.write_mapping(mp, write_all=True, write_checks=True)
        assert open(fa).read() == r["fasta_text"] and open(mp).read() == r["map_text"]
        # resume: read the .map back, set a check, write again
        from minorg_b200.grna import gRNAHits
        back = gRNAHits()
        back.parse_from_mapping(mp)
        assert back.seqs == h.seqs and [s.id for s in back.flatten_gRNAseqs()] == [s.id for s in h.flatten_gRNAseqs()]
        mp2 = str(tmp_path / "g2.map")
        back.write_mapping(mp2, write_all=True, write_checks=True)
        assert open(mp2).read() == r["map_text"]
        back.set_seqs_check("background", False, back.seqs[:1])
        assert back.get_gRNAseq_by_seq(back.seqs[0]).check("background", mode="str") == "fail"
        assert back.filter_seqs("background", accept_invalid=True) == back.seqs[1:]
        n += 1
    assert n >= 10


def test_minimum_sets_match_reference(golden, tmp_path):
    """Default minimum-set path (LAR + tie-breakers) vs sets chosen by the reference's make_get_minimum_set."""
    import os
    from minorg_b200.minimum_set import minimum_sets
    gdir = os.path.join(os.path.dirname(__file__), "golden")
    rows = golden("minimumset.json")["rows"]
    assert len(rows) >= 16 and any(len(r["sets"][0]) >= 4 for r in rows)
    for r in rows:
        assert r["error"] is None
        h = _hits_from_oracle(O.find_multi_grna(os.path.join(gdir, r["fasta"]), r["pam"], r["L"]))
        h.assign_seqid()
        h.set_seqs_check("background", True, [s for s in h.seqs if s not in r["failed"]])
        h.set_seqs_check("background", False, r["failed"])
        valid = h.filter_seqs("background")
        got = minimum_sets(h, valid, r["n_sets"])
        assert [sorted(s) for s in got] == r["sets"], (r["fasta"], r["pam"], r["n_sets"])
        mp = str(tmp_path / "final.map")
        h.write_mapping(mp, sets=got)
        assert open(mp).read() == r["map_text"]


def test_eqv_matches_reference_up_to_its_hash_order(golden, tmp_path):
    """.eqv rows, group names and memberships equal the reference's; group order is by descending coverage (the
    reference's order among equal-coverage groups depends on PYTHONHASHSEED, see grna.write_equivalents)."""
    import os
    gdir = os.path.join(os.path.dirname(__file__), "golden")
    for r in golden("eqv.json")["rows"]:
        h = _hits_from_oracle(O.find_multi_grna(os.path.join(gdir, r["fasta"]), r["pam"], r["L"]))
        fo = str(tmp_path / "g.eqv")
        h.write_equivalents(fo, write_all=True)
        got, want = open(fo).read().splitlines(), r["eqv_text"].splitlines()
        assert got[0] == want[0] and sorted(got[1:]) == sorted(want[1:]), (r["fasta"], r["pam"])
        cov = [int(l.split("\t")[1]) for l in got[1:]]
        assert cov == sorted(cov, reverse=True)
        for lines in (got, want):          # members of a group are contiguous and ordered by relative position
            seen = []
            for l in lines[1:]:
                f = l.split("\t")
                if not seen or seen[-1][0] != f[0]:
                    assert f[0] not in [x[0] for x in seen]
                    seen.append((f[0], float(f[4])))
                else:
                    assert float(f[4]) >= seen[-1][1]
                    seen[-1] = (f[0], float(f[4]))


def _emulate_enumerator(program, seq, L):
    """What K2 (csrc/enumerate.cu) computes from a PAM program, restated in Python: starts s such that some
    5' alternative ends at s and some 3' alternative starts at s+L (byte classes; empty side = no constraint)."""
    data = seq.encode("latin-1")

    def side_ok(side, at, behind):
        if side.n_alts == 0:
            return True
        for a in range(side.n_alts):
            n = side.len[a]
            s = at - n if behind else at
            if s < 0 or s + n > len(data):
                continue
            if all((side.allow[a][k][data[s + k] >> 5] >> (data[s + k] & 31)) & 1 for k in range(n)):
                return True
        return False
    return [s for s in range(len(data) - L + 1) if side_ok(program.five, s, True) and side_ok(program.three, s + L, False)]


def test_pam_program_equals_regex_on_random_input():
    """The device representation of a PAM (finite alternatives of byte classes) selects exactly the starts the
    reference's regex does (`regex.finditer(..., overlapped=True)`, IGNORECASE), for presets, sets and {m,n} repeats."""
    import random
    import regex
    rng = random.Random(8)
    pams = [".NGG", "TTTV.", "NGG", "GG", "TTN", ".NGRRT", "R{1,2}T", "T{3}V.", "[AG]GG", ".N[AG]{2}G", "Y{0,2}A.", "DTTN.",
            ".NNNNRYAC", ".", "A{2,3}.", ".W{1,3}G", "BGG."]
    for pam in pams:
        for L in (5, 20):
            p = PAM(pam, L)
            prog = p.program()
            rx = regex.compile(p.regex_string(), regex.IGNORECASE)
            for _ in range(6):
                seq = "".join(rng.choice("ACGTacgtNRYn") for _ in range(rng.randrange(0, 90)))
                want = [m.start() for m in rx.finditer(seq, overlapped=True)]
                assert _emulate_enumerator(prog, seq, L) == want, (pam, L, seq)


def test_generated_threshold_networks_exhaustively():
    """csrc/csa_gen.cuh: every emitted count_le<T,K> network (T <= 12) is simulated on ALL 2^T inputs (big-int truth
    vectors through the emitted LOP3 look-up tables) and must equal popcount(inputs) <= K."""
    import os
    import re
    src = open(os.path.join(os.path.dirname(os.path.dirname(__file__)), "minorg_b200", "csrc", "csa_gen.cuh")).read()
    funcs = re.findall(r"uint32_t count_le<(\d+), (\d+)>\(const uint32_t \(&x\)\[\d+\]\) \{\n(.*?)\n\}", src, flags=re.S)
    assert len(funcs) >= 90
    checked = 0
    for T, K, body in funcs:
        T, K = int(T), int(K)
        if T > 12:
            continue
        n = 1 << T
        full = (1 << n) - 1
        env = {"0u": 0, "0xFFFFFFFFu": full}
        for i in range(T):
            v = 0
            for a in range(n):
                if (a >> i) & 1:
                    v |= 1 << a
            env["x[%d]" % i] = v
        result = None
        for ln in body.splitlines():
            m = re.match(r"\s+const uint32_t (\w+) = lop3<0x(\w+)>\((.+), (.+), (.+)\);", ln)
            if m:
                tt, a, b, c = int(m.group(2), 16), env[m.group(3)], env[m.group(4)], env[m.group(5)]
                r = 0
                for idx in range(8):
                    if (tt >> idx) & 1:
                        r |= (a if idx & 4 else ~a & full) & (b if idx & 2 else ~b & full) & (c if idx & 1 else ~c & full)
                env[m.group(1)] = r
            else:
                result = env[re.match(r"\s+return (.+);", ln).group(1)]
        want = 0
        for a in range(n):
            if bin(a).count("1") <= K:
                want |= 1 << a
        assert result == want, (T, K)
        checked += 1
    assert checked >= 60


def test_edit_distance_rows_arithmetic_identities():
    """The gapped prefilter (csrc/screen_gapped.cu) computes two of Hyyro's five per-cell words as carry-free integer
    arithmetic (IMAD on the FMA pipe) instead of LOP3.  Check, on whole 32-bit words and over every reachable per-bit
    state, that the arithmetic column update equals the boolean one and that both equal the plain edit-distance DP."""
    import numpy as np
    M32 = 0xFFFFFFFF
    rng = np.random.default_rng(5)
    L, n_chars = 9, 400
    pattern = rng.integers(0, 4, size=(32, L))                   # 32 queries, one per bit
    text = rng.integers(0, 5, size=n_chars)                      # 4 = invalid base (mismatches everything)
    Pv = [M32] * L; Mv = [0] * L                                 # boolean formulation
    Pa = [M32] * L; Ma = [0] * L; X = [1] * L                    # arithmetic formulation: X = Mv - Pv (mod 2^32)
    col = np.tile(np.arange(1, L + 1), (32, 1))                  # plain DP column D[1..L][j] per query
    for c in text:
        mis = [sum(1 << k for k in range(32) if c == 4 or pattern[k, i] != c) for i in range(L)]
        Ph = Mh = 0
        Pha = Mha = Dh = 0
        for i in range(L):
            nD0 = mis[i] & ~Mv[i] & ~Mh & M32
            Mh_out = Pv[i] & ~nD0 & M32
            Ph_out = (Mv[i] | (nD0 & ~Pv[i])) & M32
            Pv_new = (Mh | (nD0 & ~Ph)) & M32
            Mv_new = Ph & ~nD0 & M32
            Pv[i], Mv[i], Ph, Mh = Pv_new, Mv_new, Ph_out, Mh_out
            nD0a = mis[i] & ~Ma[i] & ~Mha & M32
            Mha_out = Pa[i] & ~nD0a & M32
            Ma_new = Pha & ~nD0a & M32
            Dh_out = (nD0a + X[i]) & M32
            Pha_out = (Mha_out + Dh_out) & M32
            X[i] = (Dh - nD0a) & M32
            Pa[i] = (Ma_new - X[i]) & M32
            Ma[i] = Ma_new
            Dh, Pha, Mha = Dh_out, Pha_out, Mha_out
            assert (Pa[i], Ma[i], Pha, Mha) == (Pv[i], Mv[i], Ph, Mh)
            assert X[i] == (Ma[i] - Pa[i]) & M32 and Dh == (Pha - Mha) & M32
            assert Pv[i] & Mv[i] == 0 and Ph & Mh == 0
        new = np.empty_like(col)                                  # reference DP, D[0][j] = 0
        for k in range(32):
            up_left, up = 0, 0
            for i in range(L):
                cost = 0 if (c != 4 and pattern[k, i] == c) else 1
                v = min(up_left + cost, up + 1, col[k, i] + 1)
                up_left, up = col[k, i], v
                new[k, i] = v
        for i in range(L):                                        # vertical deltas of the new column
            for k in range(32):
                d = new[k, i] - (new[k, i - 1] if i else 0)
                assert ((Pv[i] >> k) & 1, (Mv[i] >> k) & 1) == (int(d == 1), int(d == -1))
        col = new


def test_native_fasta_reader_matches_python_reader(tmp_path):
    """mg_fasta_read (host threads inside libminorg_b200.so; no GPU involved) against the numpy reader, which follows the
    record semantics of the reference's FASTA handling: ids, offsets and sequence bytes, on awkward files."""
    import os
    import random
    import numpy as np
    from minorg_b200 import _lib
    from minorg_b200.fasta import read_fasta_arrays
    here = os.path.dirname(__file__)
    rng = random.Random(9)
    body = "".join(rng.choice("ACGTacgtNn") for _ in range(5000))
    big = "".join(rng.choice("ACGT") for _ in range(9_500_000))           # several 4 MB work items, one record
    cases = {
        "plain": ">a desc\nACGT\nAC\n>b\nTTTT\n",
        "no_final_newline": ">a\nACGT\n>b\nGG",
        "crlf": ">a x\r\nACGT\r\nAC\r\n>b\r\nT T\tT\r\n",
        "junk_before_first_header": "garbage\nlines\n>a\nAC\n",
        "empty_records": ">a\n>b\n\n\n>c\nA\n>d",
        "header_only": ">only",
        "no_header": "ACGT\nACGT\n",
        "empty": "",
        "gt_inside_line": ">a\nAC>GT\nA\n>b\n>",
        "blank_and_spaces": ">a\n\n  AC GT \n\t\nNN\n",
        "header_tabs": ">\tid1\tdescription here\nACGT\n> id2 \nGG\n",
        "long_lines": ">x\n" + body + "\n>y\n" + "\n".join(body[i:i + 61] for i in range(0, len(body), 61)) + "\n",
        "big": ">chrBig\n" + "\n".join(big[i:i + 80] for i in range(0, len(big), 80)) + "\n>tail\nACGT\n",
    }
    for name, text in cases.items():
        p = tmp_path / (name + ".fa")
        p.write_bytes(text.encode("latin-1"))
        want = read_fasta_arrays(str(p))
        for threads in (1, 3, 0):
            got = _lib.read_fasta_native(str(p), threads)
            assert got[2] == want[2], (name, got[2], want[2])
            assert np.array_equal(got[1], want[1]), (name, got[1], want[1])
            n = int(want[1][-1])
            assert np.array_equal(got[0][:n], want[0][:n]), name
    for fixture in ("sample_CDS.fasta", "synth_targets.fasta", "synth_cover.fasta"):
        path = os.path.join(here, "golden", fixture)
        if os.path.exists(path):
            want, got = read_fasta_arrays(path), _lib.read_fasta_native(path)
            assert got[2] == want[2] and np.array_equal(got[1], want[1]) and np.array_equal(got[0], want[0]), fixture
    with pytest.raises(_lib.MinorgB200Error):
        _lib.read_fasta_native(str(tmp_path / "does_not_exist.fa"))
    # gzip (magic 1f 8b) is inflated by both readers; the repo's config-1 subjects ship that way
    import gzip
    from minorg_b200.fasta import read_fasta
    for name in ("crlf", "long_lines", "big"):
        gz = tmp_path / (name + ".fa.gz")
        with gzip.open(gz, "wb") as fh:
            fh.write(cases[name].encode("latin-1"))
        want, got, py = read_fasta_arrays(str(tmp_path / (name + ".fa"))), _lib.read_fasta_native(str(gz)), read_fasta_arrays(str(gz))
        for r in (got, py):
            assert r[2] == want[2] and np.array_equal(r[1], want[1]) and np.array_equal(r[0][:int(want[1][-1])], want[0][:int(want[1][-1])]), name
        assert read_fasta(str(gz)) == read_fasta(str(tmp_path / (name + ".fa")))
    cfg = os.path.join(here, "golden", "config1", "subset_9654.fasta.gz")
    got = _lib.read_fasta_native(cfg)
    assert len(got[2]) == 8 and int(got[1][-1]) == 161377


def test_map_reader_matches_reference_on_its_fixture(golden, tmp_path):
    """gRNAHits.parse_from_mapping against the reference's reader run on its own fixture
    (examples/sample_custom_check.map, an old-format file with a user check column) and on the same rows in the other
    header versions: restored sequences, ids, ranges, checks, and the version-5 text written back."""
    from minorg_b200.grna import gRNAHits
    n = 0
    for row in golden("mapfile.json")["rows"]:
        if "error" in row:                 # version 1 does not parse in the reference either (grna.py:478: NameError)
            continue
        p = tmp_path / (row["name"] + ".map")
        p.write_text(row["map_text"])
        h = gRNAHits()
        h.parse_from_mapping(str(p))
        restored = []
        for seq in h.seqs:
            gs = h.get_gRNAseq_by_seq(seq)
            hits = sorted(h.get_hits(seq), key=lambda x: (x.target_id, x.range[0]))
            restored.append((seq, gs.id, hits))
        assert [r[0] for r in restored] == [r[0] for r in row["restored"]], row["name"]
        for (seq, gid, hits), (wseq, wid, wchecks, whits) in zip(restored, row["restored"]):
            gs = h.get_gRNAseq_by_seq(seq)
            assert gid == wid
            for k, v in wchecks.items():
                assert gs.check(k) == v, (row["name"], seq, k)
            assert len(hits) == len(whits)
            for x, w in zip(hits, whits):
                assert [x.target_id, x.range[0], x.range[1], x.strand, x.hit_id, x.target_len] == w[:6], (row["name"], w)
                for k, v in w[7].items():
                    assert x.check(k) == v, (row["name"], w, k)
        out = tmp_path / (row["name"] + ".out.map")
        h.write_mapping(str(out), write_all=True, write_checks=True)
        assert out.read_text() == row["rewritten"], row["name"]
        n += 1
    assert n == 4



def test_minimum_sets_nr_match_reference(golden):
    """prioritise_nr path (minimum_set_nr.py) vs the reference's make_get_minimum_set(prioritise_nr=True), executed under eight
    hash seeds: exact on the rows where the reference agrees with itself; on the rows where it does not (two gRNA of a
    coverage group tie on position and the reference takes whichever its hash-ordered set yields first) the sets have the
    reference's sizes, cover every target and are mutually exclusive."""
    import os
    from minorg_b200.minimum_set_nr import minimum_sets_nr
    gdir = os.path.join(os.path.dirname(__file__), "golden")
    rows = golden("minimumset_nr.json")["rows"]
    stable = [r for r in rows if r["stable_across_hash_seeds"]]
    assert len(rows) >= 32 and len(stable) >= 29 and any(len(r["sets"][0]) >= 4 for r in stable)
    for r in rows:
        assert r["error"] is None
        h = _hits_from_oracle(O.find_multi_grna(os.path.join(gdir, r["fasta"]), r["pam"], r["L"]))
        h.assign_seqid()
        h.set_seqs_check("background", True, [s for s in h.seqs if s not in r["failed"]])
        h.set_seqs_check("background", False, r["failed"])
        valid = h.filter_seqs("background")
        got = minimum_sets_nr(h, valid, r["n_sets"])
        key = (r["fasta"], r["pam"], r["n_sets"], r["drop"])
        if r["stable_across_hash_seeds"]:
            assert [sorted(s) for s in got] == r["sets"], key
        else:
            assert [len(s) for s in got] == [len(s) for s in r["sets"]], key
        targets = {hit.target_id for hit in h.flatten_hits()}
        used = [s for group in got for s in group]
        assert len(used) == len(set(used)) and set(used) <= set(valid), key
        for group in got:
            assert {hit.target_id for s in group for hit in h.get_hits(s)} >= targets, key


def test_minimum_sets_manual_mode_matches_reference(golden):
    """Manual (interactive) mode of both minimum-set paths vs the reference's make_get_minimum_set(manual_check=True) driven by
    scripted answers (reject the first listed gRNA by sequence, the first of the re-drawn set by id, an invalid entry, 'x';
    the second set accepted at once): the sets SHOWN at every prompt and the final sets are the reference's."""
    import os
    from minorg_b200.minimum_set import minimum_sets
    from minorg_b200.minimum_set_nr import minimum_sets_nr
    gdir = os.path.join(os.path.dirname(__file__), "golden")
    rows = golden("minimumset_manual.json")["rows"]
    assert len(rows) >= 8 and sum(r["stable_across_hash_seeds"] for r in rows) >= 7
    for r in rows:
        assert r["error"] is None and len(r["shown"]) == 5
        h = _hits_from_oracle(O.find_multi_grna(os.path.join(gdir, r["fasta"]), r["pam"], r["L"]))
        h.assign_seqid()
        h.set_seqs_check("background", True, [s for s in h.seqs if s not in r["failed"]])
        h.set_seqs_check("background", False, r["failed"])
        valid = h.filter_seqs("background")
        shown = []

        def review(set_no, listed, _shown=shown):
            _shown.append([set_no, [list(x) for x in listed]])
            k = len(_shown)
            return {1: listed[0][1], 2: listed[0][0], 3: "no such gRNA"}.get(k, "x")
        got = (minimum_sets_nr if r["prioritise_nr"] else minimum_sets)(h, valid, 2, review=review)
        key = (r["fasta"], r["pam"], r["prioritise_nr"])
        if r["stable_across_hash_seeds"]:
            assert shown == r["shown"], key
            assert [sorted(s) for s in got] == r["sets"], key
        else:
            assert [len(s) for s in got] == [len(s) for s in r["sets"]] and len(shown) == 5, key
        rejected = {r["shown"][0][1][0][1], dict((i, s) for i, s in r["shown"][1][1])[r["shown"][1][1][0][0]]}
        if r["stable_across_hash_seeds"]:
            assert not rejected & {s for g in got for s in g}, key
