#!/usr/bin/env python
"""Generate golden vectors by EXECUTING the reference's own Python code in this container.

Run:  python tests/golden/make_golden.py      (needs /root/reference; writes tests/golden/*.json)

What is executed verbatim from /root/reference (with the third-party stand-ins of refshim.py):
  * minorg.pam.PAM                               -> pam.json
  * minorg.ot_regex.ot_range                     -> ot_range.json
  * minorg.blast.BlastHSP + minorg.ot_regex.OffTargetExpression  -> ot_pattern.json
  * minorg.MINORg.MINORg._is_offtarget_aln_nopattern / _has_pam / _is_offtarget_pos (unbound, on a
    stand-in `self`), minorg.filter_grna.Masked   -> criteria.json and screen_*.json
  * minorg.generate_grna.find_multi_gRNA + gRNAHits.write_fasta/write_mapping  -> enumerate.json

screen_*.json: tiny genomes; EVERY local alignment of the E1 universe (contiguous
query sub-range, first/last column a base pair, <= Gcap gap characters strictly inside) is enumerated
here and judged by the reference's three predicates (pos AND aln AND pam, MINORg.py:2105-2121); a gRNA
fails iff any alignment is judged problematic.  Only the enumeration of the universe is ours; every
verdict is the reference's.

The test-suite never imports this file; it only reads the JSON it wrote.
"""
import io
import json
import os
import random
import sys
import contextlib
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
warnings.simplefilter("ignore")
import refshim  # noqa: E402

refshim.install()

with contextlib.redirect_stdout(io.StringIO()):
    import minorg.MINORg as refM  # noqa: E402
    from minorg.pam import PAM as RefPAM  # noqa: E402
    from minorg.ot_regex import ot_range as ref_ot_range, OffTargetExpression as RefOTE  # noqa: E402
    from minorg.blast import BlastHSP as RefBlastHSP  # noqa: E402
    from minorg.filter_grna import Masked as RefMasked  # noqa: E402
    from minorg.generate_grna import find_multi_gRNA as ref_find_multi_gRNA  # noqa: E402
    from minorg import MINORgError  # noqa: E402


def quiet(f, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return f(*a, **k)


class Obj:
    def __init__(self, **k):
        self.__dict__.update(k)


def dump(name, obj):
    path = os.path.join(HERE, name)
    with open(path, "w") as f:
        json.dump(obj, f, indent=None, separators=(",", ":"), sort_keys=True)
        f.write("\n")
    print("wrote", name, os.path.getsize(path), "bytes")


# ----------------------------------------------------------------------------- PAM
def gen_pam():
    pams = ["GG", "NGG", ".NGG", "ngg", ".GG", "TTN", "TTTV.", "DTTN.", ".NGRRT", ".NGRRN", ".NNNNRYAC",
            "T{3}V.", "R{1,2}T", ".", "NG", ".NG", "TTTN.", "ATV.", "YG", "NNGRRT", ".NNAGAAW", "TTV.",
            "N", "G", "tttv.", ".NGA", "BGG.", "NGGNG", "KTTV.", "A.", ".A", "NAA.", "[AG]GG", ".N[AG]G", "TT[ACG]."]
    rows = []
    for pam in pams:
        for L in (20, 23, 1):
            try:
                p = RefPAM(pam=pam, gRNA_length=L)
                row = dict(pam=pam, L=L, parsed=p.pam, regex=quiet(p.regex).pattern,
                           five=p.five_prime(), three=p.three_prime(),
                           five_max=p.five_prime_maxlen(), three_max=p.three_prime_maxlen())
            except Exception as e:  # e.g. more than one '.' -> ValueError on split
                row = dict(pam=pam, L=L, error=type(e).__name__)
            rows.append(row)
    # copy-constructor quirk (pam.py:37-42)
    src = RefPAM(pam="TTTV.", gRNA_length=23)
    cc = [dict(src_L=23, arg_L=a, got_L=RefPAM(pam=src, gRNA_length=a).gRNA_length) for a in (20, 25, None)]
    dump("pam.json", dict(rows=rows, copy_ctor=cc))


# ----------------------------------------------------------------------------- ot_range
def gen_ot_range():
    cases = ["5", "-5", "5-", "-5-", "2-4", "-2--4", "1", "-1", "1-", "-1-", "20", "-20", "10-", "-10-",
             "-11-", "11-", "3-3", "-3--3", "4-2", "-4--2", "0", "-0", "0-3", "2-0", "-2-4", "2--4", "", "a",
             "1-20", "-1--20", "6-", "-6-", "7-15", "-7--15", "--5", "5--", "12", "-12"]
    rows = []
    for c in cases:
        try:
            r = quiet(ref_ot_range, c)
            rows.append(dict(range=c, out=list(r)))
        except MINORgError:
            rows.append(dict(range=c, error=True))
    dump("ot_range.json", dict(rows=rows))


# ----------------------------------------------------------------------------- alignments / patterns
BASES = "ACGT"


def cols_to_btop(cols):
    """cols: list of (kind, qbase, sbase) with kind in '.', 'm', 'i', 'd' -> BLAST BTOP string."""
    out, run = [], 0
    for kind, qb, sb in cols:
        if kind == ".":
            run += 1
            continue
        if run:
            out.append(str(run))
            run = 0
        if kind == "m":
            out.append(qb + sb)
        elif kind == "i":      # base only in the gRNA (query): 'X-'
            out.append(qb + "-")
        else:                  # 'd': base only in the subject: '-Y'
            out.append("-" + sb)
    if run:
        out.append(str(run))
    return "".join(out)


def random_alignment(rng, L, max_gap):
    """Random local alignment of a length-L query: returns (btop, qstart, qend, kinds string)."""
    qs = rng.choice([0, 0, 0, 1, 2, rng.randrange(0, max(1, L // 2))])
    qe = L - rng.choice([0, 0, 0, 1, 2, rng.randrange(0, max(1, L // 2 - 1))])
    if qe - qs < 2:
        qs, qe = 0, L
    n_q = qe - qs
    kinds = []
    gaps = 0
    j = 0
    while j < n_q:
        first, last = (j == 0), (j == n_q - 1)
        r = rng.random()
        if not first and gaps < max_gap and r < 0.08:
            kinds.append("d")          # subject-only column, consumes no query base
            gaps += 1
            continue
        if not first and not last and gaps < max_gap and r < 0.16:
            kinds.append("i")
            gaps += 1
            j += 1
            continue
        kinds.append("m" if r > 0.8 else ".")
        j += 1
    # a 'd' may not be the final column: strip trailing d's
    while kinds and kinds[-1] == "d":
        kinds.pop()
    cols = []
    for k in kinds:
        qb = rng.choice(BASES)
        sb = qb if k == "." else rng.choice([b for b in BASES if b != qb])
        cols.append((k, qb, sb))
    return cols_to_btop(cols), qs, qe, "".join(kinds)


def gen_ot_pattern():
    rng = random.Random(20261019)
    patterns = ["0mg-10,1mg-11-", "1m-10,1mg-11-", "0mg5", "0m5", "1g20", "0g20", "0d-10", "0d10", "0d11-",
                "0i-5", "0i-4", "(0mg5,1mg6-)|(0mg6,1m7-)", "0mg5,1mg6-|0mg6,1m7-", "0m-10|0m10",
                "0g4|(2m4,0g5)", "0g4|(2m4,1g5)", "1m3,2g3", "2mg1-", "3m-1-", "1mi10,0d5-", "0md-8--3",
                "1m2-4|0g-2--4", "((0m5))", "(1m10|0g10),2mg11-", "0m1", "0m-1", "4m1-,1g1-", "2m6-12,0g6-12",
                "1g-15-", "0m-5-", "0mg-20", "5mg20", "1d3-,1i-18"]
    alns = []
    for L in (20, 20, 20, 23, 12):
        for _ in range(14):
            btop, qs, qe, kinds = random_alignment(rng, L, rng.choice([0, 1, 1, 2, 3]))
            alns.append(dict(btop=btop, qstart=qs, qend=qe, qlen=L, kinds=kinds))
    # the reference's own commented example (ot_regex.py:37-40) and SURVEY Appendix D alignments
    alns.append(dict(btop="1AG-T12-TAG1", qstart=1, qend=21, qlen=20, kinds=None))
    for btop, qs, qe in [("2AG14AG2", 0, 20), ("20", 0, 20), ("18", 2, 20), ("5-A15", 0, 20), ("15A-4", 0, 20),
                         ("10AG9", 0, 20), ("10-A10", 0, 20), ("5-A-C15", 0, 20), ("3-A4-C4-G9", 0, 20)]:
        alns.append(dict(btop=btop, qstart=qs, qend=qe, qlen=20, kinds=None))
    rows = []
    for a in alns:
        hsp = Obj(btop=a["btop"], query_start=a["qstart"], query_end=a["qend"])
        q = Obj(seq_len=a["qlen"])
        b = RefBlastHSP(hsp, q, None)
        a["sbtop"] = b.sbtop
        a["tbtop"] = list(b.tbtop)
        a["tbtop_rvs"] = list(b.tbtop_rvs)
        res = {}
        for pat in patterns:
            for uam, uag in ((True, False), (False, False), (True, True), (False, True)):
                key = "%s|%d%d" % (pat, uam, uag)
                try:
                    res[key] = bool(quiet(quiet(RefOTE, pat, unaligned_as_mismatch=uam, unaligned_as_gap=uag), b))
                except Exception as e:
                    res[key] = "error:" + type(e).__name__
        a["results"] = res
        rows.append(a)
    bad = {}
    for pat in ["1gi5", "1gd5", "m5", "1m", "1x5", "1m0", "", "1m5,", "(1m5", "1m5)"]:
        try:
            f = quiet(RefOTE, pat)
            hsp = Obj(btop="20", query_start=0, query_end=20)
            bad[pat] = bool(quiet(f, RefBlastHSP(hsp, Obj(seq_len=20), None)))
        except Exception as e:
            bad[pat] = "error:" + type(e).__name__
    dump("ot_pattern.json", dict(patterns=patterns, rows=rows, odd_patterns=bad))


# ----------------------------------------------------------------------------- criteria (per-HSP)
class ClampSeq(refshim.Seq):
    """Subject sequence whose slices clamp negative starts at the contig start (pyfaidx would wrap;
    SURVEY Appendix B: the oracle clamps to the contig instead)."""

    def __getitem__(self, k):
        if isinstance(k, slice):
            a = 0 if k.start is None else max(0, k.start)
            b = len(self._d) if k.stop is None else max(0, k.stop)
            return ClampSeq(self._d[a:b])
        return self._d[k]

    def reverse_complement(self):
        return ClampSeq(str(refshim.Seq.reverse_complement(self)))


def ref_self(ot_mismatch, ot_gap, pattern, uam, uag, pamless, pam, L, masks, fname="subject.fasta"):
    s = Obj(ot_mismatch=ot_mismatch, ot_gap=ot_gap, ot_pamless=pamless, pam=RefPAM(pam=pam, gRNA_length=L))
    if pattern is None:
        s._ot_function = lambda hsp, qr: refM.MINORg._is_offtarget_aln_nopattern(s, hsp, qr)
    else:
        s.ot_pattern = quiet(RefOTE, pattern, unaligned_as_mismatch=uam, unaligned_as_gap=uag)
        s._ot_function = lambda hsp, qr: refM.MINORg._is_offtarget_aln_pattern(s, hsp, qr)
    s.masked = {fname: tuple(RefMasked(Obj(query_id=m[0], hit_id=m[1], hit_start=m[2], hit_end=m[3])) for m in masks)}
    return s


def gen_criteria():
    rng = random.Random(7)
    rows = []
    for _ in range(400):
        L = rng.choice([20, 20, 23])
        om, og = rng.choice([0, 1, 2, 3, 5]), rng.choice([0, 1, 2, 3])
        qs = rng.choice([0, 0, 1, 2, 3])
        qe = L - rng.choice([0, 0, 1, 2, 3])
        hsp = Obj(query_start=qs, query_end=qe, mismatch_num=rng.randrange(0, 6), gap_num=rng.randrange(0, 3))
        s = Obj(ot_mismatch=om, ot_gap=og)
        out = refM.MINORg._is_offtarget_aln_nopattern(s, hsp, Obj(seq_len=L))
        rows.append(dict(L=L, ot_mismatch=om, ot_gap=og, qstart=qs, qend=qe, mm=hsp.mismatch_num,
                         gaps=hsp.gap_num, problematic=bool(out)))
    within = []
    for _ in range(200):
        ms, me = sorted(rng.sample(range(0, 60), 2))
        hs = rng.randrange(0, 60)
        he = hs + rng.randrange(1, 25)
        m = RefMasked(Obj(query_id="t", hit_id="c", hit_start=ms, hit_end=me))
        within.append(dict(mask=[ms, me], hit=[hs, he], within=bool(m.within((hs, he)))))
    dump("criteria.json", dict(nopattern=rows, within=within))


# ----------------------------------------------------------------------------- exhaustive screens
COMP = str.maketrans("ACGTNacgtn", "TGCANtgcan")


def rc(s):
    return s.translate(COMP)[::-1]


def enum_alignments(q, T, gcap):
    """Yield every alignment of the E1 universe of query q against text T (one strand):
    (qs, qe, s, t, cols) with cols a list of (kind, qbase, sbase)."""
    L, n = len(q), len(T)
    for qs in range(L):
        for s in range(n):
            # iterative DFS; state = (j, t, gaps, cols)
            stack = [(qs, s, 0, [])]
            while stack:
                j, t, g, cols = stack.pop()
                # pair column
                if j < L and t < n:
                    a, b = q[j].upper(), T[t].upper()
                    same = (a == b) and (a in BASES)
                    c2 = cols + [("." if same else "m", q[j], T[t])]
                    yield (qs, j + 1, s, t + 1, c2)
                    stack.append((j + 1, t + 1, g, c2))
                if cols and g < gcap:
                    if j < L:
                        stack.append((j + 1, t, g + 1, cols + [("i", q[j], "-")]))
                    if t < n:
                        stack.append((j, t + 1, g + 1, cols + [("d", "-", T[t])]))


def screen_with_reference(grnas, contigs, masks, crit):
    """-> (fail flags, per-gRNA count of problematic alignments)."""
    L = len(grnas[0])
    gcap = max(1, crit["ot_gap"]) - 1
    fname = "subject.fasta"
    s = ref_self(crit["ot_mismatch"], crit["ot_gap"], crit.get("pattern"), crit.get("uam", True),
                 crit.get("uag", False), crit["pamless"], crit["pam"], L, masks, fname)
    ifasta = Obj(filename=fname)
    seqs = {cid: ClampSeq(seq) for cid, seq in contigs}
    ifasta.__class__ = type("IF", (Obj,), {"__getitem__": lambda self, k: seqs[k]})
    fails, counts = [], []
    for gi, q in enumerate(grnas):
        qr = Obj(seq_len=L)
        nprob = 0
        for cid, seq in contigs:
            n = len(seq)
            for strand, T in ((1, seq), (-1, rc(seq))):
                for qs, qe, a, b, cols in enum_alignments(q, T, gcap):
                    hs, he = (a, b) if strand == 1 else (n - b, n - a)
                    hsp = Obj(query_id="g%d" % gi, hit_id=cid, hit_start=hs, hit_end=he, hit_strand=strand,
                              query_start=qs, query_end=qe, btop=cols_to_btop(cols),
                              mismatch_num=sum(1 for c in cols if c[0] == "m"),
                              gap_num=sum(1 for c in cols if c[0] in "id"))
                    if not refM.MINORg._is_offtarget_pos(s, hsp, ifasta):
                        continue
                    if not quiet(s._ot_function, hsp, qr):
                        continue
                    if not (s.ot_pamless or refM.MINORg._has_pam(s, hsp, qr, ifasta)):
                        continue
                    nprob += 1
        fails.append(nprob > 0)
        counts.append(nprob)
    return fails, counts


def mutate(rng, s, nmm):
    s = list(s)
    for p in rng.sample(range(len(s)), nmm):
        s[p] = rng.choice([b for b in BASES if b != s[p]])
    return "".join(s)


def make_case(rng, name, L, n_contigs, clen, crit, n_grna, with_n=False, with_masks=True):
    contigs = []
    for c in range(n_contigs):
        n = rng.randrange(max(4, clen // 2), clen + 1) if c else clen
        contigs.append(["ctg%d" % c, "".join(rng.choice(BASES) for _ in range(n))])
    if n_contigs > 2:
        contigs[-1][1] = contigs[-1][1][: max(3, L // 2)]      # a contig shorter than the gRNA
    grnas = []
    M = max(1, crit["ot_mismatch"]) - 1
    gcap = max(1, crit["ot_gap"]) - 1
    for i in range(n_grna):
        cid = rng.randrange(len(contigs))
        seq = contigs[cid][1]
        kind = i % 6
        if len(seq) < L + 2 or kind == 5:
            grnas.append("".join(rng.choice(BASES) for _ in range(L)))
            continue
        p = rng.randrange(0, len(seq) - L - 1)
        g = seq[p:p + L]
        if kind == 1:
            g = mutate(rng, g, M)
        elif kind == 2:
            g = mutate(rng, g, M + 1)
        elif kind == 3 and gcap:
            k = rng.randrange(2, L - 2)
            g = (g[:k] + g[k + 1:] + rng.choice(BASES)) if rng.random() < .5 else (g[:k] + rng.choice(BASES) + g[k:])[:L]
        elif kind == 4:           # hangs off a contig end
            g = (seq[:L - 2] if rng.random() < .5 else seq[-(L - 2):])
            g = ("".join(rng.choice(BASES) for _ in range(L - len(g))) + g) if rng.random() < .5 else \
                (g + "".join(rng.choice(BASES) for _ in range(L - len(g))))
        if rng.random() < 0.5:
            g = rc(g)
        grnas.append(g)
    if with_n:
        cid, seq = contigs[0]
        p = rng.randrange(5, len(seq) - 8)
        contigs[0][1] = seq[:p] + "NNN" + seq[p + 3:]
        grnas[0] = grnas[0][:3] + "N" + grnas[0][4:]
    masks = []
    if with_masks:
        for i in range(rng.randrange(1, 4)):
            cid, seq = rng.choice(contigs)
            if len(seq) < 12:
                continue
            a = rng.randrange(0, len(seq) - 8)
            b = min(len(seq), a + rng.randrange(L - 4, L + 14))
            masks.append(["tgt%d" % i, cid, a, b])
    fails, counts = screen_with_reference(grnas, [tuple(c) for c in contigs], masks, crit)
    print("  case", name, "fails", sum(fails), "/", len(fails))
    return dict(name=name, L=L, contigs=contigs, grnas=grnas, masks=masks, criteria=crit, fail=fails,
                n_problematic_alignments=counts)


def gen_screens():
    rng = random.Random(424242)
    cases = []
    base = dict(pattern=None, uam=True, uag=False, pamless=True, pam=".NGG")
    # ungapped, PAM-less, no pattern (reduction (2) of SURVEY 8c): the fast path
    for i, (om, L) in enumerate([(1, 20), (2, 20), (3, 20), (5, 20), (0, 20), (2, 23), (3, 12)]):
        cases.append(make_case(rng, "hamming_%d" % i, L, 3, 120, dict(base, ot_mismatch=om, ot_gap=0), 12,
                               with_n=(i % 2 == 0)))
    # ungapped with PAM proximity check
    for i, (om, pam, L) in enumerate([(2, ".NGG", 20), (3, "TTTV.", 23), (2, ".NGRRT", 20), (1, ".", 20), (3, "TTN", 12)]):
        cases.append(make_case(rng, "pam_%d" % i, L, 2, 90, dict(base, ot_mismatch=om, ot_gap=0, pamless=False, pam=pam), 10))
    # ungapped with pattern
    for i, (pat, uam, uag) in enumerate([("0mg-10,1mg-11-", True, False), ("1m-8,2m9-", True, False),
                                         ("0m5|1m-6", True, False), ("1m10,1m11-", False, False),
                                         ("(0mg5,1mg6-)|(0mg6,1m7-)", True, True), ("2m1-", True, False)]):
        cases.append(make_case(rng, "pattern_%d" % i, 20, 2, 80, dict(base, ot_mismatch=1, ot_gap=0, pattern=pat, uam=uam, uag=uag), 10))
    # gapped (Gcap=1), shorter gRNA so the exhaustive enumeration stays cheap
    for i, (om, og, pamless, pam) in enumerate([(1, 2, True, ".NGG"), (2, 2, True, ".NGG"), (3, 2, True, ".NGG"),
                                                (2, 2, False, ".NGG"), (2, 2, False, "TTV.")]):
        cases.append(make_case(rng, "gapped_%d" % i, 12, 2, 44, dict(base, ot_mismatch=om, ot_gap=og, pamless=pamless, pam=pam), 8))
    for i, (pat, uam, uag) in enumerate([("0mg-6,1mg-7-", True, False), ("1mg1-", True, False), ("1m1-,1g1-", True, False),
                                         ("0m6,1g-6", True, True), ("1mi-8,0d1-", True, False)]):
        cases.append(make_case(rng, "gapped_pattern_%d" % i, 12, 2, 40, dict(base, ot_mismatch=1, ot_gap=2, pattern=pat, uam=uam, uag=uag), 8))
    # two gaps: exercises the 'dd' grouping quirk of blast.py:114-116
    for i, (pat, om) in enumerate([(None, 2), ("1m1-,2g1-", 1), ("0d5,2mg6-", 1)]):
        cases.append(make_case(rng, "gap2_%d" % i, 10, 1, 30, dict(base, ot_mismatch=om, ot_gap=3, pattern=pat), 6, with_masks=(i == 0)))
    dump("screen_cases.json", dict(cases=cases))


# ----------------------------------------------------------------------------- enumeration
SYNTH_FASTA = """>recA some description
ACGTTGGCCAGGTTTACCGGAGGNGGCCTTTAACCGGTTAGGCCTTAGGGATCCAATTGGCCGG
ccggaattccggAGGtttaaaccgGGTTTACG
>recB
TTTAGGCCNNNNNGGACCTTTGCGGCCAAGGTTRYGGACGTACGTAGGCATGCATGG

>short
ACGGG
>recA
GGCCAGGTTTACCGGAGGACGTTGGCCAGGTTTACCGGAGGTTTCGGCCAAAGG
>iupac
ACGTRYKMSWBDHVNACGGTTTACGATCGATCGGAGGCCTTTGNGGATTTCCGGAAGG
"""


def hits_dump(h):
    out = []
    for seq in h.seqs:
        hs = sorted(h.get_hits(seq), key=lambda x: x.hit_id)
        out.append([seq, [[x.target_id, x.range[0], x.range[1], x.strand, x.hit_id, x.target_len] for x in hs]])
    return out


def gen_enumerate():
    import tempfile
    synth = os.path.join(HERE, "synth_targets.fasta")
    with open(synth, "w") as f:
        f.write(SYNTH_FASTA)
    rows = []
    srcs = [("sample_CDS.fasta", "/root/reference/examples/sample_CDS.fasta"), ("synth_targets.fasta", synth)]
    for label, path in srcs:
        for pam, L in [(".NGG", 20), ("TTTV.", 20), ("ATV.", 20), ("TTTV.", 23), (".", 20), ("NG", 18), (".NGRRT", 21),
                       ("T{3}V.", 20), ("R{1,2}T", 20), ("NGG", 70)]:
            h = quiet(ref_find_multi_gRNA, path, pam=pam, gRNA_len=L)
            row = dict(fasta=label, pam=pam, L=L, n_seqs=len(h), n_hits=sum(len(h.get_hits(s)) for s in h.seqs))
            if label.startswith("synth") or (pam, L) in ((".NGG", 20), ("TTTV.", 23)):
                row["hits"] = hits_dump(h)
                with tempfile.TemporaryDirectory() as td:
                    fa, mp = os.path.join(td, "g.fasta"), os.path.join(td, "g.map")
                    if len(h):
                        quiet(h.write_fasta, fa, write_all=True)
                        quiet(h.write_mapping, mp, write_all=True, write_checks=True)
                        row["fasta_text"] = open(fa).read()
                        row["map_text"] = open(mp).read()
            rows.append(row)
    dump("enumerate.json", dict(rows=rows))


if __name__ == "__main__":
    which = sys.argv[1:] or ["pam", "ot_range", "ot_pattern", "criteria", "enumerate", "screens"]
    if "pam" in which:
        gen_pam()
    if "ot_range" in which:
        gen_ot_range()
    if "ot_pattern" in which:
        gen_ot_pattern()
    if "criteria" in which:
        gen_criteria()
    if "enumerate" in which:
        gen_enumerate()
    if "screens" in which:
        gen_screens()


# ----------------------------------------------------------------------------- minimum set
def gen_minimumset():
    """Runs the reference's default minimum-set path exactly as MINORg.minimumset does (MINORg.py:2571-2626):
    filter_seqs on 'background', deepcopy, make_get_minimum_set(LAR, tie-break all_best_pos -> all_best_nr)."""
    import copy
    import hashlib
    import tempfile
    from minorg.minimum_set import make_get_minimum_set, all_best_nr, all_best_pos
    rows = []
    srcs = [("synth_targets.fasta", ".NGG", 20), ("sample_CDS.fasta", ".NGG", 20), ("sample_CDS.fasta", "TTTV.", 23),
            ("config1/targets.fasta", ".NGG", 20), ("sample_CDS.fasta", "ATV.", 20),
            ("synth_cover.fasta", ".NGG", 20), ("synth_cover.fasta", "TTN", 18), ("synth_cover.fasta", ".NG", 20)]
    for label, pam, L in srcs:
        for n_sets in (1, 4):
            h = quiet(ref_find_multi_gRNA, os.path.join(HERE, label), pam=pam, gRNA_len=L)
            h.assign_seqid()
            failed = [s for s in h.seqs if int(hashlib.md5(s.encode()).hexdigest(), 16) % 5 == 0]
            h.set_seqs_check("background", True, [s for s in h.seqs if s not in failed])
            h.set_seqs_check("background", False, failed)
            valid = copy.deepcopy(quiet(h.filter_seqs, "background", accept_invalid=False, accept_invalid_field=True,
                                        quiet=True, report_invalid_field=True))
            for_cover = copy.deepcopy(valid)
            targets = set(hit.target_id for hit in h.flatten_hits())
            tie = lambda cov, all_cov, covered: tuple(all_best_nr(all_best_pos(cov, all_cov, covered), all_cov, covered).items())[0]
            get = make_get_minimum_set(for_cover, prioritise_nr=False, manual_check=False, targets=targets,
                                       tie_breaker=tie, suppress_warning=True)
            sets, err = [], None
            try:
                while len(sets) < n_sets:
                    s = quiet(get)
                    if not s:
                        break
                    sets.append(sorted(s))
            except Exception as e:
                err = type(e).__name__
            row = dict(fasta=label, pam=pam, L=L, n_sets=n_sets, failed=failed, sets=sets, error=err)
            if sets and err is None:
                with tempfile.TemporaryDirectory() as td:
                    mp = os.path.join(td, "f.map")
                    quiet(valid.write_mapping, mp, sets=sets, version=5)
                    row["map_text"] = open(mp).read()
            rows.append(row)
            print("  minimumset", label, pam, n_sets, "->", [len(s) for s in sets], err)
    dump("minimumset.json", dict(rows=rows))


if __name__ == "__main__" and "minimumset" in sys.argv[1:]:
    gen_minimumset()


def gen_eqv():
    """gRNAHits.write_equivalents (grna.py:1090-1118, minimum_set.py:224-245) on a few enumerations."""
    import tempfile
    rows = []
    for label, pam, L in [("synth_cover.fasta", ".NGG", 20), ("synth_cover.fasta", "TTN", 18), ("sample_CDS.fasta", "TTTV.", 23),
                          ("config1/targets.fasta", ".NGG", 20), ("synth_targets.fasta", ".NGG", 20)]:
        h = quiet(ref_find_multi_gRNA, os.path.join(HERE, label), pam=pam, gRNA_len=L)
        with tempfile.TemporaryDirectory() as td:
            fo = os.path.join(td, "g.eqv")
            quiet(h.write_equivalents, fo, write_all=True)
            rows.append(dict(fasta=label, pam=pam, L=L, eqv_text=open(fo).read()))
    dump("eqv.json", dict(rows=rows))


if __name__ == "__main__" and "eqv" in sys.argv[1:]:
    gen_eqv()


# ----------------------------------------------------------------------------- .map reader
def gen_mapfile():
    """gRNAHits.parse_from_mapping (grna.py:447-514) on the reference's own fixture examples/sample_custom_check.map
    (format version 2: 'group' column, no target length, a user check column) and on the same rows rewritten in the
    other header versions (1: '[start, end)' range; 3/5: with target length; 4: 'set' without length).  Dumped: what the
    reader restored, and the version-5 text the reference's writer produces from it."""
    import tempfile
    from minorg.grna import gRNAHits
    src = open("/root/reference/examples/sample_custom_check.map").read().splitlines()
    header, data = src[0].split("\t"), [ln.split("\t") for ln in src[1:] if ln.strip()]
    gi = header.index("group")
    variants = {"v2_fixture": "\n".join(src) + "\n"}
    h1 = header[:5] + ["range (0-index, end exclusive)"] + header[7:]
    variants["v1"] = "\n".join(["\t".join(h1)] + ["\t".join(r[:5] + ["[%d, %d)" % (int(r[5]) - 1, int(r[6]))] + r[7:]) for r in data]) + "\n"
    h3 = header[:3] + ["target length"] + header[3:]
    variants["v3"] = "\n".join(["\t".join(h3)] + ["\t".join(r[:3] + ["500"] + r[3:]) for r in data]) + "\n"
    h4 = [("set" if x == "group" else x) for x in header]
    variants["v4"] = "\n".join(["\t".join(h4)] + ["\t".join(r) for r in data]) + "\n"
    h5 = [("set" if x == "group" else x) for x in h3]
    variants["v5"] = "\n".join(["\t".join(h5)] + ["\t".join(r[:3] + ["500"] + r[3:]) for r in data]) + "\n"
    assert gi == 7
    rows = []
    with tempfile.TemporaryDirectory() as td:
        for name, text in variants.items():
            p = os.path.join(td, name + ".map")
            with open(p, "w") as f:
                f.write(text)
            h = gRNAHits()
            try:
                quiet(h.parse_from_mapping, p)
            except Exception as e:             # version 1 hits `re` without an import in grna.py:478 -> NameError
                rows.append(dict(name=name, map_text=text, error=type(e).__name__))
                continue
            restored = []
            for seq in h.seqs:
                gs = h.get_gRNAseq_by_seq(seq)
                hits = sorted(h.get_hits(seq), key=lambda x: (x.target_id, x.range[0]))
                restored.append([seq, gs.id, {k: gs.check(k) for k in sorted(gs._checks)},
                                 [[x.target_id, x.range[0], x.range[1], x.strand, x.hit_id, x.target_len, x.parent_sense(),
                                   {k: x.check(k) for k in sorted(x._checks)}] for x in hits]])
            out = os.path.join(td, name + ".out.map")
            quiet(h.write_mapping, out, write_all=True, write_checks=True)
            rows.append(dict(name=name, map_text=text, restored=restored, rewritten=open(out).read()))
    dump("mapfile.json", dict(rows=rows))


if __name__ == "__main__" and "mapfile" in sys.argv[1:]:
    gen_mapfile()


# ----------------------------------------------------------------------------- minimum set, prioritise_nr
def _nr_rows():
    """The reference's non-redundancy-first minimum-set path exactly as MINORg.minimumset runs it with prioritise_nr=True
    (MINORg.py:2606-2615 -> minimum_set.make_get_minimum_set -> make_set_cover_nr -> minweight_sc)."""
    import copy
    import hashlib
    from minorg.minimum_set import make_get_minimum_set
    rows = []
    srcs = [("synth_targets.fasta", ".NGG", 20), ("sample_CDS.fasta", ".NGG", 20), ("sample_CDS.fasta", "TTTV.", 23),
            ("config1/targets.fasta", ".NGG", 20), ("sample_CDS.fasta", "ATV.", 20),
            ("synth_cover.fasta", ".NGG", 20), ("synth_cover.fasta", "TTN", 18), ("synth_cover.fasta", ".NG", 20)]
    for label, pam, L in srcs:
        for n_sets in (1, 3):
            for drop in (5, 2):
                h = quiet(ref_find_multi_gRNA, os.path.join(HERE, label), pam=pam, gRNA_len=L)
                h.assign_seqid()
                failed = [s for s in h.seqs if int(hashlib.md5(s.encode()).hexdigest(), 16) % drop == 0]
                h.set_seqs_check("background", True, [s for s in h.seqs if s not in failed])
                h.set_seqs_check("background", False, failed)
                valid = copy.deepcopy(quiet(h.filter_seqs, "background", accept_invalid=False, accept_invalid_field=True,
                                            quiet=True, report_invalid_field=True))
                targets = set(hit.target_id for hit in h.flatten_hits())
                sets, err = [], None
                try:
                    get = quiet(make_get_minimum_set, copy.deepcopy(valid), prioritise_nr=True, manual_check=False, targets=targets,
                                suppress_warning=True)
                    while len(sets) < n_sets:
                        s = quiet(get)
                        if not s:
                            break
                        sets.append(sorted(s))
                except Exception as e:
                    err = type(e).__name__
                rows.append(dict(fasta=label, pam=pam, L=L, n_sets=n_sets, drop=drop, failed=failed, sets=sets, error=err))
    return rows


def gen_minimumset_nr():
    """Rows that come out identical under eight hash seeds (the reference iterates Python sets of objects hashed by name in
    several places, minweight_sc.py:127-160): run this file eight times as a subprocess and mark what agrees."""
    import subprocess
    runs = []
    for seed in ("0", "1", "2", "3", "4", "5", "6", "7"):
        env = dict(os.environ, PYTHONHASHSEED=seed)
        out = subprocess.run([sys.executable, os.path.abspath(__file__), "_nr_rows"], env=env, capture_output=True, text=True)
        runs.append(json.loads(out.stdout.strip().splitlines()[-1]))
    rows = []
    for versions in zip(*runs):
        a = versions[0]
        a["stable_across_hash_seeds"] = all(v["sets"] == a["sets"] and v["error"] == a["error"] for v in versions)
        rows.append(a)
        print("  minimumset_nr", a["fasta"], a["pam"], a["n_sets"], a["drop"], "->", [len(s) for s in a["sets"]], a["error"],
              "stable" if a["stable_across_hash_seeds"] else "UNSTABLE")
    dump("minimumset_nr.json", dict(rows=rows))


if __name__ == "__main__" and "_nr_rows" in sys.argv[1:]:
    print(json.dumps(_nr_rows()))
    sys.exit(0)

if __name__ == "__main__" and "minimumset_nr" in sys.argv[1:]:
    gen_minimumset_nr()


# ----------------------------------------------------------------------------- minimum set, manual (interactive) mode
def _manual_rows():
    """make_get_minimum_set(manual_check=True) with scripted answers in place of the terminal (minimum_set.py:43-64, 603-637):
    every set is shown; the script rejects the first listed gRNA by SEQUENCE, then the first listed gRNA of the re-drawn set by
    ID, types something invalid, then accepts with 'x'; the second set is accepted at once."""
    import builtins
    import copy
    import hashlib
    import minorg.minimum_set as ms
    from minorg.minimum_set import make_get_minimum_set, all_best_nr, all_best_pos
    rows = []
    srcs = [("synth_targets.fasta", ".NGG", 20), ("config1/targets.fasta", ".NGG", 20), ("synth_cover.fasta", ".NGG", 20),
            ("synth_cover.fasta", "TTN", 18)]
    for label, pam, L in srcs:
        for nr in (False, True):
            h = quiet(ref_find_multi_gRNA, os.path.join(HERE, label), pam=pam, gRNA_len=L)
            h.assign_seqid()
            failed = [s for s in h.seqs if int(hashlib.md5(s.encode()).hexdigest(), 16) % 5 == 0]
            h.set_seqs_check("background", True, [s for s in h.seqs if s not in failed])
            h.set_seqs_check("background", False, failed)
            valid = copy.deepcopy(quiet(h.filter_seqs, "background", accept_invalid=False, accept_invalid_field=True,
                                        quiet=True, report_invalid_field=True))
            targets = set(hit.target_id for hit in h.flatten_hits())
            shown, state = [], {"n": 0}

            def prompt(grnas, set_num=None, _shown=shown, _state=state):
                listed = [(g.id, str(g.seq)) for g in grnas]
                _shown.append([set_num, listed])
                _state["n"] += 1
                k = _state["n"]
                return {1: listed[0][1], 2: listed[0][0], 3: "no such gRNA"}.get(k, "x")
            real = ms.manual_check_prompt
            ms.manual_check_prompt = prompt
            tie = (lambda *a: tuple()) if nr else \
                (lambda cov, all_cov, covered: tuple(all_best_nr(all_best_pos(cov, all_cov, covered), all_cov, covered).items())[0])
            sets, err = [], None
            try:
                get = quiet(make_get_minimum_set, copy.deepcopy(valid), prioritise_nr=nr, manual_check=True, targets=targets,
                            tie_breaker=tie, suppress_warning=True)
                while len(sets) < 2:
                    s = quiet(get)
                    if not s:
                        break
                    sets.append(sorted(s))
            except Exception as e:
                err = type(e).__name__
            finally:
                ms.manual_check_prompt = real
            rows.append(dict(fasta=label, pam=pam, L=L, prioritise_nr=nr, failed=failed, sets=sets, shown=shown, error=err))
    return rows


def gen_minimumset_manual():
    import subprocess
    runs = []
    for seed in ("0", "1", "2", "3", "4", "5"):
        env = dict(os.environ, PYTHONHASHSEED=seed)
        out = subprocess.run([sys.executable, os.path.abspath(__file__), "_manual_rows"], env=env, capture_output=True, text=True)
        runs.append(json.loads(out.stdout.strip().splitlines()[-1]))
    rows = []
    for versions in zip(*runs):
        a = versions[0]
        a["stable_across_hash_seeds"] = all(v["sets"] == a["sets"] and v["shown"] == a["shown"] and v["error"] == a["error"] for v in versions)
        rows.append(a)
        print("  minimumset_manual", a["fasta"], a["pam"], "nr" if a["prioritise_nr"] else "pos", "->", [len(s) for s in a["sets"]],
              len(a["shown"]), "prompts", a["error"], "stable" if a["stable_across_hash_seeds"] else "UNSTABLE")
    dump("minimumset_manual.json", dict(rows=rows))


if __name__ == "__main__" and "_manual_rows" in sys.argv[1:]:
    print(json.dumps(_manual_rows()))
    sys.exit(0)

if __name__ == "__main__" and "minimumset_manual" in sys.argv[1:]:
    gen_minimumset_manual()
