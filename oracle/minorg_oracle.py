"""CPU ORACLE (test infrastructure, NOT product code).

A plain-Python restatement of the reference's algorithm for the hot path named in BASELINE.json:
PAM-constrained gRNA enumeration and the exhaustive off-target screen with target masking.  Only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import
this file; `minorg_b200/` never does.

Parity pinning: every function below is checked in tests/test_oracle_golden.py against golden vectors
that were produced by EXECUTING the reference's own modules in the authoring container
(tests/golden/make_golden.py: minorg.pam.PAM, minorg.ot_regex, minorg.blast.BlastHSP,
MINORg._is_offtarget_aln_nopattern/_has_pam/_is_offtarget_pos, filter_grna.Masked,
generate_grna.find_multi_gRNA, gRNAHits.write_fasta/write_mapping), and against the known answers the
reference documents (docs/source/tutorial_py.rst:517-531: 201/95/267 gRNA).  The part of the screen the
reference delegates to BLASTn (which alignments exist) has no reference test; it is replaced by the
exhaustive alignment universe "E1" of SURVEY.md section 8c, and that replacement is what "exhaustive"
means in BASELINE.json's metric.

Each function cites the reference file:line it follows (paths under /root/reference/minorg/).
"""
import regex as re  # the reference uses the `regex` module under this name (pam.py:1)

# ------------------------------------------------------------------ third-party behaviour
# Bio.Data.IUPACData.ambiguous_dna_values (Biopython 1.79); letter ORDER inside a set is what PAM echoes.
AMBIGUOUS_DNA = {"A": "A", "C": "C", "G": "G", "T": "T", "M": "AC", "R": "AG", "W": "AT", "S": "CG", "Y": "CT",
                 "K": "GT", "V": "ACG", "H": "ACT", "D": "AGT", "B": "CGT", "X": "GATC", "N": "GATC"}
_RC = str.maketrans("ACGTMRWSYKVHDBNXacgtmrwsykvhdbnx", "TGCAKYWSRMBDHVNXtgcakywsrmbdhvnx")


def reverse_complement(seq):
    """Bio.Seq.reverse_complement: IUPAC-aware, case preserving, other characters unchanged."""
    return seq.translate(_RC)[::-1]


def read_fasta(path):
    """Bio.SeqIO.parse(path, 'fasta'): id = header up to first whitespace, sequence = remaining lines
    joined with whitespace removed, case preserved (generate_grna.py:34-37 consumes .id and .seq)."""
    out, rid, chunks = [], None, []
    with open(path) as f:
        for line in f:
            if line.startswith(">"):
                if rid is not None:
                    out.append((rid, "".join(chunks)))
                toks = line[1:].split()
                rid, chunks = (toks[0] if toks else ""), []
            elif rid is not None:
                chunks.append("".join(line.split()))
    if rid is not None:
        out.append((rid, "".join(chunks)))
    return out


# ------------------------------------------------------------------ PAM (pam.py:26-187)
class OraclePAM:
    def __init__(self, pam="NGG", length=20):
        self.raw = pam
        self.length = length
        self.pam = self._expand(self._infer_full(pam))

    @staticmethod
    def _infer_full(pam):                       # pam.py:55-76
        pam = pam.upper()
        if "." not in pam:
            if "N" not in pam:
                pam = "N" + pam
            pam = ("." + pam) if pam[0] == "N" else (pam + ".")
        return pam

    @staticmethod
    def _expand(pam):                           # pam.py:78-90
        out = ""
        for c in pam:
            m = AMBIGUOUS_DNA.get(c.upper(), c)
            out += m if len(m) == 1 else "[" + m + "]"
        return out

    def regex_string(self):                     # pam.py:103-125
        pre, post = self.pam.split(".")
        s = ""
        if pre:
            s += "(?<=" + pre + ")"
        s += ".{" + str(self.length) + "}"
        if post:
            s += "(?=" + post + ")"
        return s

    def five_prime(self):                       # pam.py:150-158 (parse() is idempotent)
        return self.pam.split(".")[0]

    def three_prime(self):                      # pam.py:159-167
        return self.pam.split(".")[-1]

    @staticmethod
    def pattern_max_len(pattern):               # pam.py:128-147 (including its over-count of {m,n} after a set)
        nsets = len(re.findall(r"(?<=\[).+?(?=\])", pattern))
        no_sets = "".join(re.findall(r"(?<=^|\])[^\]\[]+?(?=$|\[)", pattern))
        reps = sum(int(x.split(",")[-1]) - 1 for x in re.findall(r"(?<=\{).+?(?=\})", no_sets))
        nchars = sum(len(x) for x in re.findall(r"(?<=^|\}).+?(?=$|\{)", no_sets))
        return nchars + nsets + reps

    def five_prime_maxlen(self):
        return self.pattern_max_len(self.five_prime())

    def three_prime_maxlen(self):
        return self.pattern_max_len(self.three_prime())


# ------------------------------------------------------------------ enumeration (generate_grna.py:23-57)
def find_multi_grna(target_fasta, pam="NGG", length=20):
    """-> ordered list [(SEQ, [(target_id, start, end, strand, hit_id, target_len), ...]), ...] in
    first-occurrence order; coordinates are forward-strand, 0-based, end-exclusive (generate_grna.py:48-51)."""
    import regex
    p = OraclePAM(pam, length)
    rx = regex.compile(p.regex_string(), regex.IGNORECASE)
    table, hit_id = {}, 0
    for rid, fwd in read_fasta(target_fasta):
        n = len(fwd)
        for strand, seq in (("+", fwd), ("-", reverse_complement(fwd))):
            for m in rx.finditer(seq, overlapped=True):
                s = m.start()
                g = seq[s:s + length].upper()
                if len(g) != length:
                    continue
                a, b = (s, s + length) if strand == "+" else (n - (s + length), n - s)
                table.setdefault(g, []).append((rid, a, b, strand, hit_id, n))
                hit_id += 1
    return list(table.items())


def grna_ids(n, prefix="gRNA_", zfill=3):
    """grna.py:630-643: ids by dict order, 1-based, zfill(3) (grows past 999)."""
    return ["%s%s" % (prefix, str(i + 1).zfill(zfill)) for i in range(n)]


def grna_fasta_text(entries):
    """grna.py:1061-1089 + fasta.py:32-43 (dict_to_fasta -> Bio.SeqIO.write 'fasta': 60 columns)."""
    out = []
    for gid, (seq, _) in zip(grna_ids(len(entries)), entries):
        out.append(">" + gid)
        out.extend(seq[i:i + 60] for i in range(0, len(seq), 60))
    return "\n".join(out) + "\n" if out else ""


def grna_map_text(entries, checks=("background", "GC", "feature"), status=None):
    """grna.py:925-1060, version 5, write_all=True, write_checks=True: 1-based inclusive start/end,
    rows sorted (stable) by start, target id, gRNA id, set.  status: {(seq, check): 'pass'|'fail'|'NA'}."""
    status = status or {}
    header = ["gRNA id", "gRNA sequence", "target id", "target length", "target sense", "gRNA strand",
              "start", "end", "set"] + list(checks)
    rows = []
    for gid, (seq, hits) in zip(grna_ids(len(entries)), entries):
        for (tid, a, b, strand, _hid, tlen) in hits:
            rows.append([gid, seq, tid, tlen, "NA", strand, a + 1, b, 1] +
                        [status.get((seq, c), "NA") for c in checks])
    rows.sort(key=lambda r: int(r[6]))
    rows.sort(key=lambda r: r[2])
    rows.sort(key=lambda r: r[0])
    rows.sort(key=lambda r: r[8])
    return "\n".join("\t".join(map(str, r)) for r in [header] + rows) + "\n"


# ------------------------------------------------------------------ ot_range (ot_regex.py:63-122)
class OracleError(Exception):
    pass


def ot_range(text):
    try:
        m = re.search(r"^-*\d+", text)
        first = m.group(0)
        neg = first[0] == "-"
        rest = len(text) - len(first)
        start = int(first)
        if start == 0:
            raise OracleError
        if rest == 0:
            return (start, None) if neg else (None, start)
        if rest == 1 and text[-1] == "-":
            if neg:
                return (None, start + 1 if start < -1 else None)
            return (start - 1, None)
        last = re.search(r"(?<=-)-*\d+$", text).group(0)
        if (last[0] == "-") != neg:
            raise OracleError
        end = int(last)
        if end == 0:
            raise OracleError
        if neg:
            return (end, start + 1 if start < -1 else None)
        return (start - 1, end)
    except Exception:
        raise OracleError("cannot parse range %r" % (text,))


# ------------------------------------------------------------------ alignment strings (blast.py:46-172)
class Alignment:
    """kinds: alignment columns of the ALIGNED part only, one char each: '.', 'm', 'i' (base only in the
    gRNA), 'd' (base only in the subject).  qstart/qend/qlen as Biopython reports them."""

    def __init__(self, kinds, qstart, qend, qlen):
        self.kinds, self.qstart, self.qend, self.qlen = kinds, qstart, qend, qlen
        lead, trail = [" "] * qstart, [" "] * (qlen - qend)
        fwd, rvs = [], []
        k = kinds
        i = 0
        while i < len(k):                       # blast.py:102-117, rvs_index=False
            if k[i] != "d":
                fwd.append(k[i])
                i += 1
            else:
                fwd.append(k[i:i + 2])
                i += 2
        for c in k:                             # rvs_index=True: a 'd' joins the previous element
            if c != "d":
                rvs.append(c)
            else:
                rvs[-1] = rvs[-1] + c
        self.tuple_fwd = lead + fwd + trail
        self.tuple_rvs = lead + rvs + trail

    @classmethod
    def from_btop(cls, btop, qstart, qend, qlen):           # blast.py:46-81
        kinds, i = "", 0
        while i < len(btop):
            m = re.match(r"\d+", btop[i:])
            if m:
                kinds += "." * int(m.group(0))
                i += len(m.group(0))
            else:
                a, b = btop[i], btop[i + 1]
                kinds += "d" if a == "-" else ("i" if b == "-" else "m")
                i += 2
        return cls(kinds, qstart, qend, qlen)

    def count(self, chars, start, end, rvs):                # blast.py:122-172
        t = self.tuple_rvs if rvs else self.tuple_fwd
        a = 0 if start is None else start
        b = self.qlen if end is None else end
        step = 1 if b >= a else -1
        s = "".join(t[a:b:step])
        return sum(s.count(c) for c in chars)

    @property
    def mismatches(self):
        return self.kinds.count("m")

    @property
    def gaps(self):
        return self.kinds.count("i") + self.kinds.count("d")


# ------------------------------------------------------------------ ot pattern (ot_regex.py:124-356)
class OffTargetPattern:
    """Strictly left-to-right AND (',') / OR ('|') with parenthesised sub-expressions; True = problematic."""

    def __init__(self, pattern, unaligned_as_mismatch=True, unaligned_as_gap=False):
        self.uam, self.uag = unaligned_as_mismatch, unaligned_as_gap
        self.pattern = pattern
        self.fn, self.length = self._parse(pattern)

    def __call__(self, aln):
        return self.fn(aln)

    def _unit(self, text):                                  # ot_regex.py:222-277
        try:
            n = int(re.search(r"^\d+", text).group(0))
            mm, gap, ins, dele = "m" in text, "g" in text, "i" in text, "d" in text
            if gap and (ins or dele):
                raise OracleError       # reference: AttributeError on self.logfile -> MINORgError (:251,263)
            start, end = ot_range(re.search(r"[-\d]+$", text).group(0))
            rvs = any(x is not None and x < 0 for x in (start, end))
            chars = ("m" if mm else "") + ("i" if ins or gap else "") + ("d" if dele or gap else "")
        except Exception:
            raise OracleError("cannot parse unit %r" % (text,))

        def f(aln):
            num = aln.count(chars, start, end, rvs)
            una = aln.count(" ", start, end, rvs)
            if mm and self.uam:
                num += una
            if (ins or gap) and self.uag:
                num += una
            return num <= n
        return f

    def _parse(self, p):                                    # ot_regex.py:295-356
        fn, op = None, None

        def combine(cur, o, new):
            if cur is None:
                return new
            if o == "and":
                return lambda a: cur(a) and new(a)
            if o == "or":
                return lambda a: cur(a) or new(a)
            raise OracleError("operator missing")
        n, last = len(p), len(p) - 1
        start = i = 0
        while i < n:
            c = p[i]
            if c == "(":
                sub_fn, sub_len = self._parse(p[i + 1:])
                fn = combine(fn, op, sub_fn)
                start = i + 1 + sub_len
                i = start
            elif c in ",|":
                if start != i:
                    fn = combine(fn, op, self._unit(p[start:i]))
                op = "and" if c == "," else "or"
                start = i + 1
                i += 1
            elif c == ")":
                fn = combine(fn, op, self._unit(p[start:i]))
                i += 1
                break
            elif i >= last:
                fn = combine(fn, op, self._unit(p[start:n]))
                i += 1
                break
            else:
                i += 1
        return fn, i


# ------------------------------------------------------------------ per-alignment criteria
def nopattern_problematic(mm, gaps, qstart, qend, qlen, ot_mismatch, ot_gap):
    """MINORg.py:2018-2044."""
    unaligned = qlen - (max(qstart, qend) - min(qstart, qend))
    mq = (max(1, ot_mismatch) - 1) - mm
    gq = (max(1, ot_gap) - 1) - gaps
    return gq >= 0 and mq >= 0 and (mq + gq - unaligned) >= 0


def within_mask(hit, mask):
    """MINORg.py:1994-2000 + filter_grna.py:62-65 + functions.py:310-325; 0-based half-open both."""
    (a, b), (ms, me) = hit, mask
    return min(a, b) >= min(ms, me) and max(a, b) <= max(ms, me)


def has_pam(pam, plus_text, a, b, qstart, qend, qlen, gaps, ot_gap):
    """MINORg.py:2063-2087 restated on the strand the gRNA aligns to: `plus_text` is the contig for a
    plus-strand hit and its reverse complement for a minus-strand hit, [a,b) the aligned span in that
    text.  Windows are clamped at the contig ends (the reference's pyfaidx slices wrap at negative
    starts - undefined behaviour).  Case-insensitive (the genome is 2-bit on the device)."""
    pre_pat, post_pat = pam.five_prime(), pam.three_prime()
    gap_excess = (max(1, ot_gap) - 1) - gaps
    pre_len = pam.five_prime_maxlen() + gap_excess + qstart
    post_len = pam.three_prime_maxlen() + gap_excess + (qlen - qend)
    pre = plus_text[max(0, a - pre_len):a] if pre_len > 0 else ""
    post = plus_text[b:max(b, b + post_len)]
    found_pre = bool(pre_pat) and bool(re.search(pre_pat, pre.upper()))
    found_post = bool(post_pat) and bool(re.search(post_pat, post.upper()))
    return found_pre or found_post


# ------------------------------------------------------------------ masks (filter_grna.py:67-119, fasta.py:45-95)
def find_masks(mask_seqs, contigs):
    """All subject intervals equal to a mask sequence end-to-end on either strand.
    mask_seqs: [(id, seq)], contigs: [(id, seq)] -> set of (mask id, contig id, start, end).
    ACGT sequences compare case-insensitively (BLAST, filter_grna.py:71-73,106-118); a sequence with any
    other character compares case-sensitively, character by character (fasta.py:63-79)."""
    out = set()
    std = set("ACGTUacgtu")
    for mid, mseq in mask_seqs:
        if not mseq:
            continue
        plain = set(mseq) <= std
        for cid, cseq in contigs:
            hay = cseq.upper() if plain else cseq
            for needle in ((mseq.upper(), reverse_complement(mseq).upper()) if plain
                           else (mseq, reverse_complement(mseq))):
                pos = hay.find(needle)
                while pos >= 0:
                    if not plain or set(hay[pos:pos + len(needle)]) <= std:
                        out.add((mid, cid, pos, pos + len(needle)))
                    pos = hay.find(needle, pos + 1)
    return out


# ------------------------------------------------------------------ exhaustive screen, universe E1
_VALID = set("ACGT")


def _pairs(q, T):
    qq, TT = q.upper(), T.upper()
    return qq, TT


def _alignments(q, T, gcap):
    """Every local alignment of q (gRNA, 5'->3') against text T: contiguous query range [qs,qe), first and
    last column a base pair, at most gcap gap characters strictly inside.  Recursive generator of
    (qs, qe, s, t, kinds)."""
    L, n = len(q), len(T)

    def extend(j, t, g, kinds):
        if j < L and t < n:
            same = q[j] == T[t] and q[j] in _VALID
            k2 = kinds + ("." if same else "m")
            yield (j + 1, t + 1, k2)
            yield from extend(j + 1, t + 1, g, k2)
        if kinds and g < gcap:
            if j < L:
                yield from extend(j + 1, t, g + 1, kinds + "i")
            if t < n:
                yield from extend(j, t + 1, g + 1, kinds + "d")
    for qs in range(L):
        for s in range(n):
            for qe, t, kinds in extend(qs, s, 0, ""):
                yield qs, qe, s, t, kinds


class Criteria:
    """The reference's attributes that decide a hit (MINORg.py:686-707; defaults parse_config.py:493-513)."""

    def __init__(self, ot_mismatch=1, ot_gap=0, pattern=None, uam=True, uag=False, pamless=True, pam=".NGG",
                 length=20):
        self.ot_mismatch, self.ot_gap, self.pamless = ot_mismatch, ot_gap, pamless
        self.pattern = OffTargetPattern(pattern, uam, uag) if pattern else None
        self.pam = OraclePAM(pam, length)
        self.gcap = max(1, ot_gap) - 1


def alignment_is_problematic(crit, aln, plus_text, a, b):
    """is_offtarget_aln AND is_offtarget_pam (MINORg.py:2105-2121) for one alignment; position check is
    done by the caller in forward coordinates."""
    if crit.pattern is not None:
        if not crit.pattern(aln):                                   # MINORg.py:2015-2016
            return False
    elif not nopattern_problematic(aln.mismatches, aln.gaps, aln.qstart, aln.qend, aln.qlen,
                                   crit.ot_mismatch, crit.ot_gap):
        return False
    if crit.pamless:                                                # MINORg.py:2101
        return True
    return has_pam(crit.pam, plus_text, a, b, aln.qstart, aln.qend, aln.qlen, aln.gaps, crit.ot_gap)


def screen_exhaustive(grnas, contigs, masks, crit):
    """-> (fail flags, #problematic alignments) per gRNA.  masks: iterable of (mask id, contig id, s, e).
    Tiny inputs only (pure-Python enumeration of E1)."""
    by_contig = {}
    for _mid, cid, ms, me in masks:
        by_contig.setdefault(cid, []).append((ms, me))
    fails, counts = [], []
    for q in grnas:
        qq = q.upper()
        nprob = 0
        for cid, seq in contigs:
            n = len(seq)
            up = seq.upper()
            for strand, T in ((1, up), (-1, reverse_complement(up))):
                for qs, qe, s, t, kinds in _alignments(qq, T, crit.gcap):
                    hs, he = (s, t) if strand == 1 else (n - t, n - s)
                    if any(within_mask((hs, he), m) for m in by_contig.get(cid, ())):
                        continue                                    # on-target (MINORg.py:1994-2000)
                    if alignment_is_problematic(crit, Alignment(kinds, qs, qe, len(qq)), T, s, t):
                        nprob += 1
        fails.append(nprob > 0)
        counts.append(nprob)
    return fails, counts


def screen_hamming(grnas, contigs, masks, max_mismatch):
    """Reduction (2) of SURVEY 8c (ot_gap<=1, PAM-less, no pattern): a gRNA fails iff some full-length
    window on either strand - positions beyond a contig end count as mismatches - has <= max_mismatch
    mismatches and its on-contig span is not inside a single mask.  Returns per-gRNA hit COUNTS
    (number of such windows), which is what the device kernels report."""
    by_contig = {}
    for _mid, cid, ms, me in masks:
        by_contig.setdefault(cid, []).append((ms, me))
    counts = []
    for q in grnas:
        L = len(q)
        total = 0
        for query in (q.upper(), reverse_complement(q.upper())):
            for cid, seq in contigs:
                up, n = seq.upper(), len(seq)
                for p in range(-L + 1, n):
                    a, b = max(p, 0), min(p + L, n)
                    if L - (b - a) > max_mismatch:
                        continue
                    mm = L - (b - a)
                    for j in range(a - p, b - p):
                        c = query[j]
                        if c != up[p + j] or c not in _VALID:
                            mm += 1
                    if mm > max_mismatch:
                        continue
                    if any(within_mask((a, b), m) for m in by_contig.get(cid, ())):
                        continue
                    total += 1
        counts.append(total)
    return counts
