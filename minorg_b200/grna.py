"""gRNA containers - the boundary types the reference's grna()/filter_background() produce and consume
(grna.py:163-1298): Target, gRNAHit, gRNASeq, gRNAHits, with the accessors and writers the hot path and
its callers use (write_fasta grna.py:1061-1089, write_mapping v5 grna.py:925-1060, set_seqs_check
grna.py:580-593, assign_seqid grna.py:630-643).  Output text pinned by tests/golden/enumerate.json."""


def _status_text(v):
    return "pass" if v is True else "fail" if v is False else "NA"


def _status_value(v):
    if isinstance(v, str):
        return True if v == "pass" else False if v == "fail" else None
    return v


class CheckObj:
    """Named tri-state checks: True/False/None <-> pass/fail/NA (grna.py:28-161).  The value table is
    created on first write: a million-hit enumeration should not allocate a million dicts."""
    _default_checks = ()
    __slots__ = ("_checks",)

    def __init__(self):
        self._checks = None

    @property
    def check_names(self):
        return list(self._default_checks) + [n for n in (self._checks or ()) if n not in self._default_checks]

    def set_check(self, name, status):
        if self._checks is None:
            self._checks = {}
        self._checks[name] = _status_value(status)

    def check(self, name, mode="bool"):
        v = self._checks.get(name) if self._checks else None
        return _status_text(v) if mode == "str" else v

    def check_exists(self, name):
        return name in self._default_checks or bool(self._checks and name in self._checks)


class Target:
    def __init__(self, seq, id=None, strand=None):
        self._seq = str(seq).upper()
        self.id, self.strand = id, strand
        self.sense = None

    def __len__(self):
        return len(self._seq)

    def __str__(self):
        return self._seq

    @property
    def seq(self):
        return self._seq


class gRNASeq(CheckObj):
    _default_checks = ("background", "GC", "exclude")
    __slots__ = ("_seq", "id")

    def __init__(self, seq):
        self._checks = None
        self._seq = str(seq).upper()
        self.id = None

    @property
    def seq(self):
        return self._seq

    def __str__(self):
        return self._seq


class gRNAHit(CheckObj):
    """One occurrence: forward-strand, 0-based, end-exclusive range on its target (grna.py:1144)."""

    _default_checks = ("background", "GC", "feature", "exclude")
    __slots__ = ("target", "_range", "strand", "hit_id", "_parent_sense")

    def __init__(self, target, start, end, strand, hit_id):
        self._checks = None
        self.target, self._range, self.strand, self.hit_id = target, (start, end), strand, hit_id
        self._parent_sense = None

    @property
    def target_id(self):
        return self.target.id

    @property
    def target_len(self):
        return len(self.target)

    @property
    def range(self):
        return tuple(self._range)

    @property
    def start(self):
        return self._range[0]

    @property
    def end(self):
        return self._range[1]

    def set_parent_sense(self, sense):
        self._parent_sense = sense

    def parent_sense(self, mode="raw"):
        s = self._parent_sense
        if mode == "str":
            return "sense" if s in ("+", "sense") else "antisense" if s in ("-", "antisense") else "NA"
        return s


class gRNAHits:
    """{SEQ: gRNASeq} and {SEQ: [gRNAHit]} in first-occurrence order (grna.py:252-1118)."""

    def __init__(self):
        self._gRNAseqs, self._hits = {}, {}

    def __len__(self):
        return len(self._gRNAseqs)

    def __repr__(self):
        return "gRNAHits(gRNA = %d)" % len(self)

    @property
    def gRNAseqs(self):
        return self._gRNAseqs

    @property
    def hits(self):
        return self._hits

    @property
    def seqs(self):
        return list(self._gRNAseqs)

    def add_hit(self, seq, hit):
        seq = str(seq).upper()
        if seq not in self._gRNAseqs:
            self._gRNAseqs[seq] = gRNASeq(seq)
            self._hits[seq] = []
        self._hits[seq].append(hit)

    def get_hits(self, seq):
        return self._hits[str(seq).upper()]

    def get_gRNAseq_by_seq(self, seq):
        return self._gRNAseqs.get(str(seq).upper())

    def get_gRNAseq_by_id(self, gid):
        for s in self._gRNAseqs.values():
            if s.id == gid:
                return s
        return None

    def _index_by_id(self):
        """{id: gRNASeq}, first occurrence wins like the linear search of get_gRNAseq_by_id (one pass for many ids)."""
        out = {}
        for s in self._gRNAseqs.values():
            out.setdefault(s.id, s)
        return out

    def flatten_gRNAseqs(self):
        return list(self._gRNAseqs.values())

    def flatten_hits(self):
        return [h for hs in self._hits.values() for h in hs]

    def assign_seqid(self, prefix="gRNA_", zfill=3, assign_all=True):
        for i, s in enumerate(self._gRNAseqs.values()):
            if s.id is None or assign_all:
                s.id = "%s%s" % (prefix, str(i + 1).zfill(zfill))

    def set_seqs_check(self, check_name, status, seqs):
        for seq in seqs:
            s = self.get_gRNAseq_by_seq(seq)
            if s is not None:
                s.set_check(check_name, status)

    def filter_seqs(self, *check_names, accept_invalid=False):
        """Sequences passing every named sequence-level check (grna.py:820-866, subset used by the path)."""
        out = []
        for s in self._gRNAseqs.values():
            vals = [s.check(n) for n in check_names]
            if all(v is True or (accept_invalid and v is None) for v in vals):
                out.append(s.seq)
        return out

    def collapse(self, ids=(), seqs=()):
        """Group gRNA by identical target coverage (grna.py:892-921): -> [(group name, sorted target ids, [gRNASeq])],
        groups ordered by (-coverage, target ids) and named cg_1.. (zero-padded to the number of groups)."""
        if seqs:
            chosen = [self.get_gRNAseq_by_seq(s) for s in seqs]
        elif ids:
            by_id = self._index_by_id()
            chosen = [by_id.get(i) for i in ids]
        else:
            chosen = self.flatten_gRNAseqs()
        groups = {}
        for gs in chosen:
            if gs is None:
                continue
            cov = tuple(sorted({h.target_id for h in self.get_hits(gs.seq)}))
            groups.setdefault(cov, []).append(gs)
        ordered = sorted(groups.items(), key=lambda kv: (-len(kv[0]), kv[0]))
        width = len(str(len(ordered)))
        return [("cg_" + str(i + 1).zfill(width), cov, members) for i, (cov, members) in enumerate(ordered)]

    def relative_5prime_pos(self, seq):
        """Mean distance of the gRNA's hits from the 5' end of their targets (minimum_set.py:101-118)."""
        hits = self.get_hits(seq)
        return sum((h.range[0] if h.target.sense != "-" else h.target_len - h.range[1]) for h in hits) / len(hits)

    def write_equivalents(self, fout, ids=(), seqs=(), write_all=False):
        """.eqv file: gRNA grouped by equivalent coverage (grna.py:1090-1118, minimum_set.py:224-245).  Groups by
        descending coverage, members by relative position.  (The reference iterates a hash-ordered set here, so
        its order among groups of equal coverage changes from run to run; ours breaks those ties by group name.)"""
        if not ids and not seqs and not write_all:
            open(fout, "w").close()
            return
        self.assign_seqid(assign_all=False)
        with open(fout, "w") as fh:
            fh.write("\t".join(["coverage group", "coverage", "gRNA id", "gRNA sequence", "relative pos"]) + "\n")
            for name, cov, members in self.collapse(ids=ids, seqs=seqs):
                for gs in sorted(members, key=lambda g: (self.relative_5prime_pos(g.seq), g.id)):
                    fh.write("\t".join([name, str(len(cov)), gs.id, gs.seq, str(round(self.relative_5prime_pos(gs.seq), 2))]) + "\n")

    def parse_from_mapping(self, fname, targets=None):
        """Read a MINORg .map file written by write_mapping (version 5; versions 2-4 are accepted too, they
        only lack the 'target length' column or call 'set' 'group'; grna.py:447-514).  Restores sequences, hits,
        ids and check columns so filter/minimumset style steps can resume from files."""
        seq_targets = {}
        if targets:
            from .fasta import fasta_to_dict
            seq_targets = {k: Target(v, id=k, strand="+") for k, v in fasta_to_dict(targets).items()}
        with open(fname) as fh:
            rows = [ln.rstrip("\n").split("\t") for ln in fh if ln.strip()]
        header, data = rows[0], rows[1:]
        set_col = header.index("set") if "set" in header else header.index("group")
        has_len = "target length" in header
        checks = header[set_col + 1:]
        dummy = {}
        for row in data:
            if has_len:
                gid, seq, tid, tlen, sense, strand, start, end = row[:8]
            else:
                gid, seq, tid, sense, strand, start, end = row[:7]
                tlen = 0
            target = seq_targets.get(tid)
            if target is None:
                target = dummy.get((tid, tlen))
                if target is None:           # one dummy per (id, length): the reference builds an N-string per ROW
                    target = dummy[(tid, tlen)] = Target("N" * int(tlen), id=tid, strand="+")
            hit = gRNAHit(target, int(start) - 1, int(end), strand, gid)
            hit.set_parent_sense(sense)
            self.add_hit(seq, hit)
            gs = self.get_gRNAseq_by_seq(seq)
            gs.id = gid
            for name, val in zip(checks, row[set_col + 1:]):
                (gs if name in ("background", "exclude", "GC") else hit).set_check(name, val)

    # -- writers ---------------------------------------------------------------------------------
    def write_fasta(self, fout, ids=(), seqs=(), write_all=False):
        if write_all:
            chosen = self.flatten_gRNAseqs()
        elif seqs:
            chosen = [self.get_gRNAseq_by_seq(s) for s in seqs if self.get_gRNAseq_by_seq(s) is not None]
        elif ids:
            by_id = self._index_by_id()
            chosen = [by_id[i] for i in ids if i in by_id]
        else:
            open(fout, "w").close()
            return
        self.assign_seqid(assign_all=False)
        with open(fout, "w") as fh:
            for s in chosen:
                fh.write(">%s\n" % s.id)
                for i in range(0, len(s.seq), 60):
                    fh.write(s.seq[i:i + 60] + "\n")

    def write_mapping(self, fout, sets=(), write_all=False, write_checks=False,
                      checks=("background", "GC", "feature")):
        """MINORg .map version 5: 1-based inclusive start/end; stable sorts by start, target id, gRNA id, set."""
        checks = list(checks) if write_checks else []
        groups = [list(s) for s in sets] if sets else ([self.seqs] if write_all else None)
        if groups is None:
            open(fout, "w").close()
            return
        header = ["gRNA id", "gRNA sequence", "target id", "target length", "target sense", "gRNA strand",
                  "start", "end", "set"] + checks
        rows = []                                       # (sort key, finished line): 10^6 rows are written in seconds
        for set_no, members in enumerate(groups, 1):
            set_str = str(set_no)
            for seq in members:
                s = self.get_gRNAseq_by_seq(seq)
                sid, sseq = s.id, s.seq
                head = "%s\t%s\t" % (sid, sseq)
                for h in self.get_hits(seq):
                    a, b = h.range
                    tail = "".join("\t" + str((s if c in ("GC", "background") else h).check(c, mode="str")) for c in checks)
                    rows.append(((set_no, sid, h.target_id, a + 1),
                                 "%s%s\t%s\t%s\t%s\t%d\t%s\t%s%s\n" % (head, h.target_id, h.target_len, h.parent_sense(mode="str"), h.strand,
                                                                        a + 1, b, set_str, tail)))
        # (the reference sorts stably by start, then target id, gRNA id, set: one sort on the reversed key tuple is the same order)
        rows.sort(key=lambda r: r[0])
        with open(fout, "w") as fh:
            fh.write("\t".join(header) + "\n")
            fh.writelines(line for _, line in rows)
