"""Minimal stand-ins for the third-party modules the reference imports (Biopython, pyfaidx)
so that a few pure-Python reference modules can be EXECUTED in this container to generate
golden vectors.  Used only by tests/golden/make_golden.py (never at test/bench run time:
/root/reference does not exist on the GPU box).  Behaviour of each stand-in follows
SURVEY.md Appendix B (Biopython 1.79 / pyfaidx 0.6.4 as pinned by the reference's setup.py)."""
import sys
import types

_COMP = str.maketrans("ACGTMRWSYKVHDBNXacgtmrwsykvhdbnx", "TGCAKYWSRMBDHVNXtgcakywsrmbdhvnx")


class Seq:
    def __init__(self, data):
        self._d = str(data)

    def __str__(self):
        return self._d

    def __repr__(self):
        return "Seq(%r)" % self._d

    def __len__(self):
        return len(self._d)

    def __getitem__(self, k):
        r = self._d[k]
        return Seq(r) if isinstance(k, slice) else r

    def __eq__(self, o):
        return str(self) == str(o)

    def __hash__(self):
        return hash(self._d)

    def __iter__(self):
        return iter(self._d)

    def __add__(self, o):
        return Seq(self._d + str(o))

    def upper(self):
        return Seq(self._d.upper())

    def lower(self):
        return Seq(self._d.lower())

    def find(self, sub, start=0, end=None):
        return self._d.find(str(sub), start, len(self._d) if end is None else end)

    def count(self, sub):
        return self._d.count(str(sub))

    def complement(self):
        return Seq(self._d.translate(_COMP))

    def reverse_complement(self):
        return Seq(self._d.translate(_COMP)[::-1])


class SeqRecord:
    def __init__(self, seq, id="<unknown id>", name="", description=""):
        self.seq, self.id, self.name, self.description = seq, id, name or id, description


def _fasta_parse(handle, fmt="fasta"):
    assert fmt == "fasta"
    close = False
    if isinstance(handle, (str, bytes)) or hasattr(handle, "__fspath__"):
        handle = open(handle)
        close = True
    try:
        rid, descr, chunks = None, "", []
        for line in handle:
            if line.startswith(">"):
                if rid is not None:
                    yield SeqRecord(Seq("".join(chunks)), id=rid, description=descr)
                descr = line[1:].rstrip()
                rid = descr.split(None, 1)[0] if descr.split() else ""
                chunks = []
            elif rid is not None:
                chunks.append("".join(line.split()))
        if rid is not None:
            yield SeqRecord(Seq("".join(chunks)), id=rid, description=descr)
    finally:
        if close:
            handle.close()


def _fasta_write(records, handle, fmt="fasta"):
    close = False
    if isinstance(handle, str):
        handle = open(handle, "w")
        close = True
    n = 0
    for r in records:
        handle.write(">%s\n" % (r.description if r.description and r.description != "<unknown description>" else r.id))
        s = str(r.seq)
        for i in range(0, len(s), 60):
            handle.write(s[i:i + 60] + "\n")
        n += 1
    if close:
        handle.close()
    return n


AMBIGUOUS_DNA_VALUES = {
    "A": "A", "C": "C", "G": "G", "T": "T", "M": "AC", "R": "AG", "W": "AT", "S": "CG",
    "Y": "CT", "K": "GT", "V": "ACG", "H": "ACT", "D": "AGT", "B": "CGT", "X": "GATC", "N": "GATC",
}


def install():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    bio = mod("Bio", BiopythonWarning=Warning)
    bio.Seq = mod("Bio.Seq", Seq=Seq)
    bio.SeqRecord = mod("Bio.SeqRecord", SeqRecord=SeqRecord)
    bio.SeqIO = mod("Bio.SeqIO", parse=_fasta_parse, write=_fasta_write)
    bio.SearchIO = mod("Bio.SearchIO", parse=lambda *a, **k: iter(()))
    bio.Data = mod("Bio.Data")
    bio.Data.IUPACData = mod("Bio.Data.IUPACData", ambiguous_dna_values=AMBIGUOUS_DNA_VALUES)
    bio.Blast = mod("Bio.Blast")
    bio.Blast.Applications = mod("Bio.Blast.Applications", NcbiblastnCommandline=object)

    class _Any:
        def __init__(self, *a, **k):
            pass
    mod("Bio.Application", _Option=_Any, _Switch=_Any, _Argument=_Any, AbstractCommandline=_Any,
        ApplicationError=Exception)
    mod("pybedtools", BedTool=_Any)

    class _Stub:
        def __init__(self, *a, **k):
            pass
    mod("pyfaidx", Fasta=_Stub, FastaRecord=_Stub, Faidx=_Stub, MutableFastaRecord=_Stub,
        Sequence=_Stub, OrderedDict=dict, FastaIndexingError=Exception, UnsupportedCompressionFormat=Exception,
        BedError=Exception, FetchError=Exception)
    if "/root/reference" not in sys.path:
        sys.path.insert(0, "/root/reference")
