"""PAM description - host-side mirror of the reference's `minorg.pam.PAM` (pam.py:11-187) with the same
public names and behaviour, plus `program()` which lowers the PAM to the device representation
(include/minorg_b200.h: mg_pam) consumed by the enumeration and PAM-proximity kernels.

Behaviour pinned by tests/golden/pam.json (produced by executing the reference class)."""
import itertools

import regex

from . import _lib

DEFAULT_gRNA_LENGTH = 20
DEFAULT_PAM = "NGG"

# Bio.Data.IUPACData.ambiguous_dna_values (Biopython 1.79, the reference's pin); the order of the letters
# inside a class is what the echoed "PAM pattern:" line shows.
_IUPAC = dict(A="A", C="C", G="G", T="T", M="AC", R="AG", W="AT", S="CG", Y="CT", K="GT", V="ACG", H="ACT",
              D="AGT", B="CGT", X="GATC", N="GATC")

# preset names (pam.py:198-276 == constants.py:14-31)
PAM_PATTERNS = {}
for _name, _pat in (("SpCas9", ".NGG"), ("SaCas9N", ".NGRRN"), ("SaCas9T", ".NGRRT"), ("NmeCas9", ".NNNNGATT"),
                    ("CjCas9", ".NNNNRYAC"), ("StCas9", ".NNAGAAW"), ("Cas12a", "TTTV."), ("AacCas12b", "TTN."),
                    ("BhCas12b", "DTTN.")):
    PAM_PATTERNS[_name] = PAM_PATTERNS[_name.lower()] = _pat
globals().update(PAM_PATTERNS)


class PAM:
    """Tracks the raw PAM string (`raw_pam`), its expanded form (`pam`) and the gRNA length."""

    def __init__(self, pam=DEFAULT_PAM, gRNA_length=DEFAULT_gRNA_LENGTH):
        if isinstance(pam, PAM):
            # pam.py:37-42: a gRNA_length equal to the default keeps the source object's length
            self.raw_pam, self.pam = pam.raw_pam, pam.pam
            if gRNA_length is None:
                self.gRNA_length = 1
            elif gRNA_length != DEFAULT_gRNA_LENGTH:
                self.gRNA_length = gRNA_length
            else:
                self.gRNA_length = pam.gRNA_length
        else:
            self.raw_pam = self.pam = pam
            self.gRNA_length = gRNA_length
        self.parse()

    def __repr__(self):
        return self.pam

    # -- normalisation -------------------------------------------------------------------------
    def infer_full(self):
        """pam.py:55-76: no '.' => 3' PAM with an N spacer unless the pattern starts with a non-N base."""
        text = self.pam.upper()
        if "." not in text:
            if "N" not in text:
                text = "N" + text
            text = "." + text if text.startswith("N") else text + "."
        self.pam = text

    def expand_ambiguous(self):
        """pam.py:78-90: IUPAC letters become character classes."""
        pieces = []
        for ch in self.pam:
            members = _IUPAC.get(ch.upper(), ch)
            pieces.append(members if len(members) == 1 else "[%s]" % members)
        self.pam = "".join(pieces)

    def parse(self):
        self.infer_full()
        self.expand_ambiguous()

    # -- regex view (API parity; the scan itself runs on the device) -----------------------------
    def regex_string(self, gRNA_length=None):
        n = self.gRNA_length if gRNA_length is None else gRNA_length
        before, after = self.pam.split(".")
        return ("(?<=%s)" % before if before else "") + ".{%d}" % n + ("(?=%s)" % after if after else "")

    def regex(self, gRNA_length=None):
        """pam.py:103-125 (prints the pattern like the reference does)."""
        pattern = self.regex_string(gRNA_length)
        print("PAM pattern:", pattern)
        return regex.compile(pattern, regex.IGNORECASE)

    def _pattern_max_len(self, pattern):
        """pam.py:128-147, including its over-count when a {m,n} follows a [set]: the
        value only widens the PAM-proximity window, and parity needs the same window."""
        n_sets = len(regex.findall(r"(?<=\[).+?(?=\])", pattern))
        outside = "".join(regex.findall(r"(?<=^|\])[^\]\[]+?(?=$|\[)", pattern))
        extra = sum(int(body.split(",")[-1]) - 1 for body in regex.findall(r"(?<=\{).+?(?=\})", outside))
        literal = sum(map(len, regex.findall(r"(?<=^|\}).+?(?=$|\{)", outside)))
        return literal + n_sets + extra

    def five_prime(self):
        self.parse()
        return self.pam.split(".")[0]

    def three_prime(self):
        self.parse()
        return self.pam.split(".")[-1]

    def five_prime_maxlen(self):
        return self._pattern_max_len(self.five_prime())

    def three_prime_maxlen(self):
        return self._pattern_max_len(self.three_prime())

    # -- device lowering -------------------------------------------------------------------------
    @staticmethod
    def _atoms(side):
        """'TTT[ACG]' / 'T{3}[ACG]' / '[AG]{1,2}T' -> [(chars, lo, hi), ...]"""
        atoms, i = [], 0
        while i < len(side):
            ch = side[i]
            if ch == "[":
                j = side.index("]", i)
                chars, i = side[i + 1:j], j + 1
            elif ch in "{}]()|*+?\\^$":
                raise _lib.MinorgB200Error("PAM %r: only IUPAC letters, [sets] and {m,n} repeats are supported" % side)
            else:
                chars, i = ch, i + 1
            lo = hi = 1
            if i < len(side) and side[i] == "{":
                j = side.index("}", i)
                body = side[i + 1:j].split(",")
                lo = int(body[0]) if body[0] else 0
                hi = int(body[-1]) if body[-1] else None
                if hi is None:
                    raise _lib.MinorgB200Error("PAM %r: open-ended repeats are not supported" % side)
                i = j + 1
            atoms.append((chars, lo, hi))
        return atoms

    @classmethod
    def _alternatives(cls, side):
        atoms = cls._atoms(side)
        alts = []
        for reps in itertools.product(*[range(lo, hi + 1) for _, lo, hi in atoms]):
            alt = []
            for (chars, _, _), r in zip(atoms, reps):
                alt.extend([chars] * r)
            if len(alt) > _lib.MG_MAX_PAM_LEN:
                raise _lib.MinorgB200Error("PAM side %r longer than %d" % (side, _lib.MG_MAX_PAM_LEN))
            if alt not in alts:
                alts.append(alt)
        if len(alts) > _lib.MG_MAX_PAM_ALTS:
            raise _lib.MinorgB200Error("PAM side %r expands to more than %d alternatives" % (side, _lib.MG_MAX_PAM_ALTS))
        return alts

    @classmethod
    def _fill_side(cls, dst, side):
        if not side:
            dst.n_alts = 0
            return
        alts = cls._alternatives(side)
        dst.n_alts = len(alts)
        for a, alt in enumerate(alts):
            dst.len[a] = len(alt)
            for k, chars in enumerate(alt):
                four = 0
                for c in chars:
                    for b in {ord(c.upper()), ord(c.lower())}:       # re.IGNORECASE
                        dst.allow[a][k][b >> 5] |= 1 << (b & 31)
                    if c.upper() in "ACGT":
                        four |= 1 << "ACGT".index(c.upper())
                dst.allow4[a][k] = four

    def program(self):
        """The mg_pam structure for this PAM (ctypes)."""
        p = _lib.Pam()
        self._fill_side(p.five, self.five_prime())
        self._fill_side(p.three, self.three_prime())
        p.five_maxlen = self.five_prime_maxlen()
        p.three_maxlen = self.three_prime_maxlen()
        return p
