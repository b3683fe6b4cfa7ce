"""Off-target pattern expressions - host-side mirror of the reference's `minorg.ot_regex`
(`ot_range` ot_regex.py:63-122, `OffTargetExpression` ot_regex.py:124-356): same names, same accept/reject
behaviour, same left-to-right (no precedence) semantics - parsed into a small tree that is (a) evaluable on
the host on any object exposing `_num_char_in_range` like the reference's BlastHSP, and (b) lowered to the
postfix unit program of `mg_criteria` that the device verifier executes on every candidate alignment.

Behaviour pinned by tests/golden/ot_range.json and ot_pattern.json (produced by executing the reference)."""
import re

from . import _lib


class MINORgError(Exception):
    """Same role as minorg.MINORgError (minorg/__init__.py)."""


CHAR_UNALIGNED, CHAR_DEL, CHAR_INS, CHAR_MISMATCH, CHAR_MATCH, CHAR_GAP = " ", "d", "i", "m", ".", "g"

_HEAD = re.compile(r"-*\d+")
_TAIL = re.compile(r"-(-*\d+)$")


def ot_range(range_pattern):
    """'5' -> (None, 5); '-5' -> (-5, None); '5-' -> (4, None); '-5-' -> (None, -4); '2-4' -> (1, 4);
    '-2--4' -> (-4, -1).  The pair is used as a Python slice [start:end] on the per-position alignment
    tuple.  Zero, mixed signs and anything unparsable raise MINORgError (ot_regex.py:63-122)."""
    bad = MINORgError("Unable to parse off-target pattern range: %s" % (range_pattern,))
    head = _HEAD.match(range_pattern or "")
    if head is None:
        raise bad
    first_txt = head.group(0)
    try:
        first, negative = int(first_txt), first_txt.startswith("-")
    except ValueError:          # e.g. '--5'
        raise bad
    remainder = range_pattern[len(first_txt):]
    if first == 0:
        raise bad
    if remainder == "":
        return (first, None) if negative else (None, first)
    if remainder == "-":
        if negative:
            return (None, first + 1 if first < -1 else None)
        return (first - 1, None)
    tail = _TAIL.search(range_pattern)
    if tail is None:
        raise bad
    second_txt = tail.group(1)
    if second_txt.startswith("-") != negative:
        raise bad
    try:
        second = int(second_txt)
    except ValueError:
        raise bad
    if second == 0:
        raise bad
    if negative:
        return (second, first + 1 if first < -1 else None)
    return (first - 1, second)


class _Unit:
    __slots__ = ("text", "n", "mismatch", "ins", "dele", "gap", "start", "end", "rvs")

    def __init__(self, text):
        lead = re.match(r"\d+", text)
        rng = re.search(r"[-\d]+$", text)
        if lead is None or rng is None:
            raise MINORgError("Unable to parse off-target pattern: %s" % text)
        self.text = text
        self.n = int(lead.group(0))
        self.mismatch = CHAR_MISMATCH in text
        self.gap = CHAR_GAP in text
        self.ins = CHAR_INS in text
        self.dele = CHAR_DEL in text
        if self.gap and (self.ins or self.dele):
            # the reference touches a missing attribute here (ot_regex.py:251) and re-raises (:263-264)
            raise MINORgError("Unable to parse off-target pattern: %s" % text)
        try:
            self.start, self.end = ot_range(rng.group(0))
        except MINORgError:
            raise MINORgError("Unable to parse off-target pattern: %s" % text)
        self.rvs = any(v is not None and v < 0 for v in (self.start, self.end))

    @property
    def chars(self):
        return ((CHAR_MISMATCH,) if self.mismatch else ()) + ((CHAR_INS,) if self.ins or self.gap else ()) + \
               ((CHAR_DEL,) if self.dele or self.gap else ())


class OffTargetExpression:
    """Callable like the reference class: expr(blasthsp) -> True if the hit is problematic."""

    def __init__(self, pattern, unaligned_as_mismatch=True, unaligned_as_gap=False):
        self.pattern = pattern
        self.unaligned_as_mismatch = unaligned_as_mismatch
        self.unaligned_as_gap = unaligned_as_gap
        self.tree, self.length = self._scan(pattern)
        self.f = self.evaluate

    def __len__(self):
        return self.length

    def __call__(self, v):
        return self.evaluate(v)

    # ---- parsing: one left-to-right pass, operators applied as they are met -------------------
    def _scan(self, text):
        tree, pending_op = None, None

        def attach(node):
            nonlocal tree
            if tree is None:
                tree = node
            elif pending_op in ("and", "or"):
                tree = (pending_op, tree, node)
            else:
                raise MINORgError("self._op must be 'and' or 'or' if self.f is not None")
        if not isinstance(text, str):
            raise TypeError("pattern must be str")
        size = len(text)
        if size == 0:
            raise TypeError("empty off-target pattern")     # the reference fails on f(None) at call time
        unit_from = pos = 0
        while pos < size:
            ch = text[pos]
            if ch == "(":
                sub, used = self._scan(text[pos + 1:])
                if sub is None:
                    raise MINORgError("Unable to parse off-target pattern: %s" % text)
                attach(sub)
                unit_from = pos = pos + 1 + used
            elif ch == "," or ch == "|":
                if unit_from != pos:
                    attach(("unit", _Unit(text[unit_from:pos])))
                pending_op = "and" if ch == "," else "or"
                pos += 1
                unit_from = pos
            elif ch == ")":
                attach(("unit", _Unit(text[unit_from:pos])))
                pos += 1
                break
            elif pos >= size - 1:
                attach(("unit", _Unit(text[unit_from:size])))
                pos += 1
                break
            else:
                pos += 1
        return tree, pos

    # ---- host evaluation (API parity with the reference object) -------------------------------
    def _unit_true(self, u, hsp):
        total = hsp._num_char_in_range(*u.chars, start=u.start, end=u.end, rvs_index=u.rvs)
        una = hsp._num_char_in_range(CHAR_UNALIGNED, start=u.start, end=u.end, rvs_index=u.rvs)
        if u.mismatch and self.unaligned_as_mismatch:
            total += una
        if (u.ins or u.gap) and self.unaligned_as_gap:
            total += una
        return total <= u.n

    def evaluate(self, hsp):
        def walk(node):
            if node[0] == "unit":
                return self._unit_true(node[1], hsp)
            a = walk(node[1])
            if node[0] == "and":
                return a and walk(node[2])
            return a or walk(node[2])
        return walk(self.tree)

    # ---- device lowering ----------------------------------------------------------------------
    def units(self):
        out = []

        def walk(node):
            if node[0] == "unit":
                out.append(node[1])
            else:
                walk(node[1])
                walk(node[2])
        walk(self.tree)
        return out

    def lower(self, crit):
        """Fill units/ops of an `_lib.Criteria` (postfix order)."""
        units, ops = [], []

        def walk(node):
            if node[0] == "unit":
                units.append(node[1])
                ops.append(len(units) - 1)
            else:
                walk(node[1])
                walk(node[2])
                ops.append(_lib.MG_OP_AND if node[0] == "and" else _lib.MG_OP_OR)
        walk(self.tree)
        if len(units) > _lib.MG_MAX_PATTERN_UNITS or len(ops) > _lib.MG_MAX_PATTERN_OPS:
            raise MINORgError("off-target pattern too long for the device program")
        crit.n_units, crit.n_ops = len(units), len(ops)
        none = _lib.MG_SLICE_NONE
        for i, u in enumerate(units):
            d = crit.units[i]
            d.n = u.n
            d.chars = (_lib.MG_CH_MISMATCH if u.mismatch else 0) | (_lib.MG_CH_INS if (u.ins or u.gap) else 0) | \
                      (_lib.MG_CH_DEL if (u.dele or u.gap) else 0)
            d.start = none if u.start is None else u.start
            d.end = none if u.end is None else u.end
            d.rvs = int(u.rvs)
            d.add_unaligned_m = int(u.mismatch and self.unaligned_as_mismatch)
            d.add_unaligned_g = int((u.ins or u.gap) and self.unaligned_as_gap)
        for i, o in enumerate(ops):
            crit.ops[i] = o

    def edit_bound(self, length, max_gap):
        """A bound B such that `expression true` implies mismatches + gap characters + unaligned positions
        <= B for the alignment (hence full-length edit cost <= B): the budget of the device prefilter.
        Returns -1 when the pattern leaves some position or error type unconstrained (then every cell is
        handed to the exact verifier).  Sound by construction: a unit 'n m.. <range>' that also counts
        unaligned positions caps the errors inside its range by n (+ max_gap if it ignores a gap type); the
        caps of units whose ranges jointly cover every tuple element add up."""
        def slice_mask(u, n):
            a = 0 if u.start is None else u.start
            b = length if u.end is None else u.end          # blast.py:138-139: None -> query length
            step = 1 if b >= a else -1
            m = 0
            for i in list(range(n))[a:b:step]:
                m |= 1 << i
            return m
        worst = 0
        for clause in self.clauses():
            usable = [u for u in clause if u.mismatch and self.unaligned_as_mismatch]
            best = None
            sizes = [length] + [length + k for k in range(1, max_gap // 2 + 1)]
            # a cover must hold for every possible tuple length (extra 'dd' elements, blast.py:114-116)
            cand = usable[:12]
            for pick in range(1, 1 << len(cand)):
                chosen = [u for i, u in enumerate(cand) if (pick >> i) & 1]
                ok = True
                for n in sizes:
                    full = (1 << n) - 1
                    cov = 0
                    for u in chosen:
                        cov |= slice_mask(u, n if not u.rvs else length)
                    if (n == length and cov != full) or (n > length and any(not u.rvs for u in chosen) and cov != full):
                        ok = False
                        break
                if not ok:
                    continue
                cost = sum(u.n for u in chosen)
                if any(not (u.gap or (u.ins and u.dele)) for u in chosen):
                    cost += max_gap
                if best is None or cost < best:
                    best = cost
            if best is None:
                return -1
            worst = max(worst, best)
        return worst

    def exact_pieces(self, length, max_gap, min_piece=None):
        """Pigeonhole prefilter: gRNA position ranges [(start, len)] such that every alignment satisfying the
        pattern matches the subject exactly and gap-free on at least one of them, or None if no such set of
        useful (>= min_piece bases) ranges can be proven.  For each clause of the disjunctive normal form take
        the unit 'n m.. <range>' (counting unaligned positions as mismatches) whose range holds the longest
        run R of consecutive positions: at most n errors (+ max_gap if it ignores a gap type) fall inside R,
        so one of n_eff + 1 equal parts of R is error-free.  Only for max_gap <= 1 (no 'dd' tuple elements, so
        tuple index == gRNA position, blast.py:102-117)."""
        if max_gap > 1:
            return None
        if min_piece is None:
            min_piece = max(5, min(8, length // 2))     # 4^-8 survivors per cell at L >= 16
        pieces = []
        for clause in self.clauses():
            best = None
            for u in clause:
                if not (u.mismatch and self.unaligned_as_mismatch):
                    continue
                a = 0 if u.start is None else u.start
                b = length if u.end is None else u.end
                idx = sorted(list(range(length))[a:b:(1 if b >= a else -1)])
                run, cur = [], []
                for i in idx:                      # longest run of consecutive positions
                    cur = cur + [i] if cur and i == cur[-1] + 1 else [i]
                    if len(cur) > len(run):
                        run = cur
                if not run:
                    continue
                n_eff = u.n + (0 if (u.gap or (u.ins and u.dele)) else max_gap)
                part = len(run) // (n_eff + 1)
                if part >= min_piece and (best is None or part > best[0]):
                    best = (part, run[0], n_eff + 1)
            if best is None:
                return None
            part, first, k = best
            pieces.extend((first + i * part, part) for i in range(k))
        pieces = sorted(set(pieces))
        return pieces if 0 < len(pieces) <= _lib.MG_MAX_PIECES else None

    def clauses(self):
        """Disjunctive normal form: list of clauses, each a list of units (the tree has only AND/OR)."""
        def walk(node):
            if node[0] == "unit":
                return [[node[1]]]
            a, b = walk(node[1]), walk(node[2])
            if node[0] == "or":
                return a + b
            return [x + y for x in a for y in b]
        return walk(self.tree)


def make_criteria(ot_mismatch=1, ot_gap=0, ot_pattern=None, ot_unaligned_as_mismatch=True,
                  ot_unaligned_as_gap=False, ot_pamless=True, grna_length=None):
    """The reference's screen attributes (MINORg.py:686-707) -> `_lib.Criteria`.
    max_mismatch / max_gap follow MINORg.py:2033-2034.  grna_length lets the host derive the prefilter
    budget of a pattern (without it every cell goes through the exact verifier)."""
    c = _lib.Criteria()
    c.max_mismatch = max(1, int(ot_mismatch)) - 1
    c.max_gap = max(1, int(ot_gap)) - 1
    if c.max_gap > _lib.MG_MAX_GAPS:
        raise MINORgError("ot_gap > %d is not supported by the device screen" % (_lib.MG_MAX_GAPS + 1))
    c.pamless = int(bool(ot_pamless))
    c.prefilter_edits = -1
    if ot_pattern is not None:
        expr = ot_pattern if isinstance(ot_pattern, OffTargetExpression) else \
            OffTargetExpression(ot_pattern, ot_unaligned_as_mismatch, ot_unaligned_as_gap)
        expr.lower(c)
        if grna_length is not None:
            c.prefilter_edits = expr.edit_bound(int(grna_length), c.max_gap)
            pieces = expr.exact_pieces(int(grna_length), c.max_gap)
            # without gaps a finite Hamming budget is the tighter prefilter; pieces serve gapped / unbounded patterns
            if pieces and (c.max_gap > 0 or c.prefilter_edits < 0):
                c.n_pieces = len(pieces)
                for i, (a, n) in enumerate(pieces):
                    c.piece_start[i], c.piece_len[i] = a, n
    return c
