"""ctypes loader of the C oracle (oracle/screen_oracle.c).  Test infrastructure only - see the header of
screen_oracle.c.  Used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libscreen_oracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "screen_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        # -march=native is avoided: the .so built in the authoring container travels to the GPU box
        subprocess.check_call(["gcc", "-O3", "-mpopcnt", "-shared", "-fPIC", "-pthread", src, "-o", _SO])
    return _SO


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
    return _lib


def screen_hamming(contigs, masks, grnas, max_mismatch, threads=1):
    """contigs: [(name, seq)], masks: [(contig index, start, end)], grnas: [str] -> uint32 counts per gRNA."""
    bufs = [c[1].encode() if isinstance(c[1], str) else bytes(c[1]) for c in contigs]
    offsets = np.zeros(len(bufs) + 1, dtype=np.int64)
    np.cumsum([len(b) for b in bufs], out=offsets[1:])
    blob = np.frombuffer(b"".join(bufs) or b"\0", dtype=np.uint8)
    return screen_hamming_arrays(blob, offsets, masks, grnas, max_mismatch, threads)


def screen_hamming_arrays(blob, offsets, masks, grnas, max_mismatch, threads=1, win_ranges=None, bitsliced=False):
    """win_ranges (optional): per contig (lo, hi) - only windows whose start lies in [lo, hi) are counted.
    bitsliced=True: the bit-sliced CPU port (oracle_screen_hamming_bitsliced; same counts, ~10x faster) - the CPU baseline."""
    masks = list(masks)
    mc = np.array([m[0] for m in masks] or [0], dtype=np.int32)
    ms = np.array([m[1] for m in masks] or [0], dtype=np.int64)
    me = np.array([m[2] for m in masks] or [0], dtype=np.int64)
    g = [s.encode() if isinstance(s, str) else bytes(s) for s in grnas]
    L = len(g[0]) if g else 1
    gb = np.frombuffer(b"".join(g) or b"\0", dtype=np.uint8)
    out = np.zeros(max(1, len(g)), dtype=np.uint32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    if win_ranges is None:
        lo = hi = None
    else:
        lo = np.array([r[0] for r in win_ranges], dtype=np.int64)
        hi = np.array([r[1] for r in win_ranges], dtype=np.int64)
        assert len(lo) == len(offsets) - 1
    fn = _load().oracle_screen_hamming_bitsliced if bitsliced else _load().oracle_screen_hamming_win
    rc = fn(p(blob), p(offsets), C.c_int(len(offsets) - 1), p(mc), p(ms), p(me),
                                           C.c_int(len(masks)), p(gb), C.c_int(len(g)), C.c_int(L), C.c_int(max_mismatch),
                                           C.c_int(threads), p(out), None if lo is None else p(lo), None if hi is None else p(hi))
    if rc != 0:
        raise RuntimeError("oracle_screen_hamming failed: %d" % rc)
    return out[:len(g)]


def _pam_side(side_regex):
    """'TTT[ACG]' / '[GATC][AG]{1,2}T' -> (n alternatives, int32 lengths, uint8 allow[alt*16+k]) for the C oracle."""
    import itertools
    atoms, i = [], 0
    while i < len(side_regex):
        if side_regex[i] == "[":
            j = side_regex.index("]", i)
            chars, i = side_regex[i + 1:j], j + 1
        else:
            chars, i = side_regex[i], i + 1
        lo = hi = 1
        if i < len(side_regex) and side_regex[i] == "{":
            j = side_regex.index("}", i)
            body = side_regex[i + 1:j].split(",")
            lo, hi, i = int(body[0]), int(body[-1]), j + 1
        atoms.append((chars, lo, hi))
    alts = []
    for reps in itertools.product(*[range(lo, hi + 1) for _, lo, hi in atoms]):
        alt = [c for (c, _, _), r in zip(atoms, reps) for _ in range(r)]
        if alt not in alts:
            alts.append(alt)
    lens = np.array([len(a) for a in alts] or [0], dtype=np.int32)
    allow = np.zeros(max(1, len(alts)) * 16, dtype=np.uint8)
    for a, alt in enumerate(alts):
        for k, chars in enumerate(alt):
            allow[a * 16 + k] = sum(1 << "ACGT".index(ch) for ch in set(chars.upper()) if ch in "ACGT")
    return (len(alts) if side_regex else 0), lens, allow


def _win_arrays(win_ranges, n_contigs):
    if win_ranges is None:
        return None, None
    lo = np.array([r[0] for r in win_ranges], dtype=np.int64)
    hi = np.array([r[1] for r in win_ranges], dtype=np.int64)
    assert len(lo) == n_contigs
    return lo, hi


def screen_nopattern(contigs, masks, grnas, max_mismatch, max_gap, threads=1, pam=None, win_ranges=None):
    """Exhaustive no-pattern screen with gaps/trimming: uint32[2N] distinct-anchor counts per oriented query
    ([i] = '+' strand, [N+i] = '-' strand) - see oracle_screen_nopattern in screen_oracle.c.  pam = an
    oracle.minorg_oracle.OraclePAM to require PAM proximity (ot_pamless=False), None for the PAM-less default."""
    bufs = [c[1].encode() if isinstance(c[1], str) else bytes(c[1]) for c in contigs]
    offsets = np.zeros(len(bufs) + 1, dtype=np.int64)
    np.cumsum([len(b) for b in bufs], out=offsets[1:])
    blob = np.frombuffer(b"".join(bufs) or b"\0", dtype=np.uint8)
    masks = list(masks)
    mc = np.array([m[0] for m in masks] or [0], dtype=np.int32)
    ms = np.array([m[1] for m in masks] or [0], dtype=np.int64)
    me = np.array([m[2] for m in masks] or [0], dtype=np.int64)
    g = [s.encode() if isinstance(s, str) else bytes(s) for s in grnas]
    L = len(g[0]) if g else 1
    gb = np.frombuffer(b"".join(g) or b"\0", dtype=np.uint8)
    out = np.zeros(max(1, 2 * len(g)), dtype=np.uint32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    if pam is None:
        n5, len5, allow5, n3, len3, allow3, max5, max3, pamless = 0, np.zeros(1, np.int32), np.zeros(16, np.uint8), 0, \
            np.zeros(1, np.int32), np.zeros(16, np.uint8), 0, 0, 1
    else:
        n5, len5, allow5 = _pam_side(pam.five_prime())
        n3, len3, allow3 = _pam_side(pam.three_prime())
        max5, max3, pamless = pam.five_prime_maxlen(), pam.three_prime_maxlen(), 0
    lo, hi = _win_arrays(win_ranges, len(offsets) - 1)
    rc = _load().oracle_screen_nopattern_win(p(blob), p(offsets), C.c_int(len(offsets) - 1), p(mc), p(ms), p(me),
                                             C.c_int(len(masks)), p(gb), C.c_int(len(g)), C.c_int(L), C.c_int(max_mismatch),
                                             C.c_int(max_gap), C.c_int(threads), p(out), C.c_int(pamless),
                                             C.c_int(n5), p(len5), p(allow5), C.c_int(max5), C.c_int(n3), p(len3), p(allow3), C.c_int(max3),
                                             None if lo is None else p(lo), None if hi is None else p(hi))
    if rc != 0:
        raise RuntimeError("oracle_screen_nopattern failed: %d" % rc)
    return out[:2 * len(g)]


def pattern_parse_ok(pattern, uam=True, uag=False):
    """True iff the C parser accepts the pattern (the reference raises MINORgError otherwise)."""
    n = C.c_int(0)
    return _load().oracle_pattern_parse(pattern.encode(), C.c_int(int(uam)), C.c_int(int(uag)), C.byref(n)) == 0


def pattern_eval(pattern, kinds, qstart, qend, qlen, uam=True, uag=False):
    """OffTargetExpression(pattern)(alignment) by the C restatement: True / False, None on a parse error."""
    r = _load().oracle_pattern_eval(pattern.encode(), C.c_int(int(uam)), C.c_int(int(uag)), kinds.encode(), C.c_int(qstart),
                                    C.c_int(qend), C.c_int(qlen))
    return None if r < 0 else bool(r)


def screen_pattern(contigs, masks, grnas, pattern, ot_gap=0, uam=True, uag=False, threads=1, pam=None, prune=True, win_ranges=None):
    """Exhaustive --ot-pattern screen (oracle_screen_pattern): uint32[2N] distinct-anchor counts per oriented query.
    ot_gap is the reference attribute (gap characters allowed inside an alignment = max(1, ot_gap) - 1)."""
    bufs = [c[1].encode() if isinstance(c[1], str) else bytes(c[1]) for c in contigs]
    offsets = np.zeros(len(bufs) + 1, dtype=np.int64)
    np.cumsum([len(b) for b in bufs], out=offsets[1:])
    blob = np.frombuffer(b"".join(bufs) or b"\0", dtype=np.uint8)
    masks = list(masks)
    mc = np.array([m[0] for m in masks] or [0], dtype=np.int32)
    ms = np.array([m[1] for m in masks] or [0], dtype=np.int64)
    me = np.array([m[2] for m in masks] or [0], dtype=np.int64)
    g = [s.encode() if isinstance(s, str) else bytes(s) for s in grnas]
    L = len(g[0]) if g else 1
    gb = np.frombuffer(b"".join(g) or b"\0", dtype=np.uint8)
    out = np.zeros(max(1, 2 * len(g)), dtype=np.uint32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    if pam is None:
        n5, len5, allow5, n3, len3, allow3, max5, max3, pamless = 0, np.zeros(1, np.int32), np.zeros(16, np.uint8), 0, \
            np.zeros(1, np.int32), np.zeros(16, np.uint8), 0, 0, 1
    else:
        n5, len5, allow5 = _pam_side(pam.five_prime())
        n3, len3, allow3 = _pam_side(pam.three_prime())
        max5, max3, pamless = pam.five_prime_maxlen(), pam.three_prime_maxlen(), 0
    lo, hi = _win_arrays(win_ranges, len(offsets) - 1)
    rc = _load().oracle_screen_pattern(p(blob), p(offsets), C.c_int(len(offsets) - 1), p(mc), p(ms), p(me), C.c_int(len(masks)),
                                       p(gb), C.c_int(len(g)), C.c_int(L), pattern.encode(), C.c_int(int(uam)), C.c_int(int(uag)),
                                       C.c_int(max(1, ot_gap) - 1), C.c_int(int(prune)), C.c_int(threads), p(out), C.c_int(pamless),
                                       C.c_int(n5), p(len5), p(allow5), C.c_int(max5), C.c_int(n3), p(len3), p(allow3), C.c_int(max3),
                                       None if lo is None else p(lo), None if hi is None else p(hi))
    if rc != 0:
        raise RuntimeError("oracle_screen_pattern failed: %d" % rc)
    return out[:2 * len(g)]
