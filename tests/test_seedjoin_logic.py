"""The ALGORITHM of the seed-and-join screen (minorg_b200/csrc/screen_seedjoin.cu), restated in plain Python and checked on the
CPU against the exhaustive C oracle: the pigeonhole routes R1..R8 over the gRNA halves reach every problematic anchor, the
stage-1 tests are necessary conditions, and the canonical-alignment ownership rule counts every anchor exactly once - for the
one-gap criterion (t = (M+1)//2, g = M-t-1) and without gaps (t = M//2).  No GPU: this pins the argument, the GPU tests pin the
kernel (tests/test_gpu_seedjoin.py)."""
import itertools
import random

import numpy as np
import pytest

from oracle import c_oracle, minorg_oracle as O

L = 20


def _neighbours(key, radius):
    """All strings within Hamming distance `radius` (<= 2) of key, each once."""
    out = [key]
    if radius >= 1:
        for i, c in enumerate(key):
            for x in "ACGT":
                if x != c:
                    k1 = key[:i] + x + key[i + 1:]
                    out.append(k1)
                    if radius >= 2:
                        for j in range(i + 1, len(key)):
                            for y in "ACGT":
                                if y != key[j]:
                                    out.append(k1[:j] + y + k1[j + 1:])
    return out


def _diag(q, s, o):
    """bit j set: query position j mismatches the text at s[o + j]."""
    return sum(1 << j for j in range(L) if q[j] != s[o + j])


H1M, H2M = (1 << 10) - 1, ((1 << 10) - 1) << 10


def _pc(x):
    return bin(x).count("1")


def _canonical(D0, DM, DP, M, t, gapless):
    """sj_canonical: the route (and gap position) that owns a problematic anchor, (0, 0) if it is not problematic."""
    h, c2 = _pc(D0), _pc(D0 & H2M)
    if gapless:
        return ((1 if c2 <= t else 2), 0) if h <= M else (0, 0)
    if h <= M or (h == M + 1 and (D0 & 0x80001)):
        return (1 if c2 <= t else 2), 0
    for k in range(1, 20):                                          # deletion before k
        if _pc(DM & ((1 << k) - 1)) + _pc(D0 >> k) <= M:
            if k <= 9:
                return (1 if c2 <= t else 5), k
            if k == 10:
                return (1 if c2 <= t else 3), k
            return (3 if _pc(DM & H1M) <= t else 7), k
    for k in range(1, 19):                                          # insertion at k
        if _pc(DP & ((1 << k) - 1)) + _pc(D0 >> (k + 1)) <= M:
            if k <= 9:
                return (1 if c2 <= t else 6), k
            return (4 if _pc(DP & H1M) <= t else 8), k
    return 0, 0


def _stage1(r, k, D0, DM, DP, M, gapless):
    """The cheap tests of the kernel: must hold whenever route r (and k) owns a problematic anchor."""
    low = (1 << k) - 1
    if r == 2:
        return _pc(D0) <= (M if gapless else M + 1)
    if r == 1:
        return _pc(D0) <= M if gapless else min(_pc(DM & D0 & H1M), _pc(DP & D0 & H1M)) + _pc(D0 & H2M) <= M + 1
    if r == 3:
        return _pc(DM & H1M) + _pc(DM & D0 & H2M) <= M
    if r == 4:
        return _pc(DP & H1M) + _pc(DP & D0 & H2M) <= M + 1
    if r in (5, 7):
        return _pc(DM & low) + _pc(D0 & ~low) <= M
    return _pc(DP & low) + _pc(D0 & ~(2 * low + 1)) <= M


def _model(contig, grnas, M, gapless, lo, hi):
    """Per oriented query: anchors counted by the seed-and-join scheme over the forward windows [lo, hi)."""
    n, N = len(contig), len(grnas)
    t = M // 2 if gapless else (M + 1) // 2
    g = -1 if gapless else M - t - 1
    tbl = [dict(), dict(), dict(), dict()]
    for i, q in enumerate(grnas):
        tbl[0].setdefault(q[:10], []).append((i, 0))
        tbl[1].setdefault(q[10:], []).append((i, 0))
        for k in range(1, 10):
            tbl[2].setdefault(q[:k] + q[k + 1:10], []).append((i, k))
        for k in range(0, 9):
            tbl[3].setdefault(q[10:10 + k] + q[11 + k:], []).append((i, k))
    counts = np.zeros(2 * N, dtype=np.int64)
    seen = set()
    for strand, s in ((1, contig), (-1, O.reverse_complement(contig))):
        def consider(i, r, k, w):
            wf = w if strand > 0 else n - w - L
            if not (lo <= wf < hi):
                return
            q = grnas[i]
            D0, DM, DP = _diag(q, s, w), _diag(q, s, w - 1), _diag(q, s, w + 1)
            owner, kq = _canonical(D0, DM, DP, M, t, gapless)
            if owner == r and (r < 5 or kq == k):
                assert _stage1(r, k, D0, DM, DP, M, gapless), (r, k)
                assert (strand, i, w) not in seen                   # exactly once (each probe string is generated once)
                seen.add((strand, i, w))
                counts[i if strand > 0 else N + i] += 1
        for p in range(13, n - 33):
            for key in _neighbours(s[p:p + 10], t):
                for i, _ in tbl[0].get(key, ()):
                    consider(i, 2, 0, p)
                    if not gapless:
                        consider(i, 3, 0, p + 1)
                        consider(i, 4, 0, p - 1)
                for i, _ in tbl[1].get(key, ()):
                    consider(i, 1, 0, p - 10)
            if g < 0:
                continue
            for k in range(1, 10):
                if s[p + k] == s[p + k - 1]:
                    continue                                        # the same string as with the smaller k
                for key in _neighbours(s[p:p + k] + s[p + k + 1:p + 11], g):
                    for i, _ in tbl[0].get(key, ()):
                        consider(i, 5, k, p + 1)
                    for i, _ in tbl[1].get(key, ()):
                        consider(i, 7, 10 + k, p - 9)
            for key in _neighbours(s[p:p + 9], g):
                for i, ke in tbl[2].get(key, ()):
                    consider(i, 6, ke, p - 1)
                for i, ke in tbl[3].get(key, ()):
                    consider(i, 8, 10 + ke, p - 11)
    return counts


def _world(seed, n, n_grna):
    rng = random.Random(seed)
    contig = "".join(rng.choice("ACGT") for _ in range(n))
    grnas = []
    for i in range(n_grna):
        p = rng.randrange(40, n - 60)
        q = list(contig[p:p + L])
        for _ in range(rng.choice([0, 1, 2, 3, 4, 4, 5])):
            q[rng.randrange(L)] = rng.choice("ACGT")
        r = rng.random()
        if r < 0.35:
            k = rng.randrange(1, 19)
            q = q[:k] + q[k + 1:] + [rng.choice("ACGT")]
        elif r < 0.7:
            k = rng.randrange(1, 19)
            q = q[:k] + [rng.choice("ACGT")] + q[k:19]
        q = "".join(q)
        grnas.append(O.reverse_complement(q) if i % 2 else q)
    return contig, grnas


@pytest.mark.parametrize("M,gapless", [(4, False), (3, False), (2, False), (1, False), (0, False), (4, True), (5, True), (3, True)])
def test_routes_and_ownership_count_every_anchor_once(M, gapless):
    c_oracle.build()
    n = 1500
    contig, grnas = _world(40 + M + (10 if gapless else 0), n, 400)
    lo, hi = 40, n - 60                                             # windows with clean surroundings (the rest is the DP path's)
    got = _model(contig, grnas, M, gapless, lo, hi)
    if gapless:
        blob = np.frombuffer(contig.encode(), dtype=np.uint8)
        both = c_oracle.screen_hamming_arrays(blob, np.array([0, n], dtype=np.int64), [], grnas, M, threads=4, win_ranges=[(lo, hi)])
        assert np.array_equal(got[:len(grnas)] + got[len(grnas):], both.astype(np.int64))
    else:
        want = c_oracle.screen_nopattern([("c", contig)], [], grnas, M, 1, threads=4, win_ranges=[(lo, hi)]).astype(np.int64)
        bad = np.nonzero(got != want)[0]
        assert bad.size == 0, (bad[:5], got[bad[:5]], want[bad[:5]])
    assert got.sum() >= (100 if M >= 2 else 10)


def test_probe_counts_match_the_kernel_tables():
    """436 / 279 / 28 probe strings per position for t = 2, g = 1."""
    assert len(_neighbours("ACGTACGTAC", 2)) == 1 + 30 + 405
    assert len(set(_neighbours("ACGTACGTAC", 2))) == 436
    assert 9 * len(_neighbours("ACGTACGTAC", 1)) == 279
    assert len(_neighbours("ACGTACGTA", 1)) == 28
