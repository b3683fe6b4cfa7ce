#!/usr/bin/env python
"""Code generator for the carry-save adder trees of the bit-sliced mismatch scan.

For every gRNA length L in 1..32 it emits `csa_count<L>(x, b)`: x[0..L) are 32-bit words whose bit i
says "query i of this group mismatches the window at position j"; b[0..NB) receive the bit planes of the
per-query mismatch COUNT (b[k] bit i = bit k of count_i).  Full adders are two LOP3 (xor3 0x96, majority
0xE8), half adders two (xor, and).  Run:  python gen_csa.py > csa_gen.cuh
"""
import sys


def split_of(L):
    """Positions summed before the warp-level early-out test when the budget is only known at run time
    (0 = no split)."""
    if L < 10:
        return 0
    return (3 * L + 3) // 4          # 20 -> 15, 23 -> 18, 32 -> 24


def best_split(L, K):
    """Split minimising the expected loads per (warp, 32-query group) step on i.i.d. uniform sequence:
    T1 + P(some of the 32x32 cells is still within budget after T1 positions) * (L - T1)."""
    from math import comb

    def loads(t):
        p = sum(comb(t, k) * .75 ** k * .25 ** (t - k) for k in range(min(K, t) + 1))
        return t + (1 - (1 - p) ** 1024) * (L - t)
    return min(range(4, L), key=loads)


def best_pair_split(L, K):
    """Pairs of positions tested before the vote in the pair-word scan: a pair word is set when EITHER position
    mismatches (probability 15/16 on uniform sequence), so #set pair words <= #mismatches; survivors (and only
    they) pay the exact count over all L positions."""
    from math import comb
    n_pairs = (L + 1) // 2

    def loads(t):
        p = sum(comb(t, k) * (15 / 16) ** k * (1 / 16) ** (t - k) for k in range(min(K, t) + 1))
        return t + (1 - (1 - p) ** 1024) * L
    return min(range(2, n_pairs + 1), key=loads)


TUNED_L = (20, 23)        # gRNA lengths with compile-time budgets (SpCas9, Cas12a)
TUNED_K = range(0, 6)


def gen(L, first=None):
    """first=None: count x[0..L).  first=T1: add x[0..L-T1) to the partial count planes p[0..NB(T1))."""
    lines, ops = [], 0
    if first is None:
        lists = {0: ["x[%d]" % j for j in range(L)]}
    else:
        lists = {w: ["p[%d]" % w] for w in range(first.bit_length())}
        lists[0] = lists.get(0, []) + ["x[%d]" % j for j in range(L - first)]
    tmp = [0]

    def new():
        tmp[0] += 1
        return "t%d" % tmp[0]
    w = 0
    nb = L.bit_length()
    while w < nb:
        items = lists.get(w, [])
        while len(items) >= 3:
            a, b, c = items[0], items[1], items[2]
            items = items[3:]
            s, cy = new(), new()
            if w == nb - 1:      # the count never reaches 2^nb: the carry out of the top weight is always zero
                lines.append("    uint32_t %s = lop3_xor3(%s, %s, %s);" % (s, a, b, c))
                ops += 1
            else:
                lines.append("    uint32_t %s = lop3_xor3(%s, %s, %s), %s = lop3_maj(%s, %s, %s);" % (s, a, b, c, cy, a, b, c))
                ops += 2
                lists.setdefault(w + 1, []).append(cy)
            items.append(s)
        if len(items) == 2:
            a, b = items
            s, cy = new(), new()
            if w == nb - 1:
                lines.append("    uint32_t %s = %s ^ %s;" % (s, a, b))
                ops += 1
            else:
                lines.append("    uint32_t %s = %s ^ %s, %s = %s & %s;" % (s, a, b, cy, a, b))
                ops += 2
                lists.setdefault(w + 1, []).append(cy)
            items = [s]
        lines.append("    b[%d] = %s;" % (w, items[0] if items else "0u"))
        w += 1
    assert not lists.get(nb), (L, lists.get(nb))
    return nb, ops, lines


# ---------------------------------------------------------------------------------------------------------------
# Threshold networks: count_le<T, K>(x) = mask of bit columns whose count over the T words is <= K, WITHOUT forming
# the exact count: adders run only below the top weight 2^wt <= K; at that weight "at most one carry, and if one
# then the low part <= K - 2^wt" decides.  Boolean functions of <= 3 words are fused into single LOP3s.
# Every emitted network is checked exhaustively in tests/test_host_mirror.py (T <= 12).
class Net:
    """Signals are (vars, tt): a boolean function of <= 3 named words (tt indexed like LOP3: a=0xF0, b=0xCC, c=0xAA).
    Functions are fused lazily; a signal is materialised (one LOP3) only when a consumer needs more than 3 leaves."""
    def __init__(self):
        self.lines, self.n, self.cache = [], 0, {}
    def var(self, name):
        return ((name,), 0xF0)
    def const(self, v):
        return ((), 0xFF if v else 0x00)
    def _eval(self, sig, env):
        vs, tt = sig
        idx = 0
        for k in range(3):
            bit = env[vs[k]] if k < len(vs) else 0
            idx |= bit << (2 - k)
        return (tt >> idx) & 1
    def materialise(self, sig):
        vs, tt = sig
        if len(vs) == 1 and tt == 0xF0:
            return sig
        if len(vs) == 0:
            name = "0xFFFFFFFFu" if tt & 1 else "0u"
            return ((name,), 0xF0)
        if sig in self.cache:
            return self.cache[sig]
        self.n += 1
        name = "t%d" % self.n
        a = list(vs) + [vs[0]] * (3 - len(vs))
        # re-index tt for padded operands: operands beyond len(vs) alias vs[0]; evaluate consistently
        tt2 = 0
        for idx in range(8):
            bits = [(idx >> 2) & 1, (idx >> 1) & 1, idx & 1]
            env = {}
            ok = True
            for k, v in enumerate(a):
                if v in env and env[v] != bits[k]:
                    ok = False
                env[v] = bits[k]
            if ok:
                tt2 |= self._eval(sig, env) << idx
            else:   # unreachable operand combination: copy the value of a consistent assignment
                env = {v: bits[a.index(v)] for v in set(a)}
                tt2 |= self._eval(sig, env) << idx
        self.lines.append("    const uint32_t %s = lop3<0x%02X>(%s, %s, %s);" % (name, tt2, a[0], a[1], a[2]))
        self.cache[sig] = ((name,), 0xF0)
        return self.cache[sig]
    def combine(self, f, *sigs):
        sigs = list(sigs)
        while True:
            leaves = []
            for s in sigs:
                for v in s[0]:
                    if v not in leaves:
                        leaves.append(v)
            if len(leaves) <= 3:
                break
            # materialise the widest non-trivial signal
            k = max(range(len(sigs)), key=lambda i: (len(sigs[i][0]) > 1 or sigs[i][1] != 0xF0, len(sigs[i][0])))
            sigs[k] = self.materialise(sigs[k])
        tt = 0
        for idx in range(8):
            env = {v: (idx >> (2 - k)) & 1 for k, v in enumerate(leaves)}
            tt |= (1 if f(*[self._eval(s, env) for s in sigs]) else 0) << idx
        return (tuple(leaves), tt)

def gen_le(T, K):
    net = Net()
    xs = [net.var("x[%d]" % i) for i in range(T)]
    if K >= T:
        return net, net.const(1)
    if K == 0:
        acc = xs[0]
        i = 1
        while i < T:
            grp = xs[i:i + 2]
            acc = net.combine(lambda *b: int(any(b)), acc, *grp)
            i += 2
        return net, net.combine(lambda a: 1 - a, acc)
    wt = K.bit_length() - 1
    r = K - (1 << wt)
    lists = {0: list(xs)}
    low = []
    for w in range(wt):
        items = lists.get(w, [])
        while len(items) >= 3:
            a, b, c = items[:3]
            items = items[3:]
            items.append(net.combine(lambda p, q, s: p ^ q ^ s, a, b, c))
            lists.setdefault(w + 1, []).append(net.combine(lambda p, q, s: int(p + q + s >= 2), a, b, c))
        if len(items) == 2:
            a, b = items
            items = [net.combine(lambda p, q: p ^ q, a, b)]
            lists.setdefault(w + 1, []).append(net.combine(lambda p, q: p & q, a, b))
        low.append(items[0] if items else net.const(0))
    C = list(lists.get(wt, []))
    deads = []
    while len(C) > 1:
        if len(C) >= 3:
            a, b, c = C[:3]; C = C[3:]
            C.append(net.combine(lambda p, q, s: p ^ q ^ s, a, b, c))
            deads.append(net.combine(lambda p, q, s: int(p + q + s >= 2), a, b, c))
        else:
            a, b = C; C = [net.combine(lambda p, q: p ^ q, a, b)]
            deads.append(net.combine(lambda p, q: p & q, a, b))
    one = C[0] if C else net.const(0)
    dead = net.const(0)
    for d in deads:
        dead = net.combine(lambda p, q: p | q, dead, d)
    # low > r ?
    def low_gt(*bits):
        v = sum(b << i for i, b in enumerate(bits))
        return int(v > r)
    lowgt = net.combine(low_gt, *low) if low else net.const(0)
    alive = net.combine(lambda d, o, g: int((not d) and not (o and g)), dead, one, lowgt)
    return net, alive



def main():
    out = ["// GENERATED by gen_csa.py - do not edit.", "#pragma once", "#include <cstdint>", "",
           "namespace mgcsa {",
           "__device__ __forceinline__ uint32_t lop3_xor3(uint32_t a, uint32_t b, uint32_t c) {",
           "    uint32_t r; asm(\"lop3.b32 %0, %1, %2, %3, 0x96;\" : \"=r\"(r) : \"r\"(a), \"r\"(b), \"r\"(c)); return r; }",
           "__device__ __forceinline__ uint32_t lop3_maj(uint32_t a, uint32_t b, uint32_t c) {",
           "    uint32_t r; asm(\"lop3.b32 %0, %1, %2, %3, 0xE8;\" : \"=r\"(r) : \"r\"(a), \"r\"(b), \"r\"(c)); return r; }",
           "template <int L> struct CountBits { static constexpr int value = (L >= 32) ? 6 : (L >= 16) ? 5 : (L >= 8) ? 4 : (L >= 4) ? 3 : (L >= 2) ? 2 : 1; };",
           "template <int L> __device__ __forceinline__ void csa_count(const uint32_t (&x)[L], uint32_t (&b)[CountBits<L>::value]);", ""]
    out.append("// SplitOf<L, K>: positions summed before the early-out vote (0 = no split); K < 0 = budget known at run time")
    out.append("template <int L, int K> struct SplitOf { static constexpr int value = (L < 10) ? 0 : (3 * L + 3) / 4; };")
    for L in TUNED_L:
        for K in TUNED_K:
            out.append("template <> struct SplitOf<%d, %d> { static constexpr int value = %d; };" % (L, K, best_split(L, K)))
    out.append("// PairSplit<L, K>: position PAIRS whose 'some mismatch' words are summed before the vote of the pair-word scan")
    out.append("template <int L, int K> struct PairSplit { static constexpr int value = (L + 1) / 2; };")
    for L in TUNED_L:
        for K in TUNED_K:
            out.append("template <> struct PairSplit<%d, %d> { static constexpr int value = %d; };" % (L, K, best_pair_split(L, K)))
    out.append("template <int L, int T1> __device__ __forceinline__ void csa_count_rest(const uint32_t (&p)[CountBits<(T1 > 0 ? T1 : 1)>::value], const uint32_t (&x)[L - T1], uint32_t (&b)[CountBits<L>::value]);")
    out.append("")
    for L in range(1, 33):
        nb, ops, lines = gen(L)
        out.append("// L=%d: %d logic ops, %d count bits" % (L, ops, nb))
        out.append("template <> __device__ __forceinline__ void csa_count<%d>(const uint32_t (&x)[%d], uint32_t (&b)[%d]) {" % (L, L, nb))
        out.extend(lines)
        out.append("}")
        out.append("")
        splits = {split_of(L)} | ({best_split(L, K) for K in TUNED_K} if L in TUNED_L else set())
        for t1 in sorted(splits - {0}):
            nb2, ops2, lines2 = gen(L, first=t1)
            out.append("// L=%d after the first %d positions: %d logic ops" % (L, t1, ops2))
            out.append("template <> __device__ __forceinline__ void csa_count_rest<%d, %d>(const uint32_t (&p)[%d], const uint32_t (&x)[%d], uint32_t (&b)[%d]) {" % (L, t1, t1.bit_length(), L - t1, nb2))
            out.extend(lines2)
            out.append("}")
            out.append("")
    out.append("template <int LUT> __device__ __forceinline__ uint32_t lop3(uint32_t a, uint32_t b, uint32_t c) {")
    out.append("    uint32_t r; asm(\"lop3.b32 %0, %1, %2, %3, %4;\" : \"=r\"(r) : \"r\"(a), \"r\"(b), \"r\"(c), \"n\"(LUT)); return r; }")
    out.append("template <int T, int K> struct HasCountLe { static constexpr bool value = (T >= 2 && T <= 16 && K >= 0 && K <= 7 && K < T); };")
    out.append("template <int T, int K> __device__ __forceinline__ uint32_t count_le(const uint32_t (&x)[T]);")
    out.append("")
    for T in range(2, 17):
        for K in range(0, min(8, T)):
            net, alive = gen_le(T, K)
            res = net.materialise(alive)
            out.append("// at most %d of %d words set: %d logic ops" % (K, T, len(net.lines)))
            out.append("template <> __device__ __forceinline__ uint32_t count_le<%d, %d>(const uint32_t (&x)[%d]) {" % (T, K, T))
            out.extend(net.lines)
            out.append("    return %s;" % res[0][0])
            out.append("}")
            out.append("")
    out.append("}  // namespace mgcsa")
    sys.stdout.write("\n".join(out) + "\n")


if __name__ == "__main__":
    main()
