// Exact evaluation of the reference's per-hit predicates on the device, over the exhaustive alignment
// universe E1: for one anchor (gRNA, strand, virtual end e) enumerate every local alignment
// whose last diagonal ends at e - contiguous query range [qs,qe), first/last column a base pair, at most
// Gcap gap characters strictly inside - and decide  is_offtarget_pos AND is_offtarget_aln AND
// is_offtarget_pam  (MINORg.py:2105-2121) for each:
//   * position : MINORg._is_offtarget_pos (MINORg.py:1994-2000) - span inside ONE masked interval => on-target
//   * alignment: MINORg._is_offtarget_aln_nopattern (MINORg.py:2018-2044), or the OffTargetExpression
//                program (ot_regex.py:222-356) over the per-position tuples of BlastHSP (blast.py:46-172,
//                including the 'dd' grouping of blast.py:114-116)
//   * PAM      : MINORg._has_pam (MINORg.py:2063-2087), windows clamped to the contig, case-insensitive.
// Runs only on prefilter survivors (cold path).
#pragma once
#include "common.cuh"

namespace mg {

struct UnitDev {
    int32_t n, chars, start, end, rvs, add_m, add_g;
    uint32_t mask;   // positions selected by [start:end] on a tuple of exactly L elements (Python slice rules)
};

struct PamSideDev {
    int32_t n_alts;
    int32_t len[MG_MAX_PAM_ALTS];
    uint8_t allow4[MG_MAX_PAM_ALTS][MG_MAX_PAM_LEN];
};

struct GeneralCtx {
    int32_t L, M, Gp, pamless, n_units, n_ops, n_grna;
    UnitDev units[MG_MAX_PATTERN_UNITS];
    int32_t ops[MG_MAX_PATTERN_OPS];
    PamSideDev five, three;
    int32_t five_max, three_max;
    int32_t prune_pattern;    // 1: partial pattern evaluation is a sound lower bound (Gp <= 1: no 'dd' elements)
    int32_t edit_bound;       // >= 0: host-proved bound of a gapped pattern (true => mm + gaps + unaligned <= this); -1: none
    int32_t n_pieces;         // exact-piece prefilter: position masks in gRNA orientation and mirrored (rc queries)
    uint32_t piece_fwd[MG_MAX_PIECES], piece_rev[MG_MAX_PIECES];
    int32_t piece_start[MG_MAX_PIECES], piece_len[MG_MAX_PIECES];   // the same ranges as (start, length) in gRNA orientation
    // prefilter survivors are queued here and verified by a second kernel with one thread per candidate
    // (verifying inline would run one lane of a warp at a time and drag a 2 KB stack into the scan kernel)
    struct Cand* cand;
    unsigned long long* cand_count;
    unsigned long long cand_cap;
};

struct Cand {
    int32_t q;        // oriented query index: q < N = gRNA q on '+', q >= N = gRNA q-N on '-'
    int32_t tag;      // 0, or kCandTagged | piece << 8 | (offset + Gp): which (piece, offset) of the exact-piece prefilter queued it
    int64_t e;        // anchor: virtual end of the alignment in strand coordinates
};
constexpr int32_t kCandTagged = 1 << 16;
constexpr int kPieceMax = 16;          // positions of a piece that are compared (a longer piece is cut: still necessary)

// Python slice [start:end:step] with concrete bounds on a sequence of n elements; step = +1 if end >= start
// else -1 (blast.py:140-147).  Visits indices first, first+step, ... while (step > 0 ? i < last : i > last).
__host__ __device__ inline void py_slice_bounds(int start, int end, int n, int& first, int& last, int& step) {
    step = (end >= start) ? 1 : -1;
    if (step > 0) {
        first = start < 0 ? (start + n < 0 ? 0 : start + n) : (start > n ? n : start);
        last = end < 0 ? (end + n < 0 ? 0 : end + n) : (end > n ? n : end);
    } else {
        first = start < 0 ? (start + n < 0 ? -1 : start + n) : (start > n - 1 ? n - 1 : start);
        last = end < 0 ? (end + n < 0 ? -1 : end + n) : (end > n - 1 ? n - 1 : end);
    }
}

__host__ __device__ inline bool py_slice_has(int first, int last, int step, int idx) {
    return step > 0 ? (idx >= first && idx < last) : (idx <= first && idx > last);
}

__host__ __device__ inline uint32_t py_slice_mask(int start_or_none, int end_or_none, int qlen, int n) {
    const int start = start_or_none == MG_SLICE_NONE ? 0 : start_or_none;
    const int end = end_or_none == MG_SLICE_NONE ? qlen : end_or_none;      // blast.py:138-139: None -> qlen
    int first, last, step;
    py_slice_bounds(start, end, n, first, last, step);
    uint32_t m = 0;
    for (int i = 0; i < n && i < 32; ++i) if (py_slice_has(first, last, step, i)) m |= 1u << i;
    return m;
}

// Strand-aware text: strand +1 = forward genome; -1 = reverse complement, position i' = glen-1-i.
__device__ __forceinline__ int tbase(const GenomeView& gv, int strand, int64_t i) {
    if (strand > 0) return base_at(gv, i);
    const int c = base_at(gv, gv.glen - 1 - i);
    return c == kCodeInvalid ? c : 3 - c;
}

struct AlnState {
    uint32_t mbits;   // mismatch at query position j
    uint32_t ibits;   // insertion (base only in the gRNA) at query position j
    uint32_t dA, dB;  // 2-bit count of deletions immediately BEFORE query position j (bit planes)
    int mm, gaps;
};

__device__ inline int unit_count(const GeneralCtx& c, const UnitDev& u, const AlnState& a, uint32_t ubits, bool has_dd) {
    int cnt = 0;
    if (!u.rvs && has_dd) {
        // forward tuple with 'dd' elements of their own (blast.py:114-116): walk the elements explicitly
        int extra = 0;
        for (int j = 0; j < c.L; ++j) extra += (int)(((a.dA >> j) & 1u) + 2u * ((a.dB >> j) & 1u)) >> 1;
        const int n = c.L + extra;
        const int start = u.start == MG_SLICE_NONE ? 0 : u.start;
        const int end = u.end == MG_SLICE_NONE ? c.L : u.end;
        int first, last, step;
        py_slice_bounds(start, end, n, first, last, step);
        int idx = 0;
        for (int j = 0; j < c.L; ++j) {
            const int k = (int)(((a.dA >> j) & 1u) + 2u * ((a.dB >> j) & 1u));
            for (int r = 0; r < (k >> 1); ++r, ++idx)
                if (py_slice_has(first, last, step, idx) && (u.chars & MG_CH_DEL)) cnt += 2;
            if (py_slice_has(first, last, step, idx)) {
                if ((u.chars & MG_CH_DEL) && (k & 1)) cnt += 1;
                if ((u.chars & MG_CH_MISMATCH) && ((a.mbits >> j) & 1u)) cnt += 1;
                if ((u.chars & MG_CH_INS) && ((a.ibits >> j) & 1u)) cnt += 1;
                if ((ubits >> j) & 1u) cnt += u.add_m + u.add_g;
            }
            ++idx;
        }
        return cnt;
    }
    const uint32_t m = u.mask;
    if (u.chars & MG_CH_MISMATCH) cnt += __popc(a.mbits & m);
    if (u.chars & MG_CH_INS) cnt += __popc(a.ibits & m);
    if (u.chars & MG_CH_DEL) {
        if (u.rvs) cnt += __popc((a.dA >> 1) & m) + 2 * __popc((a.dB >> 1) & m);   // joined to the previous position
        else cnt += __popc(a.dA & m) + 2 * __popc(a.dB & m);                       // joined to the next position
    }
    cnt += (u.add_m + u.add_g) * __popc(ubits & m);
    return cnt;
}

__device__ inline bool pattern_true(const GeneralCtx& c, const AlnState& a, uint32_t ubits) {
    const bool has_dd = a.dB != 0;
    unsigned long long stack = 0;   // bit stack
    int sp = 0;
    for (int k = 0; k < c.n_ops; ++k) {
        const int op = c.ops[k];
        if (op >= 0) {
            const UnitDev& u = c.units[op];
            const bool v = unit_count(c, u, a, ubits, has_dd) <= u.n;
            stack = (stack << 1) | (v ? 1ull : 0ull);
            ++sp;
        } else {
            const unsigned long long b = stack & 1ull, aa = (stack >> 1) & 1ull;
            stack >>= 2;
            stack = (stack << 1) | (op == MG_OP_AND ? (aa & b) : (aa | b));
            --sp;
        }
    }
    return (stack & 1ull) != 0;
}

__device__ inline bool pam_in_window(const GenomeView& gv, int strand, const PamSideDev& side, int64_t w0, int64_t w1) {
    // re.search(side, window): some alternative matches somewhere inside [w0, w1)
    if (side.n_alts == 0) return false;        // MINORg.py:2085-2086: an empty side never matches
    for (int64_t x = w0; x <= w1; ++x) {
        for (int a = 0; a < side.n_alts; ++a) {
            const int len = side.len[a];
            if (x + len > w1) continue;
            bool ok = true;
            for (int k = 0; k < len && ok; ++k) {
                const int b = tbase(gv, strand, x + k);
                ok = b < 4 && ((side.allow4[a][k] >> b) & 1);
            }
            if (ok) return true;
        }
    }
    return false;
}

// One complete alignment: query [qs,qe) x subject [s,t) (strand coordinates) inside contig [cs,ce).
__device__ inline bool alignment_problematic(const GenomeView& gv, const GeneralCtx& c, int strand, const AlnState& a,
                                             int qs, int qe, int64_t s, int64_t t, int64_t cs, int64_t ce) {
    const int L = c.L;
    const int u = L - (qe - qs);
    if (c.n_units == 0) {
        const int mq = c.M - a.mm, gq = c.Gp - a.gaps;                       // MINORg.py:2033-2038
        if (mq < 0 || gq < 0 || mq + gq - u < 0) return false;
    } else {
        const uint32_t all = L >= 32 ? 0xFFFFFFFFu : ((1u << L) - 1u);
        const uint32_t aligned = (qe >= 32 ? 0xFFFFFFFFu : ((1u << qe) - 1u)) & ~((1u << qs) - 1u);
        if (!pattern_true(c, a, all & ~aligned)) return false;
    }
    // position check on the forward-strand span
    const int64_t fa = strand > 0 ? s : gv.glen - t, fb = strand > 0 ? t : gv.glen - s;
    if (gv.n_masks && span_masked(gv, fa, fb)) return false;
    if (c.pamless) return true;
    const int gap_excess = c.Gp - a.gaps;                                    // MINORg.py:2073-2076
    const int pre_len = c.five_max + gap_excess + qs;
    const int post_len = c.three_max + gap_excess + (L - qe);
    {
        const int64_t w0 = (s - (pre_len > 0 ? pre_len : 0) < cs) ? cs : s - (pre_len > 0 ? pre_len : 0);
        if (pam_in_window(gv, strand, c.five, w0, s)) return true;
    }
    {
        const int64_t w1 = (t + (post_len > 0 ? post_len : 0) > ce) ? ce : t + (post_len > 0 ? post_len : 0);
        if (pam_in_window(gv, strand, c.three, t, w1)) return true;
    }
    return false;
}

// Sound pruning of a PARTIAL alignment (columns known from the 3' end back to query position j):
// all counters only grow when more columns are added or the 5' end is trimmed.
__device__ inline bool partial_may_succeed(const GeneralCtx& c, const AlnState& a, int qe) {
    if (a.gaps > c.Gp) return false;
    if (c.n_units == 0) {
        if (a.mm > c.M) return false;
        return (c.M - a.mm) + (c.Gp - a.gaps) - (c.L - qe) >= 0;
    }
    if (!c.prune_pattern) return true;
    const uint32_t all = c.L >= 32 ? 0xFFFFFFFFu : ((1u << c.L) - 1u);
    const uint32_t tail = all & ~(qe >= 32 ? 0xFFFFFFFFu : ((1u << qe) - 1u));   // known unaligned 3' positions
    return pattern_true(c, a, tail);
}

// Closed-form verifier for the no-pattern criterion with at most ONE gap character (the reference's default
// and the common --ot-gap 2 case).  With <= 1 gap an alignment ending (virtually) at e is: a stretch of the main
// diagonal D0; or a prefix on the neighbouring diagonal D-1 joined to D0 by a deletion before query position k;
// or a prefix on D+1 joined by an insertion AT position k.  Mismatch bit-masks of the three diagonals turn every
// (qs, qe, k) choice into two popcounts, instead of the column-by-column search below.
__device__ inline uint32_t bit_range(int a, int b) {      // bits [a, b), 0 <= a <= b <= 32
    const uint32_t hi = b >= 32 ? 0xFFFFFFFFu : ((1u << b) - 1u);
    const uint32_t lo = a >= 32 ? 0xFFFFFFFFu : ((1u << a) - 1u);
    return hi & ~lo;
}

// even bits of x (bit 2i -> bit i)
__device__ __forceinline__ uint32_t compress_even(uint32_t lo, uint32_t hi) {
    auto half = [](uint32_t v) {
        v &= 0x55555555u;
        v = (v | (v >> 1)) & 0x33333333u;
        v = (v | (v >> 2)) & 0x0F0F0F0Fu;
        v = (v | (v >> 4)) & 0x00FF00FFu;
        return (v | (v >> 8)) & 0x0000FFFFu;
    };
    return half(lo) | (half(hi) << 16);
}

// Mismatch / off-contig masks of the three diagonals a one-gap alignment ending at e can use, from ONE 32-base window
// of the strand's text (x0 = e-L-1 .. x0+31) and the packed query: bit j of m0 / mM / mP = query position j against
// subject e-L+j / e-L+j-1 / e-L+j+1 mismatches (invalid or off-contig bases mismatch), f* = that subject base is off
// the contig.  Needs L <= 30 and the window inside the buffers; the caller falls back to per-base fetches otherwise.
__device__ __forceinline__ void diagonal_masks_packed(const GenomeView& gv, int strand, int64_t x0, int L, int64_t cs, int64_t ce,
                                                      const uint4 qp, uint32_t& m0, uint32_t& mM, uint32_t& mP,
                                                      uint32_t& f0, uint32_t& fM, uint32_t& fP) {
    const int n = L + 2;
    uint32_t lo, hi, inv;
    if (strand > 0) {
        load_window(gv, x0, lo, hi, inv);
    } else {
        // strand position x0+i is forward base glen-1-x0-i: load the forward window that ENDS there and reverse it
        uint32_t flo, fhi, finv;
        load_window(gv, gv.glen - x0 - 32, flo, fhi, finv);
        auto rev2 = [](uint32_t v) { v = __brev(v); return ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1); };
        lo = ~rev2(fhi);
        hi = ~rev2(flo);
        inv = __brev(finv);
    }
    // window index i <-> strand position x0 + i; off-contig part of the window
    const int64_t a = cs - x0, b = ce - x0;
    const uint32_t on = bit_range((int)max((int64_t)0, min((int64_t)32, a)), (int)max((int64_t)0, min((int64_t)32, b)));
    const uint32_t off = ~on;
    const uint32_t all = bit_range(0, L);
    const uint32_t bad = inv | off;                        // window positions that mismatch everything
    const unsigned long long w = ((unsigned long long)hi << 32) | lo;
    const unsigned long long q = ((unsigned long long)qp.y << 32) | qp.x;
    auto diag = [&](int o) {                               // query position j against window index j + o
        const unsigned long long x = (w >> (2 * o)) ^ q;
        const unsigned long long d = x | (x >> 1);
        return (compress_even((uint32_t)d, (uint32_t)(d >> 32)) | (bad >> o) | qp.z) & all;
    };
    mM = diag(0); m0 = diag(1); mP = diag(2);
    fM = off & all; f0 = (off >> 1) & all; fP = (off >> 2) & all;
    (void)n;
}

__device__ inline bool verify_nopattern_one_gap(const GenomeView& gv, const GeneralCtx& c, const uint8_t* __restrict__ q,
                                                const uint4* __restrict__ qpacked, int strand, int64_t e, int64_t cs, int64_t ce) {
    const int L = c.L;
    const int64_t base = e - L;                            // D0: query position j <-> subject base + j
    uint32_t m0 = 0, mM = 0, mP = 0, f0 = 0, fM = 0, fP = 0;   // mismatch / off-contig masks per diagonal
    const int64_t x0 = base - 1;
    const int64_t fwd0 = strand > 0 ? x0 : gv.glen - x0 - 32;   // first forward base of the 32-base window
    if (qpacked != nullptr && L <= 30 && fwd0 >= 0 && fwd0 + 32 <= gv.glen) {
        diagonal_masks_packed(gv, strand, x0, L, cs, ce, *qpacked, m0, mM, mP, f0, fM, fP);
    } else {
        int prev = -1, cur = -1, next = -1;
        {
            const int64_t x1 = base;
            prev = (x0 >= cs && x0 < ce) ? tbase(gv, strand, x0) : -1;
            cur = (x1 >= cs && x1 < ce) ? tbase(gv, strand, x1) : -1;
        }
        for (int j = 0; j < L; ++j) {
            const int64_t x2 = base + j + 1;
            next = (x2 >= cs && x2 < ce) ? tbase(gv, strand, x2) : -1;
            const int qb = q[j];
            const uint32_t b = 1u << j;
            if (cur < 0) f0 |= b;
            if (prev < 0) fM |= b;
            if (next < 0) fP |= b;
            if (qb >= 4 || qb != cur) m0 |= b;
            if (qb >= 4 || qb != prev) mM |= b;
            if (qb >= 4 || qb != next) mP |= b;
            prev = cur;
            cur = next;
        }
    }
    const int M = c.M, Gp = c.Gp, B = M + Gp;
    {   // Early reject (most prefilter survivors are unit-cost edit-distance hits with 2+ gaps, or ungapped hits with
        // M+1 mismatches in the middle).  Necessary conditions, ignoring contig edges, masks and PAM:
        //  - without a gap some trim [qs,qe) must have mm <= M and mm + u <= B;
        //  - with one gap, trimming a position adds 1 to u and removes at most 1 mismatch, so every trim costs at least
        //    the full-length alignment through the same gap position k: 1 + mm_full(k) <= B (which implies mm <= M).
        bool feasible = false;
        for (int qe = L; qe >= 1 && L - qe <= B && !feasible; --qe)
            for (int qs = 0; qs < qe && L - qe + qs <= B; ++qs) {
                const int mm = __popc(m0 & bit_range(qs, qe));
                if (mm <= M && mm + L - qe + qs <= B) { feasible = true; break; }
            }
        if (!feasible && Gp >= 1) {
            // del(k) = popc(mM & [0,k)) + popc(m0 & [k,L));  ins(k) = popc(mP & [0,k)) + popc(m0 & [k+1,L));  k = 1..L-1
            int del = (int)(mM & 1u) + __popc(m0 & bit_range(1, L));
            int ins = (int)(mP & 1u) + __popc(m0 & bit_range(2, L));
            int best = min(del, ins);
            for (int k = 1; k < L - 1; ++k) {
                del += (int)((mM >> k) & 1u) - (int)((m0 >> k) & 1u);
                ins += (int)((mP >> k) & 1u) - (int)((m0 >> (k + 1)) & 1u);
                best = min(best, min(del, ins));
            }
            feasible = 1 + best <= B;
        }
        if (!feasible) return false;
    }
    const bool edge = (f0 | fM | fP) != 0u;                // some base of the three diagonals is off the contig
    for (int qe = L; qe >= 1 && L - qe <= B; --qe) {
        const int trim3 = L - qe;
        for (int qs = 0; qs < qe && trim3 + qs <= B; ++qs) {          // no gap
            const uint32_t r = bit_range(qs, qe);
            if (f0 & r) continue;
            const int mm = __popc(m0 & r);
            if (mm > M || mm + trim3 + qs > B) continue;
            const AlnState a{m0 & r, 0u, 0u, 0u, mm, 0};
            if (alignment_problematic(gv, c, strand, a, qs, qe, base + qs, base + qe, cs, ce)) return true;
        }
        if (Gp < 1) continue;
        for (int qs = 0; qs < qe && trim3 + qs + 1 <= B; ++qs) {
            const int u = trim3 + qs;
            const int T = min(M, B - 1 - u);                            // most mismatches a one-gap alignment may have here
            if (!edge) {
                // the mismatch count moves by one bit of each diagonal per step of the gap position
                int mm = (int)((mM >> qs) & 1u) + __popc(m0 & bit_range(qs + 1, qe));
                for (int k = qs + 1; k <= qe - 1; ++k) {                // deletion before query position k
                    if (mm <= T) {
                        const AlnState a{(mM & bit_range(qs, k)) | (m0 & bit_range(k, qe)), 0u, 1u << k, 0u, mm, 1};
                        if (alignment_problematic(gv, c, strand, a, qs, qe, base - 1 + qs, base + qe, cs, ce)) return true;
                    }
                    mm += (int)((mM >> k) & 1u) - (int)((m0 >> k) & 1u);
                }
                mm = (int)((mP >> qs) & 1u) + __popc(m0 & bit_range(qs + 2, qe));
                for (int k = qs + 1; k <= qe - 2; ++k) {                // insertion at query position k
                    if (mm <= T) {
                        const AlnState a{(mP & bit_range(qs, k)) | (m0 & bit_range(k + 1, qe)), 1u << k, 0u, 0u, mm, 1};
                        if (alignment_problematic(gv, c, strand, a, qs, qe, base + 1 + qs, base + qe, cs, ce)) return true;
                    }
                    mm += (int)((mP >> k) & 1u) - (int)((m0 >> (k + 1)) & 1u);
                }
                continue;
            }
            for (int k = qs + 1; k <= qe - 1; ++k) {                    // deletion before query position k
                const uint32_t rp = bit_range(qs, k), rs = bit_range(k, qe);
                if ((fM & rp) | (f0 & rs)) continue;
                const uint32_t mb = (mM & rp) | (m0 & rs);
                const int mm = __popc(mb);
                if (mm > T) continue;
                const AlnState a{mb, 0u, 1u << k, 0u, mm, 1};
                if (alignment_problematic(gv, c, strand, a, qs, qe, base - 1 + qs, base + qe, cs, ce)) return true;
            }
            for (int k = qs + 1; k <= qe - 2; ++k) {                    // insertion at query position k
                const uint32_t rp = bit_range(qs, k), rs = bit_range(k + 1, qe);
                if ((fP & rp) | (f0 & rs)) continue;
                const uint32_t mb = (mP & rp) | (m0 & rs);
                const int mm = __popc(mb);
                if (mm > T) continue;
                const AlnState a{mb, 1u << k, 0u, 0u, mm, 1};
                if (alignment_problematic(gv, c, strand, a, qs, qe, base + 1 + qs, base + qe, cs, ce)) return true;
            }
        }
    }
    return false;
}

// Early reject for the no-pattern criterion with 2..MG_MAX_GAPS gaps (the exhaustive walk below costs ~10^5 operations on
// a candidate that has no alignment, and most unit-cost edit-distance survivors have none within the gap budget).
// Lower bound: a trimmed alignment with (mm, g gaps, u unaligned) extends, with the same gaps, to a full-length
// alignment ending at e with mm' <= mm + u mismatches (off-contig or invalid subject bases count as mismatches), so
// min over g <= Gp of (fewest mismatches of a full-length alignment with exactly g gaps) + g must be <= B, where B is
// M + Gp for the no-pattern criterion or the host-proved edit bound of a pattern (ot_regex.edit_bound).
// f[d][g] = fewest mismatches of query [j, L) against the subject ending at e when query position j sits on diagonal
// d (subject position e - L + j + d) after g gap characters; the walk goes from j = L down to 0.
__device__ inline bool gapped_cost_exceeds(const GenomeView& gv, const GeneralCtx& c, const uint8_t* __restrict__ q,
                                           int strand, int64_t e, int64_t cs, int64_t ce, int B) {
    constexpr int GM = MG_MAX_GAPS, W = 2 * GM + 1, INF = 1 << 20;
    const int L = c.L, Gp = c.Gp;
    const int64_t base = e - L;
    int f[W][GM + 1];
#pragma unroll
    for (int d = 0; d < W; ++d)
#pragma unroll
        for (int g = 0; g <= GM; ++g) f[d][g] = INF;
    f[GM][0] = 0;
    int win[W];                                             // subject bases at base + j - 1 + d, d = -GM..GM (-1: off the contig)
#pragma unroll
    for (int d = 0; d < W; ++d) {
        const int64_t x = base + L - 1 + (d - GM);
        win[d] = (x >= cs && x < ce) ? tbase(gv, strand, x) : -1;
    }
    for (int j = L; j >= 1; --j) {
        // deletions at query position j (subject base consumed alone): diagonal d -> d - 1, one more gap
#pragma unroll
        for (int d = W - 1; d >= 1; --d)
#pragma unroll
            for (int g = 0; g < GM; ++g)
                if (g < Gp) f[d - 1][g + 1] = min(f[d - 1][g + 1], f[d][g]);
        // query position j - 1: paired with the subject base of its diagonal, or inserted (diagonal d -> d + 1)
        const int qb = q[j - 1];
        int nf[W][GM + 1];
#pragma unroll
        for (int d = 0; d < W; ++d)
#pragma unroll
            for (int g = 0; g <= GM; ++g) {
                const int mis = (qb >= 4 || qb != win[d]) ? 1 : 0;
                int v = f[d][g] + mis;
                if (d >= 1 && g >= 1 && g <= Gp) v = min(v, f[d - 1][g - 1]);
                nf[d][g] = min(v, INF);
            }
#pragma unroll
        for (int d = 0; d < W; ++d)
#pragma unroll
            for (int g = 0; g <= GM; ++g) f[d][g] = nf[d][g];
        // slide the subject window one base to the left
#pragma unroll
        for (int d = W - 1; d >= 1; --d) win[d] = win[d - 1];
        const int64_t x = base + j - 2 - GM;
        win[0] = (x >= cs && x < ce) ? tbase(gv, strand, x) : -1;
    }
    int best = INF;
#pragma unroll
    for (int d = 0; d < W; ++d)
#pragma unroll
        for (int g = 0; g <= GM; ++g)
            if (g <= Gp) best = min(best, f[d][g] + g);
    return best > B;
}

// cold call site of the exact no-pattern search below (keeps the unrolled DP small)
static __device__ __noinline__ bool nopattern_end_ok(const GenomeView& gv, const GeneralCtx& c, int strand, int mm, int gaps,
                                                     int qs, int qe, int64_t s, int64_t t, int64_t cs, int64_t ce) {
    const AlnState a{0u, 0u, 0u, 0u, mm, gaps};
    return alignment_problematic(gv, c, strand, a, qs, qe, s, t, cs, ce);
}

// Exact verifier for the no-pattern criterion with 2..MG_MAX_GAPS gaps: the criterion (MINORg.py:2018-2044), the position
// check and the PAM windows only depend on (qs, qe, subject span, #mismatches, #gap characters), so instead of walking
// every alignment it is enough to know, for each 3' trim qe and each (qs, diagonal, #gaps), the FEWEST mismatches of an
// E1 alignment (first and last column a base pair, gap characters strictly inside, everything on the contig) - one
// banded dynamic programme per qe over f[diagonal][gaps], walked from the 3' end like the search it replaces.
__device__ inline bool verify_nopattern_gaps(const GenomeView& gv, const GeneralCtx& c, const uint8_t* __restrict__ q,
                                             int strand, int64_t e, int64_t cs, int64_t ce) {
    constexpr int GM = MG_MAX_GAPS, W = 2 * GM + 1, INF = 1 << 20;
    const int L = c.L, M = c.M, Gp = c.Gp, B = M + Gp;
    const int64_t base = e - L;
    int sub[MG_MAX_GRNA_LEN + 2 * GM + 2];                  // sub[k] = subject base at base - GM - 1 + k (-1: off the contig)
    for (int k = 0; k <= L + 2 * GM; ++k) {
        const int64_t x = base - GM - 1 + k;
        sub[k] = (x >= cs && x < ce) ? tbase(gv, strand, x) : -1;
    }
    for (int qe = L; qe >= 1 && L - qe <= B; --qe) {
        const int64_t t_end = e - (L - qe);
        if (t_end - 1 < cs || t_end > ce) continue;          // the last column must be on the contig
        const int trim3 = L - qe;
        int f[W][GM + 1];                                    // fewest mismatches of query [j, qe); diagonal d - GM; g gaps
#pragma unroll
        for (int d = 0; d < W; ++d)
#pragma unroll
            for (int g = 0; g <= GM; ++g) f[d][g] = INF;
        f[GM][0] = 0;
        for (int j = qe; j >= 1; --j) {
            if (j < qe) {
                // deletion before query position j: consumes subject base + j - 1 + delta; a pair must still fit on the contig
#pragma unroll
                for (int d = W - 1; d >= 1; --d) {
                    const bool room = base + j + (d - GM) - 2 >= cs;
#pragma unroll
                    for (int g = 0; g < GM; ++g)
                        if (g < Gp && room) f[d - 1][g + 1] = min(f[d - 1][g + 1], f[d][g]);
                }
            }
            const int qb = q[j - 1];
            const int qs = j - 1, u = trim3 + qs;
            int nf[W][GM + 1];
#pragma unroll
            for (int d = 0; d < W; ++d) {
                const int sb = sub[j + d];                   // subject base + j - 1 + (d - GM)
                const int mis = (qb >= 4 || qb != sb) ? 1 : 0;
#pragma unroll
                for (int g = 0; g <= GM; ++g) {
                    int v = (sb >= 0 && f[d][g] < INF) ? f[d][g] + mis : INF;     // base pair
                    if (v > M) v = INF;                      // mismatches only grow
                    if (v < INF && g <= Gp && (M - v) + (Gp - g) - u >= 0 &&
                        nopattern_end_ok(gv, c, strand, v, g, qs, qe, base + j - 1 + (d - GM), t_end, cs, ce)) return true;
                    // insertion of query position j - 1 (not as the last column; a pair must follow)
                    if (j < qe && j >= 2 && d >= 1 && g >= 1 && g <= Gp) v = min(v, f[d - 1][g - 1]);
                    nf[d][g] = v;
                }
            }
#pragma unroll
            for (int d = 0; d < W; ++d)
#pragma unroll
                for (int g = 0; g <= GM; ++g) f[d][g] = nf[d][g];
        }
    }
    return false;
}

constexpr int kMaxDepth = MG_MAX_GRNA_LEN + MG_MAX_GAPS + 2;

struct Frame {
    AlnState a;
    int j;          // query positions [j, qe) are consumed
    int64_t t;      // subject positions [t, t_end) are consumed
    int next;       // next choice to try: 0 pair, 1 insertion, 2 deletion
    int last_pair;  // the most recent (leftmost) column is a base pair
};

// gRNA `q` (codes in gRNA orientation, 5'->3'), strand, virtual end e in strand coordinates.
static __device__ __noinline__ bool verify_anchor(const GenomeView& gv, const GeneralCtx& c, const uint8_t* __restrict__ q,
                                           const uint4* __restrict__ qpacked, int strand, int64_t e) {
    const int L = c.L;
    // the contig (strand coordinates) that intersects [e - L - Gp, e)
    const int64_t lo = e - L - c.Gp, hi = e;
    const int64_t flo = strand > 0 ? lo : gv.glen - hi, fhi = strand > 0 ? hi : gv.glen - lo;
    int a0 = 0, a1 = gv.n_contigs - 1, ci = -1;
    while (a0 <= a1) {
        const int mid = (a0 + a1) >> 1;
        if (gv.cstart[mid] < fhi) { ci = mid; a0 = mid + 1; } else a1 = mid - 1;
    }
    if (ci < 0 || gv.cend[ci] <= flo) return false;
    const int64_t cs = strand > 0 ? gv.cstart[ci] : gv.glen - gv.cend[ci];
    const int64_t ce = strand > 0 ? gv.cend[ci] : gv.glen - gv.cstart[ci];

    if (c.n_units == 0 && c.Gp <= 1) return verify_nopattern_one_gap(gv, c, q, qpacked, strand, e, cs, ce);
    if (c.n_units == 0) {
        if (gapped_cost_exceeds(gv, c, q, strand, e, cs, ce, c.M + c.Gp)) return false;
        if (c.Gp <= MG_MAX_GAPS) return verify_nopattern_gaps(gv, c, q, strand, e, cs, ce);
    }

    if (c.edit_bound >= 0 && gapped_cost_exceeds(gv, c, q, strand, e, cs, ce, c.edit_bound)) return false;

    Frame st[kMaxDepth];
    for (int qe = L; qe >= 1; --qe) {
        const int64_t t_end = e - (L - qe);
        if (t_end - 1 < cs || t_end > ce) continue;          // the last column must be on the contig
        int sp = 0;
        st[0].a = AlnState{0u, 0u, 0u, 0u, 0, 0};
        st[0].j = qe; st[0].t = t_end; st[0].next = 0; st[0].last_pair = 0;
        {   // cheap bound on trimming before any work: the 3' trim alone may already break the budget
            if (!partial_may_succeed(c, st[0].a, qe)) continue;
        }
        while (sp >= 0) {
            Frame& f = st[sp];
            const int choice = f.next++;
            if (choice > 2) { --sp; continue; }
            const bool first_col = (f.j == qe && f.t == t_end);
            Frame n = f;
            n.next = 0;
            if (choice == 0) {
                if (f.j < 1 || f.t - 1 < cs) continue;
                const int qb = q[f.j - 1];
                const int sb = tbase(gv, strand, f.t - 1);
                const bool same = (qb == sb) && qb < 4;
                n.j = f.j - 1; n.t = f.t - 1; n.last_pair = 1;
                if (!same) { n.a.mbits |= 1u << (f.j - 1); n.a.mm++; }
                if (!partial_may_succeed(c, n.a, qe)) continue;
                if (alignment_problematic(gv, c, strand, n.a, n.j, qe, n.t, t_end, cs, ce)) return true;
            } else {
                if (first_col || f.a.gaps >= c.Gp) continue;
                if (choice == 1) {           // insertion: consumes query position j-1 only; a pair must follow
                    if (f.j < 2) continue;
                    n.j = f.j - 1; n.last_pair = 0;
                    n.a.ibits |= 1u << (f.j - 1); n.a.gaps++;
                } else {                     // deletion: consumes subject t-1 only, sits before query position j
                    if (f.j < 1 || f.t - 2 < cs) continue;
                    n.t = f.t - 1; n.last_pair = 0;
                    const uint32_t bit = 1u << f.j;
                    if (f.j < 32) {
                        if (n.a.dA & bit) { n.a.dA &= ~bit; n.a.dB ^= bit; if (!(n.a.dB & bit)) { /* 4 deletions: beyond MG_MAX_GAPS */ } }
                        else n.a.dA |= bit;
                    }
                    n.a.gaps++;
                }
                if (!partial_may_succeed(c, n.a, qe)) continue;
            }
            if (sp + 1 < kMaxDepth) st[++sp] = n;
        }
    }
    return false;
}

}  // namespace mg

namespace mg {

// Hand a prefilter survivor to the verifier pass.  The counter keeps counting past the capacity: the host sees
// the overflow, learns the survivor density from it and re-runs the range in slices that fit.
__device__ __forceinline__ void emit_candidate(const GeneralCtx& c, int q, int64_t e, int32_t tag = 0) {
    const unsigned long long slot = atomicAdd(c.cand_count, 1ull);
    if (slot < c.cand_cap) c.cand[slot] = Cand{q, tag, e};
}

}  // namespace mg
