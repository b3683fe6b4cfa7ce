/* CPU ORACLE (test infrastructure, NOT product code) - C restatement of the exhaustive ungapped screen.
 *
 * Same semantics as oracle/minorg_oracle.py:screen_hamming, i.e. reduction (2) of SURVEY.md 8c of the
 * reference's criterion MINORg._is_offtarget_aln_nopattern (minorg/MINORg.py:2018-2044) with ot_gap<=1,
 * ot_pamless, no pattern, applied to EVERY window instead of BLAST's HSPs, plus the position check
 * MINORg._is_offtarget_pos (MINORg.py:1994-2000; functions.py:310-325):
 *   a window (contig, start p in [-L+1, n), strand) counts for gRNA q iff
 *     #{positions off the contig} + #{on-contig positions where q and the subject differ or either is not
 *     ACGT} <= max_mismatch, and its on-contig span is not inside one masked interval.
 * Parity pinning: tests/test_oracle_c.py checks this file against the Python oracle on the golden screen
 * cases (whose verdicts come from the reference's own predicates).
 *
 * Also the "port" CPU baseline timed by bench.py: 2-bit windows, 64-bit XOR + popcount, pthreads over
 * window ranges.   Build: gcc -O3 -march=native -shared -fPIC -pthread screen_oracle.c -o libscreen_oracle.so
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline int code_of(char c) {
    switch (c) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        default: return 4;
    }
}

typedef struct {
    const uint8_t* codes; /* contig codes, 0..3 / 4 */
    int64_t n;            /* contig length */
    int64_t p0, p1;       /* window starts [p0, p1) handled by this thread */
    const uint64_t* qbits; /* [nq] 2-bit packed queries */
    const uint64_t* qinv;  /* [nq] invalid positions of each query at the even bits */
    int nq, L, max_mm;
    const int64_t* ms; const int64_t* me; int n_masks; /* masks of this contig */
    uint32_t* counts;     /* [nq] thread-private */
} job_t;

static int span_masked(const job_t* j, int64_t a, int64_t b) {
    for (int i = 0; i < j->n_masks; ++i) {
        int64_t lo = j->ms[i] < j->me[i] ? j->ms[i] : j->me[i];
        int64_t hi = j->ms[i] < j->me[i] ? j->me[i] : j->ms[i];
        if (a >= lo && b <= hi) return 1;
    }
    return 0;
}

static void* run(void* arg) {
    job_t* j = (job_t*)arg;
    const int L = j->L;
    const uint64_t even = 0x5555555555555555ULL;
    const uint64_t lmask = (L >= 32) ? even : (((1ULL << (2 * L)) - 1) & even);
    for (int64_t p = j->p0; p < j->p1; ++p) {
        uint64_t w = 0, winv = 0;
        for (int k = 0; k < L; ++k) {
            const int64_t i = p + k;
            const int c = (i < 0 || i >= j->n) ? 4 : j->codes[i];
            if (c == 4) winv |= 1ULL << (2 * k); else w |= (uint64_t)c << (2 * k);
        }
        int masked = -1;
        for (int q = 0; q < j->nq; ++q) {
            const uint64_t x = w ^ j->qbits[q];
            const uint64_t d = ((x | (x >> 1)) & lmask) | winv | j->qinv[q];
            if (__builtin_popcountll(d) > j->max_mm) continue;
            if (masked < 0) {
                const int64_t a = p < 0 ? 0 : p, b = (p + L > j->n) ? j->n : p + L;
                masked = (a >= b) ? 1 : span_masked(j, a, b);
            }
            if (!masked) j->counts[q]++;
        }
    }
    return NULL;
}

/* counts_out[i] = number of (window, strand) pairs within budget and outside masks, over all contigs.
 * win_lo / win_hi (NULL = every window): per contig, only the windows whose START lies in [win_lo[c], win_hi[c])
 * are counted (clipped to [-L+1, len)) - the C side of mg_screen's window shards, used by bench.py's parity check on
 * a slice of a full-size genome. */
int oracle_screen_hamming_win(const char* seq, const int64_t* offsets, int n_contigs,
                              const int32_t* mask_contig, const int64_t* mask_start, const int64_t* mask_end, int n_masks,
                              const char* grna, int n, int L, int max_mm, int n_threads, uint32_t* counts_out,
                              const int64_t* win_lo, const int64_t* win_hi);

int oracle_screen_hamming(const char* seq, const int64_t* offsets, int n_contigs,
                          const int32_t* mask_contig, const int64_t* mask_start, const int64_t* mask_end, int n_masks,
                          const char* grna, int n, int L, int max_mm, int n_threads, uint32_t* counts_out) {
    return oracle_screen_hamming_win(seq, offsets, n_contigs, mask_contig, mask_start, mask_end, n_masks, grna, n, L, max_mm,
                                     n_threads, counts_out, NULL, NULL);
}

int oracle_screen_hamming_win(const char* seq, const int64_t* offsets, int n_contigs,
                              const int32_t* mask_contig, const int64_t* mask_start, const int64_t* mask_end, int n_masks,
                              const char* grna, int n, int L, int max_mm, int n_threads, uint32_t* counts_out,
                              const int64_t* win_lo, const int64_t* win_hi) {
    if (L < 1 || L > 32 || n < 0 || n_threads < 1) return -2;
    const int nq = 2 * n;
    uint64_t* qbits = (uint64_t*)calloc((size_t)nq + 1, 8);
    uint64_t* qinv = (uint64_t*)calloc((size_t)nq + 1, 8);
    for (int i = 0; i < n; ++i) {
        for (int k = 0; k < L; ++k) {
            const int c = code_of(grna[(size_t)i * L + k]);
            if (c == 4) { qinv[i] |= 1ULL << (2 * k); qinv[n + i] |= 1ULL << (2 * (L - 1 - k)); }
            else { qbits[i] |= (uint64_t)c << (2 * k); qbits[n + i] |= (uint64_t)(3 - c) << (2 * (L - 1 - k)); }
        }
    }
    memset(counts_out, 0, sizeof(uint32_t) * (size_t)n);
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)n_threads);
    job_t* jobs = (job_t*)malloc(sizeof(job_t) * (size_t)n_threads);
    uint32_t* priv = (uint32_t*)calloc((size_t)n_threads * (size_t)(nq + 1), 4);
    for (int c = 0; c < n_contigs; ++c) {
        const int64_t len = offsets[c + 1] - offsets[c];
        int nm = 0;
        int64_t* ms = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n_masks + 1));
        int64_t* me = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n_masks + 1));
        for (int i = 0; i < n_masks; ++i) if (mask_contig[i] == c) { ms[nm] = mask_start[i]; me[nm] = mask_end[i]; ++nm; }
        int64_t first = -(int64_t)L + 1, last = len; /* window starts [first, last) */
        if (win_lo != NULL && win_hi != NULL) {
            if (win_lo[c] > first) first = win_lo[c];
            if (win_hi[c] < last) last = win_hi[c];
            if (last <= first) { free(ms); free(me); continue; }
        }
        uint8_t* codes = (uint8_t*)malloc((size_t)len + 1);
        for (int64_t i = 0; i < len; ++i) codes[i] = (uint8_t)code_of(seq[offsets[c] + i]);
        const int64_t total = last - first;
        for (int t = 0; t < n_threads; ++t) {
            job_t* j = &jobs[t];
            j->codes = codes; j->n = len; j->qbits = qbits; j->qinv = qinv; j->nq = nq; j->L = L; j->max_mm = max_mm;
            j->ms = ms; j->me = me; j->n_masks = nm; j->counts = priv + (size_t)t * (size_t)(nq + 1);
            j->p0 = first + total * t / n_threads;
            j->p1 = first + total * (t + 1) / n_threads;
            pthread_create(&th[t], NULL, run, j);
        }
        for (int t = 0; t < n_threads; ++t) pthread_join(th[t], NULL);
        free(codes); free(ms); free(me);
    }
    for (int t = 0; t < n_threads; ++t)
        for (int i = 0; i < n; ++i)
            counts_out[i] += priv[(size_t)t * (size_t)(nq + 1) + i] + priv[(size_t)t * (size_t)(nq + 1) + n + i];
    free(priv); free(jobs); free(th); free(qbits); free(qinv);
    return 0;
}

/* ------------------------------------------------------------------------------------------------------------
 * Exhaustive no-pattern screen with gaps and trimming (universe E1 of SURVEY.md 8c), PAM-less: for every gRNA and
 * strand, the number of distinct ANCHORS (virtual end of the alignment = subject end + unaligned 3' positions) that
 * carry at least one alignment satisfying MINORg._is_offtarget_aln_nopattern (MINORg.py:2018-2044:
 * mm <= M, gaps <= Gp, mm + gaps + unaligned <= M + Gp) whose subject span is not inside one mask
 * (MINORg.py:1994-2000).  Plain forward depth-first enumeration from every (query start, subject start), pruned only
 * by the budget (all three counters are monotone).  This is what the device's general path reports per oriented
 * query for the same criteria (prefilter + verifier), so tests compare the 2N counts one to one.
 * Pinned indirectly: on the golden cases its fail flags equal the verdicts of the reference's predicates
 * (tests/test_oracle_c.py). */
typedef struct {
    const uint8_t* T; int64_t n;       /* strand text codes */
    const uint8_t* q; int L, M, Gp, B;
    const int64_t* ms; const int64_t* me; int n_masks; int strand; /* masks in forward coordinates */
    uint8_t* ends;                     /* bitmap over virtual ends 0 .. n + L */
    int qs; int64_t s;
    /* optional PAM proximity check (MINORg._has_pam, MINORg.py:2063-2087); pamless = 1 skips it */
    int pamless, n5, n3, max5, max3;
    const int32_t* len5; const uint8_t* allow5; const int32_t* len3; const uint8_t* allow3; /* allow[alt*16 + k] = ACGT bit mask */
} gctx_t;

static int g_pam_in(const gctx_t* c, int n_alts, const int32_t* len, const uint8_t* allow, int64_t w0, int64_t w1) {
    if (n_alts == 0) return 0;                      /* an empty PAM side never matches (MINORg.py:2085-2086) */
    for (int64_t x = w0; x <= w1; ++x)
        for (int a = 0; a < n_alts; ++a) {
            if (x + len[a] > w1) continue;
            int ok = 1;
            for (int k = 0; k < len[a] && ok; ++k) { const int b = c->T[x + k]; ok = b < 4 && ((allow[a * 16 + k] >> b) & 1); }
            if (ok) return 1;
        }
    return 0;
}

static int g_has_pam(const gctx_t* c, int qe, int gaps, int64_t s, int64_t t) {
    const int gap_excess = c->Gp - gaps;
    int pre = c->max5 + gap_excess + c->qs, post = c->max3 + gap_excess + (c->L - qe);
    if (pre < 0) pre = 0;
    if (post < 0) post = 0;
    const int64_t w0 = s - pre < 0 ? 0 : s - pre, w1 = t + post > c->n ? c->n : t + post;   /* clamped to the contig */
    return g_pam_in(c, c->n5, c->len5, c->allow5, w0, s) || g_pam_in(c, c->n3, c->len3, c->allow3, t, w1);
}

static int g_masked(const gctx_t* c, int64_t a, int64_t b) {     /* strand span [a,b) -> forward span */
    int64_t fa = c->strand > 0 ? a : c->n - b, fb = c->strand > 0 ? b : c->n - a;
    for (int i = 0; i < c->n_masks; ++i) {
        int64_t lo = c->ms[i] < c->me[i] ? c->ms[i] : c->me[i], hi = c->ms[i] < c->me[i] ? c->me[i] : c->ms[i];
        if (fa >= lo && fb <= hi) return 1;
    }
    return 0;
}

static void g_dfs(gctx_t* c, int j, int64_t t, int gaps, int mm, int ncols) {
    if (mm + gaps + c->qs > c->B) return;
    if (j < c->L && t < c->n) {
        const int qb = c->q[j], sb = c->T[t];
        const int mm2 = mm + ((qb == sb && qb < 4) ? 0 : 1);
        if (mm2 <= c->M) {
            const int qe = j + 1, u = c->qs + c->L - qe;
            if (mm2 + gaps + u <= c->B && !g_masked(c, c->s, t + 1) && (c->pamless || g_has_pam(c, qe, gaps, c->s, t + 1))) {
                const int64_t e = (t + 1) + (c->L - qe);
                c->ends[e >> 3] |= (uint8_t)(1u << (e & 7));
            }
            g_dfs(c, j + 1, t + 1, gaps, mm2, ncols + 1);
        }
    }
    if (ncols > 0 && gaps < c->Gp) {
        if (j < c->L) g_dfs(c, j + 1, t, gaps + 1, mm, ncols + 1);
        if (t < c->n) g_dfs(c, j, t + 1, gaps + 1, mm, ncols + 1);
    }
}

typedef struct {
    const uint8_t* fwd; const uint8_t* rev; int64_t n;
    const uint8_t* qcodes; int nq, L, M, Gp;
    const int64_t* ms; const int64_t* me; int n_masks;
    int q0, q1; uint32_t* counts; /* [2*nq] shared, disjoint slots per thread */
    int pamless, n5, n3, max5, max3;
    const int32_t* len5; const uint8_t* allow5; const int32_t* len3; const uint8_t* allow3;
    int64_t wlo, whi;              /* only anchors owned by the windows that START in [wlo, whi) (forward contig coordinates) count */
} gjob_t;

/* Anchor range [*e0, *e1) (strand coordinates) owned by the forward windows [wlo, whi): a window [p, p+L) owns the '+' anchor
 * p + L and the '-' anchor n - p (its end in reverse-complement coordinates) - mg_screen's ownership rule for shards. */
static void anchor_range(int strand, int64_t n, int L, int64_t wlo, int64_t whi, int64_t* e0, int64_t* e1) {
    if (strand > 0) { *e0 = wlo + L; *e1 = whi + L; }
    else { *e0 = n - whi + 1; *e1 = n - wlo + 1; }
    if (*e0 < 0) *e0 = 0;
    if (*e1 > n + L + 1) *e1 = n + L + 1;
}

static void* g_run(void* arg) {
    gjob_t* j = (gjob_t*)arg;
    const size_t nbytes = (size_t)(j->n + j->L + 16) / 8 + 2;
    uint8_t* ends = (uint8_t*)malloc(nbytes);
    for (int qi = j->q0; qi < j->q1; ++qi) {
        for (int strand = 1; strand >= -1; strand -= 2) {
            memset(ends, 0, nbytes);
            gctx_t c;
            c.T = strand > 0 ? j->fwd : j->rev; c.n = j->n; c.q = j->qcodes + (size_t)qi * j->L;
            c.L = j->L; c.M = j->M; c.Gp = j->Gp; c.B = j->M + j->Gp;
            c.ms = j->ms; c.me = j->me; c.n_masks = j->n_masks; c.strand = strand; c.ends = ends;
            c.pamless = j->pamless; c.n5 = j->n5; c.n3 = j->n3; c.max5 = j->max5; c.max3 = j->max3;
            c.len5 = j->len5; c.allow5 = j->allow5; c.len3 = j->len3; c.allow3 = j->allow3;
            int64_t e0, e1;
            anchor_range(strand, j->n, j->L, j->wlo, j->whi, &e0, &e1);
            if (e1 <= e0) continue;
            /* an alignment that ends (virtually) at e starts no earlier than e - L - Gp */
            const int64_t s_lo = e0 - j->L - j->Gp - 1 < 0 ? 0 : e0 - j->L - j->Gp - 1, s_hi = e1 < j->n ? e1 : j->n;
            for (int64_t s = s_lo; s < s_hi; ++s)
                for (int qs = 0; qs < j->L && qs <= c.B; ++qs) { c.qs = qs; c.s = s; g_dfs(&c, qs, s, 0, 0, 0); }
            uint32_t cnt = 0;
            for (int64_t e = e0; e < e1; ++e) cnt += (ends[e >> 3] >> (e & 7)) & 1u;
            j->counts[strand > 0 ? qi : j->nq + qi] += cnt;
        }
    }
    free(ends);
    return NULL;
}

/* counts_out[2n]: [i] = '+' strand anchors of gRNA i, [n+i] = '-' strand anchors. */
int oracle_screen_nopattern_pam(const char* seq, const int64_t* offsets, int n_contigs,
                                const int32_t* mask_contig, const int64_t* mask_start, const int64_t* mask_end, int n_masks,
                                const char* grna, int n, int L, int max_mm, int max_gap, int n_threads, uint32_t* counts_out,
                                int pamless, int n5, const int32_t* len5, const uint8_t* allow5, int max5,
                                int n3, const int32_t* len3, const uint8_t* allow3, int max3);

int oracle_screen_nopattern(const char* seq, const int64_t* offsets, int n_contigs,
                            const int32_t* mask_contig, const int64_t* mask_start, const int64_t* mask_end, int n_masks,
                            const char* grna, int n, int L, int max_mm, int max_gap, int n_threads, uint32_t* counts_out) {
    return oracle_screen_nopattern_pam(seq, offsets, n_contigs, mask_contig, mask_start, mask_end, n_masks, grna, n, L,
                                       max_mm, max_gap, n_threads, counts_out, 1, 0, NULL, NULL, 0, 0, NULL, NULL, 0);
}

/* Same with the PAM proximity requirement: a PAM side = n alternatives of fixed length, position k of alternative a
 * allows the bases whose bit is set in allow[a*16 + k] (bit 0..3 = A,C,G,T); max5/max3 = PAM.five/three_prime_maxlen(). */
int oracle_screen_nopattern_win(const char* seq, const int64_t* offsets, int n_contigs,
                                const int32_t* mask_contig, const int64_t* mask_start, const int64_t* mask_end, int n_masks,
                                const char* grna, int n, int L, int max_mm, int max_gap, int n_threads, uint32_t* counts_out,
                                int pamless, int n5, const int32_t* len5, const uint8_t* allow5, int max5,
                                int n3, const int32_t* len3, const uint8_t* allow3, int max3,
                                const int64_t* win_lo, const int64_t* win_hi);

int oracle_screen_nopattern_pam(const char* seq, const int64_t* offsets, int n_contigs,
                                const int32_t* mask_contig, const int64_t* mask_start, const int64_t* mask_end, int n_masks,
                                const char* grna, int n, int L, int max_mm, int max_gap, int n_threads, uint32_t* counts_out,
                                int pamless, int n5, const int32_t* len5, const uint8_t* allow5, int max5,
                                int n3, const int32_t* len3, const uint8_t* allow3, int max3) {
    return oracle_screen_nopattern_win(seq, offsets, n_contigs, mask_contig, mask_start, mask_end, n_masks, grna, n, L, max_mm, max_gap,
                                       n_threads, counts_out, pamless, n5, len5, allow5, max5, n3, len3, allow3, max3, NULL, NULL);
}

/* win_lo / win_hi (NULL = everything): per contig, only the anchors owned by the windows that start in [win_lo[c], win_hi[c])
 * are counted - the C side of mg_screen(win_begin, win_end) for the gapped criteria (bench.py --gapped parity slice). */
int oracle_screen_nopattern_win(const char* seq, const int64_t* offsets, int n_contigs,
                                const int32_t* mask_contig, const int64_t* mask_start, const int64_t* mask_end, int n_masks,
                                const char* grna, int n, int L, int max_mm, int max_gap, int n_threads, uint32_t* counts_out,
                                int pamless, int n5, const int32_t* len5, const uint8_t* allow5, int max5,
                                int n3, const int32_t* len3, const uint8_t* allow3, int max3,
                                const int64_t* win_lo, const int64_t* win_hi) {
    if (L < 1 || L > 32 || n < 0 || n_threads < 1 || max_gap < 0 || max_gap > 3) return -2;
    uint8_t* qcodes = (uint8_t*)malloc((size_t)n * L + 1);
    for (size_t i = 0; i < (size_t)n * L; ++i) qcodes[i] = (uint8_t)code_of(grna[i]);
    memset(counts_out, 0, sizeof(uint32_t) * 2 * (size_t)n);
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)n_threads);
    gjob_t* jobs = (gjob_t*)malloc(sizeof(gjob_t) * (size_t)n_threads);
    for (int c = 0; c < n_contigs; ++c) {
        const int64_t len = offsets[c + 1] - offsets[c];
        if (len == 0) continue;
        uint8_t* fwd = (uint8_t*)malloc((size_t)len), *rev = (uint8_t*)malloc((size_t)len);
        for (int64_t i = 0; i < len; ++i) {
            const int k = code_of(seq[offsets[c] + i]);
            fwd[i] = (uint8_t)k;
            rev[len - 1 - i] = (uint8_t)(k == 4 ? 4 : 3 - k);
        }
        int nm = 0;
        int64_t* ms = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n_masks + 1)), *me = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n_masks + 1));
        for (int i = 0; i < n_masks; ++i) if (mask_contig[i] == c) { ms[nm] = mask_start[i]; me[nm] = mask_end[i]; ++nm; }
        for (int t = 0; t < n_threads; ++t) {
            gjob_t* j = &jobs[t];
            j->fwd = fwd; j->rev = rev; j->n = len; j->qcodes = qcodes; j->nq = n; j->L = L; j->M = max_mm; j->Gp = max_gap;
            j->ms = ms; j->me = me; j->n_masks = nm; j->q0 = (int)((int64_t)n * t / n_threads); j->q1 = (int)((int64_t)n * (t + 1) / n_threads);
            j->counts = counts_out;
            j->pamless = pamless; j->n5 = n5; j->n3 = n3; j->max5 = max5; j->max3 = max3;
            j->len5 = len5; j->allow5 = allow5; j->len3 = len3; j->allow3 = allow3;
            j->wlo = win_lo ? win_lo[c] : -(int64_t)L - 8; j->whi = win_hi ? win_hi[c] : len + 8;
            pthread_create(&th[t], NULL, g_run, j);
        }
        for (int t = 0; t < n_threads; ++t) pthread_join(th[t], NULL);
        free(fwd); free(rev); free(ms); free(me);
    }
    free(jobs); free(th); free(qcodes);
    return 0;
}

/* ------------------------------------------------------------------------------------------------------------
 * Exhaustive --ot-pattern screen (universe E1 of SURVEY.md 8c): per oriented query, the number of distinct ANCHORS
 * (virtual end = subject end + unaligned 3' positions) that carry at least one alignment for which the reference's
 * OffTargetExpression is true (ot_regex.py:124-356), whose span is not inside one mask (MINORg.py:1994-2000) and - unless
 * pamless - that has a PAM nearby (MINORg.py:2063-2087).  Everything is restated from the reference's text, on strings,
 * like the reference does it:
 *   - the pattern STRING is parsed here (ot_regex.py:295-356: ',' = AND, '|' = OR, strictly left to right, parentheses
 *     recurse, a ')' or the last character closes a unit; unit = <n><types><range>, ot_regex.py:245-262; range ->
 *     Python slice, ot_regex.py:63-122), independently of the product's parser;
 *   - an alignment is its column string ('.', 'm', 'i' = base only in the gRNA, 'd' = base only in the subject), grouped
 *     into one element per query position as blast.py:84-120 does: a 'd' is glued to the NEXT column for positive ranges
 *     (two 'd' in a row glue to each other: the 'dd' element of blast.py:114-116) or to the PREVIOUS element for negative
 *     ranges, ' ' for every unaligned position; a unit counts its characters in the Python slice [start:end:step] of
 *     that tuple (blast.py:138-172: None -> 0 / qlen, step -1 when end < start) and adds the unaligned positions of the
 *     slice for 'm' units (unaligned_as_mismatch) and 'i'/'g' units (unaligned_as_gap), ot_regex.py:265-277.
 * Enumeration: depth-first from the LAST column (a base pair) backwards; an alignment is complete whenever its leftmost
 * column is a base pair; at most gcap gap columns strictly inside.  Optional pruning (prune != 0, only when gcap <= 1 so
 * that no 'dd' element can shift the tuple): a partial alignment is dropped when the expression is false even if every
 * unit only counts what is already fixed (the known columns and the unaligned 3' tail) - counts only grow as columns are
 * added and the expression has no negation, so no satisfying alignment is lost.  tests/test_oracle_c.py checks the
 * pruned against the unpruned enumeration and both against the Python oracle and the golden vectors of the reference. */
#include <ctype.h>
#include <stdio.h>

#define PMAX_UNITS 64
#define PMAX_NODES 160
#define PMAX_COLS 80

typedef struct { int n, cm, ci, cd, has_s, s, has_e, e, rvs, add_m, add_g; } punit_t;
typedef struct { int kind; int unit; int l, r; } pnode_t;           /* kind 0 = unit, 1 = and, 2 = or */
typedef struct { punit_t units[PMAX_UNITS]; int n_units; pnode_t nodes[PMAX_NODES]; int n_nodes; int root; int uam, uag; } pexpr_t;

static int p_range(const char* t, int len, punit_t* u) {            /* ot_regex.py:63-122 */
    int i = 0;
    while (i < len && t[i] == '-') ++i;
    const int neg_start = i > 0;
    int d0 = i;
    while (i < len && isdigit((unsigned char)t[i])) ++i;
    if (i == d0) return -1;
    long start = strtol(t + d0, NULL, 10);
    if (d0 % 2 == 1) start = -start; else if (d0 > 0) start = start;    /* int('--5') is invalid in Python; see below */
    if (d0 > 1) return -1;                                           /* int("--5") raises ValueError -> MINORgError */
    if (start == 0) return -1;
    const int rest = len - i;
    if (rest == 0) {
        if (neg_start) { u->has_s = 1; u->s = (int)start; u->has_e = 0; }
        else { u->has_s = 0; u->has_e = 1; u->e = (int)start; }
        return 0;
    }
    if (rest == 1 && t[len - 1] == '-') {
        if (neg_start) { u->has_s = 0; if (start < -1) { u->has_e = 1; u->e = (int)start + 1; } else u->has_e = 0; }
        else { u->has_s = 1; u->s = (int)start - 1; u->has_e = 0; }
        return 0;
    }
    /* end = the digits at the very end with the dashes before them, preceded by one more '-' (lookbehind) */
    int k = len;
    while (k > 0 && isdigit((unsigned char)t[k - 1])) --k;
    if (k == len) return -1;
    int ds = k;
    while (ds > 0 && t[ds - 1] == '-') --ds;                         /* ds..k = all dashes before the digits */
    /* "(?<=-)-*\d+$": the match starts right after a '-', taking as many of the remaining dashes as possible */
    if (k - ds < 1) return -1;
    const int n_dash = k - ds - 1;                                   /* dashes that belong to the number */
    if (n_dash > 1) return -1;                                       /* int("--5") */
    long end = strtol(t + k, NULL, 10);
    if (n_dash == 1) end = -end;
    const int neg_end = n_dash == 1;
    if (neg_end != neg_start) return -1;
    if (end == 0) return -1;
    if (ds != i) return -1;                                          /* anything between the two numbers: not a range the regexes
                                                                        accept consistently; the product and the tests never use it */
    if (neg_end) {
        u->has_s = 1; u->s = (int)end;
        if (start < -1) { u->has_e = 1; u->e = (int)start + 1; } else u->has_e = 0;
    } else {
        u->has_s = 1; u->s = (int)start - 1; u->has_e = 1; u->e = (int)end;
    }
    return 0;
}

static int p_unit(pexpr_t* x, const char* t, int len) {              /* ot_regex.py:222-264 -> unit index or -1 */
    if (x->n_units >= PMAX_UNITS || len <= 0 || !isdigit((unsigned char)t[0])) return -1;
    punit_t u;
    memset(&u, 0, sizeof u);
    u.n = (int)strtol(t, NULL, 10);
    int mm = 0, gap = 0, ins = 0, del = 0;
    for (int i = 0; i < len; ++i) { mm |= t[i] == 'm'; gap |= t[i] == 'g'; ins |= t[i] == 'i'; del |= t[i] == 'd'; }
    if (gap && (ins || del)) return -1;                              /* self.logfile does not exist -> MINORgError (:251,263) */
    int k = len;
    while (k > 0 && (t[k - 1] == '-' || isdigit((unsigned char)t[k - 1]))) --k;    /* "[-\d]+$" */
    if (k == len) return -1;
    if (p_range(t + k, len - k, &u) != 0) return -1;
    u.rvs = (u.has_s && u.s < 0) || (u.has_e && u.e < 0);
    u.cm = mm; u.ci = ins || gap; u.cd = del || gap;
    u.add_m = mm && x->uam; u.add_g = (ins || gap) && x->uag;
    x->units[x->n_units] = u;
    return x->n_units++;
}

static int p_node(pexpr_t* x, int kind, int unit, int l, int r) {
    if (x->n_nodes >= PMAX_NODES) return -1;
    x->nodes[x->n_nodes].kind = kind; x->nodes[x->n_nodes].unit = unit; x->nodes[x->n_nodes].l = l; x->nodes[x->n_nodes].r = r;
    return x->n_nodes++;
}

/* ot_regex.py:295-356.  Returns the node of the expression parsed from p[0..n) (-1: error) and its length in *used. */
static int p_parse(pexpr_t* x, const char* p, int n, int* used) {
    int f = -1, op = 0, start = 0, i = 0;                            /* op: 0 none, 1 and, 2 or */
    const int last = n - 1;
#define P_BUILD(newnode) do { const int nn_ = (newnode); if (nn_ < 0) return -1; \
        if (f < 0) f = nn_; else if (op == 0) return -1; else { f = p_node(x, op, -1, f, nn_); if (f < 0) return -1; } } while (0)
    while (i < n) {
        const char c = p[i];
        if (c == '(') {
            int sub_len = 0;
            const int sub = p_parse(x, p + i + 1, n - i - 1, &sub_len);
            P_BUILD(sub);
            start = i + 1 + sub_len;
            i = start;
        } else if (c == ',' || c == '|') {
            if (start != i) { const int u = p_unit(x, p + start, i - start); if (u < 0) return -1; P_BUILD(p_node(x, 0, u, -1, -1)); }
            op = c == ',' ? 1 : 2;
            start = i + 1;
            ++i;
        } else if (c == ')') {
            const int u = p_unit(x, p + start, i - start); if (u < 0) return -1; P_BUILD(p_node(x, 0, u, -1, -1));
            ++i;
            break;
        } else if (i >= last) {
            const int u = p_unit(x, p + start, n - start); if (u < 0) return -1; P_BUILD(p_node(x, 0, u, -1, -1));
            ++i;
            break;
        } else ++i;
    }
#undef P_BUILD
    *used = i;
    return f;                                                         /* -1 when nothing was parsed (self.f None -> TypeError) */
}

typedef struct { uint8_t m, i, d, sp; } pelem_t;                     /* characters of one tuple element */

/* Python: len(range(*slice(start, end, step).indices(n))) members, visited in order; cb adds the element counts */
static void p_slice(int has_s, int s, int has_e, int e, int qlen, int n, int* first, int* stop, int* step) {
    long a = has_s ? s : 0, b = has_e ? e : qlen;                    /* blast.py:138-139 */
    *step = b >= a ? 1 : -1;                                         /* blast.py:141 */
    if (*step > 0) {
        if (a < 0) { a += n; if (a < 0) a = 0; } else if (a > n) a = n;
        if (b < 0) { b += n; if (b < 0) b = 0; } else if (b > n) b = n;
    } else {
        if (a < 0) { a += n; if (a < 0) a = -1; } else if (a >= n) a = n - 1;
        if (b < 0) { b += n; if (b < 0) b = -1; } else if (b >= n) b = n - 1;
    }
    *first = (int)a; *stop = (int)b;
}

static int p_unit_count(const punit_t* u, const pelem_t* tup, int n, int qlen) {
    int first, stop, step, cnt = 0;
    p_slice(u->has_s, u->s, u->has_e, u->e, qlen, n, &first, &stop, &step);
    for (int k = first; step > 0 ? k < stop : k > stop; k += step) {
        if (u->cm) cnt += tup[k].m;
        if (u->ci) cnt += tup[k].i;
        if (u->cd) cnt += tup[k].d;
        cnt += (u->add_m + u->add_g) * tup[k].sp;
    }
    return cnt;
}

static int p_eval(const pexpr_t* x, int node, const pelem_t* fwd, int nf, const pelem_t* rvs, int nr, int qlen) {
    const pnode_t* nd = &x->nodes[node];
    if (nd->kind == 0) {
        const punit_t* u = &x->units[nd->unit];
        return p_unit_count(u, u->rvs ? rvs : fwd, u->rvs ? nr : nf, qlen) <= u->n;
    }
    const int a = p_eval(x, nd->l, fwd, nf, rvs, nr, qlen);
    if (nd->kind == 1) return a && p_eval(x, nd->r, fwd, nf, rvs, nr, qlen);
    return a || p_eval(x, nd->r, fwd, nf, rvs, nr, qlen);
}

/* kinds[0..nc): the alignment's columns in forward order; qs / qe / L as Biopython reports them.  Builds both tuples
 * (blast.py:84-120) and evaluates the expression. */
static int p_tuples(const char* kinds, int nc, int qs, int qe, int L, pelem_t* fwd, int* nf, pelem_t* rvs, int* nr) {
    int a = 0, b = 0;
    pelem_t sp = {0, 0, 0, 1}, z = {0, 0, 0, 0};
    for (int k = 0; k < qs; ++k) { fwd[a++] = sp; rvs[b++] = sp; }
    for (int k = 0; k < nc;) {                                        /* rvs_index = False */
        pelem_t e = z;
        if (kinds[k] != 'd') { e.m = kinds[k] == 'm'; e.i = kinds[k] == 'i'; ++k; }
        else {                                                        /* 'd' + the next character, whatever it is */
            e.d = 1;
            if (k + 1 < nc) { const char c2 = kinds[k + 1]; e.m = c2 == 'm'; e.i = c2 == 'i'; e.d += c2 == 'd'; }
            k += 2;
        }
        fwd[a++] = e;
    }
    for (int k = 0; k < nc; ++k) {                                    /* rvs_index = True */
        if (kinds[k] != 'd') { pelem_t e = z; e.m = kinds[k] == 'm'; e.i = kinds[k] == 'i'; rvs[b++] = e; }
        else if (b > 0) rvs[b - 1].d++;
    }
    for (int k = qe; k < L; ++k) { fwd[a++] = sp; rvs[b++] = sp; }
    *nf = a; *nr = b;
    return 0;
}

typedef struct {
    const pexpr_t* x;
    gctx_t g;                          /* text, masks, PAM tables (qs is filled per alignment) */
    int gcap, prune;
    int qe; int64_t t_end;             /* last column: query position qe-1 against subject t_end-1 */
    char cols[PMAX_COLS];              /* columns so far, in REVERSE order (cols[0] = last column) */
} pctx_t;

static int p_complete_true(pctx_t* c, int nc, int qs) {
    char kinds[PMAX_COLS];
    for (int k = 0; k < nc; ++k) kinds[k] = c->cols[nc - 1 - k];
    pelem_t fwd[PMAX_COLS], rvs[PMAX_COLS];
    int nf, nr;
    p_tuples(kinds, nc, qs, c->qe, c->g.L, fwd, &nf, rvs, &nr);
    return p_eval(c->x, c->x->root, fwd, nf, rvs, nr, c->g.L);
}

/* Lower bound for every completion of the partial alignment (known: query positions >= j and the 3' tail): the columns
 * known so far as an alignment that starts at query position j, nothing counted before it.  Only used with gcap <= 1. */
static int p_may_succeed(pctx_t* c, int nc, int j) {
    char kinds[PMAX_COLS];
    for (int k = 0; k < nc; ++k) kinds[k] = c->cols[nc - 1 - k];
    pelem_t fwd[PMAX_COLS], rvs[PMAX_COLS], z = {0, 0, 0, 0};
    int a = 0, b = 0, nf, nr;
    const int L = c->g.L;
    /* positions [0, j) unknown: optimistic zero elements, so that the tuple keeps one element per query position */
    pelem_t tf[PMAX_COLS], tr[PMAX_COLS];
    p_tuples(kinds, nc, 0, c->qe, L, tf, &nf, tr, &nr);      /* elements of positions j.. (a leading 'd' glues forward) */
    int lead_d = nc > 0 && kinds[0] == 'd';
    for (int k = 0; k < j; ++k) { fwd[a++] = z; rvs[b++] = z; }
    for (int k = 0; k < nf; ++k) fwd[a++] = tf[k];
    if (lead_d && b > 0) rvs[b - 1].d++;                     /* for negative ranges a leading 'd' belongs to position j-1 */
    for (int k = 0; k < nr; ++k) rvs[b++] = tr[k];
    if (a != L || b != L) return 1;                          /* unexpected shape: do not prune */
    return p_eval(c->x, c->x->root, fwd, a, rvs, b, L);
}

static void p_dfs(pctx_t* c, int j, int64_t t, int gaps, int mm, int nc, int last_pair) {
    /* columns cover query [j, qe) and subject [t, t_end); the leftmost column is a pair iff last_pair */
    if (nc + 2 >= PMAX_COLS) return;
    /* extend with a base pair (j-1, t-1) */
    if (j >= 1 && t >= 1) {
        const int qb = c->g.q[j - 1], sb = c->g.T[t - 1];
        const int same = qb == sb && qb < 4;
        c->cols[nc] = same ? '.' : 'm';
        const int ok = !c->prune || p_may_succeed(c, nc + 1, j - 1);
        if (ok) {
            const int qs = j - 1;
            const int64_t e = c->t_end + (c->g.L - c->qe);
            if (!((c->g.ends[e >> 3] >> (e & 7)) & 1) && p_complete_true(c, nc + 1, qs) && !g_masked(&c->g, t - 1, c->t_end)) {
                c->g.qs = qs;
                if (c->g.pamless || g_has_pam(&c->g, c->qe, gaps, t - 1, c->t_end)) c->g.ends[e >> 3] |= (uint8_t)(1u << (e & 7));
            }
            p_dfs(c, j - 1, t - 1, gaps, mm + !same, nc + 1, 1);
        }
    }
    if (nc > 0 && gaps < c->gcap) {
        if (j >= 2) {                 /* insertion: query position j-1 alone; a pair must still follow on the left */
            c->cols[nc] = 'i';
            if (!c->prune || p_may_succeed(c, nc + 1, j - 1)) p_dfs(c, j - 1, t, gaps + 1, mm, nc + 1, 0);
        }
        if (j >= 1 && t >= 2) {       /* deletion: subject base t-1 alone */
            c->cols[nc] = 'd';
            if (!c->prune || p_may_succeed(c, nc + 1, j)) p_dfs(c, j, t - 1, gaps + 1, mm, nc + 1, 0);
        }
    }
    (void)last_pair;
}

typedef struct {
    const pexpr_t* x; const uint8_t* fwd; const uint8_t* rev; int64_t n;
    const uint8_t* qcodes; int nq, L, gcap, prune;
    const int64_t* ms; const int64_t* me; int n_masks;
    int q0, q1; uint32_t* counts;
    int pamless, n5, n3, max5, max3;
    const int32_t* len5; const uint8_t* allow5; const int32_t* len3; const uint8_t* allow3;
    int64_t wlo, whi;              /* forward window starts whose anchors count (see anchor_range) */
} pjob_t;

static void* p_run(void* arg) {
    pjob_t* j = (pjob_t*)arg;
    const size_t nbytes = (size_t)(j->n + j->L + 16) / 8 + 2;
    uint8_t* ends = (uint8_t*)malloc(nbytes);
    for (int qi = j->q0; qi < j->q1; ++qi) {
        for (int strand = 1; strand >= -1; strand -= 2) {
            memset(ends, 0, nbytes);
            pctx_t c;
            memset(&c, 0, sizeof c);
            c.x = j->x; c.gcap = j->gcap; c.prune = j->prune && j->gcap <= 1;
            c.g.T = strand > 0 ? j->fwd : j->rev; c.g.n = j->n; c.g.q = j->qcodes + (size_t)qi * j->L;
            c.g.L = j->L; c.g.M = 0; c.g.Gp = j->gcap; c.g.B = 0;
            c.g.ms = j->ms; c.g.me = j->me; c.g.n_masks = j->n_masks; c.g.strand = strand; c.g.ends = ends;
            c.g.pamless = j->pamless; c.g.n5 = j->n5; c.g.n3 = j->n3; c.g.max5 = j->max5; c.g.max3 = j->max3;
            c.g.len5 = j->len5; c.g.allow5 = j->allow5; c.g.len3 = j->len3; c.g.allow3 = j->allow3;
            int64_t e0, e1;
            anchor_range(strand, j->n, j->L, j->wlo, j->whi, &e0, &e1);
            if (e1 <= e0) continue;
            for (int qe = j->L; qe >= 1; --qe) {
                c.qe = qe;
                if (c.prune && !p_may_succeed(&c, 0, qe)) continue;     /* the 3' tail alone already decides */
                /* anchor e = t + (L - qe) must lie in [e0, e1) */
                int64_t t0 = e0 - (j->L - qe), t1 = e1 - (j->L - qe);
                if (t0 < 1) t0 = 1;
                if (t1 > j->n + 1) t1 = j->n + 1;
                for (int64_t t = t0; t < t1; ++t) { c.t_end = t; p_dfs(&c, qe, t, 0, 0, 0, 0); }
            }
            uint32_t cnt = 0;
            for (int64_t e = e0; e < e1; ++e) cnt += (ends[e >> 3] >> (e & 7)) & 1u;
            j->counts[strand > 0 ? qi : j->nq + qi] += cnt;
        }
    }
    free(ends);
    return NULL;
}

/* Parses `pattern` with the reference's rules; returns 0 and the number of units, or -1 if the reference would raise. */
int oracle_pattern_parse(const char* pattern, int uam, int uag, int* n_units_out) {
    pexpr_t x;
    memset(&x, 0, sizeof x);
    x.uam = uam; x.uag = uag;
    int used = 0;
    const int root = p_parse(&x, pattern, (int)strlen(pattern), &used);
    if (n_units_out) *n_units_out = x.n_units;
    return root < 0 ? -1 : 0;
}

/* One alignment given as the reference sees it (column string + query_start / query_end / qlen): expression value (0 / 1),
 * -1 on a parse error.  Used to pin this file on the golden pattern evaluations produced by the reference's own classes. */
int oracle_pattern_eval(const char* pattern, int uam, int uag, const char* kinds, int qs, int qe, int qlen) {
    pexpr_t x;
    memset(&x, 0, sizeof x);
    x.uam = uam; x.uag = uag;
    int used = 0;
    x.root = p_parse(&x, pattern, (int)strlen(pattern), &used);
    if (x.root < 0) return -1;
    const int nc = (int)strlen(kinds);
    if (nc + qlen >= PMAX_COLS) return -1;
    pelem_t fwd[PMAX_COLS * 2], rvs[PMAX_COLS * 2];
    int nf, nr;
    p_tuples(kinds, nc, qs, qe, qlen, fwd, &nf, rvs, &nr);
    return p_eval(&x, x.root, fwd, nf, rvs, nr, qlen);
}

/* counts_out[2n]: [i] = '+' strand anchors of gRNA i, [n+i] = '-' strand anchors.  gcap = max(1, ot_gap) - 1. */
int oracle_screen_pattern(const char* seq, const int64_t* offsets, int n_contigs,
                          const int32_t* mask_contig, const int64_t* mask_start, const int64_t* mask_end, int n_masks,
                          const char* grna, int n, int L, const char* pattern, int uam, int uag, int gcap, int prune,
                          int n_threads, uint32_t* counts_out,
                          int pamless, int n5, const int32_t* len5, const uint8_t* allow5, int max5,
                          int n3, const int32_t* len3, const uint8_t* allow3, int max3,
                          const int64_t* win_lo, const int64_t* win_hi /* per contig, NULL = everything */) {
    if (L < 1 || L > 32 || n < 0 || n_threads < 1 || gcap < 0 || gcap > 3) return -2;
    pexpr_t x;
    memset(&x, 0, sizeof x);
    x.uam = uam; x.uag = uag;
    int used = 0;
    x.root = p_parse(&x, pattern, (int)strlen(pattern), &used);
    if (x.root < 0) return -3;
    uint8_t* qcodes = (uint8_t*)malloc((size_t)n * L + 1);
    for (size_t i = 0; i < (size_t)n * L; ++i) qcodes[i] = (uint8_t)code_of(grna[i]);
    memset(counts_out, 0, sizeof(uint32_t) * 2 * (size_t)n);
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)n_threads);
    pjob_t* jobs = (pjob_t*)malloc(sizeof(pjob_t) * (size_t)n_threads);
    for (int c = 0; c < n_contigs; ++c) {
        const int64_t len = offsets[c + 1] - offsets[c];
        if (len == 0) continue;
        uint8_t* fwd = (uint8_t*)malloc((size_t)len), *rev = (uint8_t*)malloc((size_t)len);
        for (int64_t i = 0; i < len; ++i) {
            const int k = code_of(seq[offsets[c] + i]);
            fwd[i] = (uint8_t)k;
            rev[len - 1 - i] = (uint8_t)(k == 4 ? 4 : 3 - k);
        }
        int nm = 0;
        int64_t* ms = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n_masks + 1)), *me = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n_masks + 1));
        for (int i = 0; i < n_masks; ++i) if (mask_contig[i] == c) { ms[nm] = mask_start[i]; me[nm] = mask_end[i]; ++nm; }
        for (int t = 0; t < n_threads; ++t) {
            pjob_t* j = &jobs[t];
            j->x = &x; j->fwd = fwd; j->rev = rev; j->n = len; j->qcodes = qcodes; j->nq = n; j->L = L; j->gcap = gcap; j->prune = prune;
            j->ms = ms; j->me = me; j->n_masks = nm; j->q0 = (int)((int64_t)n * t / n_threads); j->q1 = (int)((int64_t)n * (t + 1) / n_threads);
            j->counts = counts_out;
            j->pamless = pamless; j->n5 = n5; j->n3 = n3; j->max5 = max5; j->max3 = max3;
            j->len5 = len5; j->allow5 = allow5; j->len3 = len3; j->allow3 = allow3;
            j->wlo = win_lo ? win_lo[c] : -(int64_t)L - 8; j->whi = win_hi ? win_hi[c] : len + 8;
            pthread_create(&th[t], NULL, p_run, j);
        }
        for (int t = 0; t < n_threads; ++t) pthread_join(th[t], NULL);
        free(fwd); free(rev); free(ms); free(me);
    }
    free(jobs); free(th); free(qcodes);
    return 0;
}

/* ------------------------------------------------------------------------------------------------------------
 * Bit-sliced CPU port of the ungapped screen - the CPU BASELINE bench.py times beside the GPU ("port"), so that the
 * comparison is not against a strawman: the same counts as oracle_screen_hamming_win (tests/test_oracle_c.py compares
 * them), computed the way the device's pair-word scan does it, with 64 oriented queries per machine word:
 *   stage 1  per window and group of 64 queries: one table word per PAIR of gRNA positions ("query i mismatches at least
 *            one base of the pair", indexed by the 4-bit pair of genome bases), bit-sliced count of the set words
 *            (carry-save adders), compare with the budget - the number of set pair words never exceeds the mismatches;
 *   stage 2  survivors only: exact 2-bit XOR/popcount of the window against the query.
 * ~35 word operations per 64 cells instead of ~6 per cell.  Plain C, no intrinsics beyond popcount; pthreads over windows. */
typedef struct {
    const uint8_t* codes; int64_t n; int64_t p0, p1;
    const uint64_t* tab2;              /* [n_groups][16*NP + 1] */
    const uint64_t* qbits; const uint64_t* qinv;
    int nq, n_groups, L, NP, max_mm;
    const int64_t* ms; const int64_t* me; int n_masks;
    uint32_t* counts;
} bjob_t;

static inline void bs_add(uint64_t cnt[5], uint64_t x) {          /* bit-sliced counter += x (ripple of half adders) */
    for (int b = 0; b < 5 && x; ++b) { const uint64_t t = cnt[b] & x; cnt[b] ^= x; x = t; }
}

static inline uint64_t bs_le(const uint64_t cnt[5], int K) {        /* bit i: count_i <= K */
    uint64_t lt = 0, eq = ~0ULL;
    for (int b = 4; b >= 0; --b) {
        if ((K >> b) & 1) { lt |= eq & ~cnt[b]; eq &= cnt[b]; } else eq &= ~cnt[b];
    }
    return lt | eq;
}

static void* brun(void* arg) {
    bjob_t* j = (bjob_t*)arg;
    const int L = j->L, NP = j->NP, S2 = 16 * NP + 1;
    const uint64_t even = 0x5555555555555555ULL;
    const uint64_t lmask = (L >= 32) ? even : (((1ULL << (2 * L)) - 1) & even);
    job_t mj;                                                        /* for span_masked() */
    mj.ms = j->ms; mj.me = j->me; mj.n_masks = j->n_masks;
    int off[16];
    for (int64_t p = j->p0; p < j->p1; ++p) {
        uint64_t w = 0, winv = 0;
        for (int k = 0; k < L; ++k) {
            const int64_t i = p + k;
            const int c = (i < 0 || i >= j->n) ? 4 : j->codes[i];
            if (c == 4) winv |= 1ULL << (2 * k); else w |= (uint64_t)c << (2 * k);
        }
        for (int k = 0; k < NP; ++k) {
            const int bad = (int)((winv >> (4 * k)) & 5ULL & ((2 * k + 1 < L) ? 5ULL : 1ULL));
            off[k] = bad ? 16 * NP : 16 * k + (int)((w >> (4 * k)) & 15ULL);
        }
        int masked = -1;
        for (int g = 0; g < j->n_groups; ++g) {
            const uint64_t* t = j->tab2 + (size_t)g * S2;
            uint64_t alive;
            if (NP == 10) {                                          /* carry-save tree for the common 20-nt case */
#define FA(a, b, c, s, cy) do { const uint64_t a_ = (a), b_ = (b), c_ = (c); s = a_ ^ b_ ^ c_; cy = (a_ & b_) | (c_ & (a_ ^ b_)); } while (0)
                uint64_t s0, c0, s1, c1, s2, c2, t0, c3, u0, d0, b0, b1, c4, d1;
                FA(t[off[0]], t[off[1]], t[off[2]], s0, c0);
                FA(t[off[3]], t[off[4]], t[off[5]], s1, c1);
                FA(t[off[6]], t[off[7]], t[off[8]], s2, c2);
                FA(s0, s1, s2, t0, c3);
                b0 = t0 ^ t[off[9]]; c4 = t0 & t[off[9]];
                FA(c0, c1, c2, u0, d0);
                FA(u0, c3, c4, b1, d1);
                uint64_t cnt[5] = {b0, b1, d0 ^ d1, d0 & d1, 0};
#undef FA
                alive = bs_le(cnt, j->max_mm);
            } else {
                uint64_t cnt[5] = {0, 0, 0, 0, 0};
                for (int k = 0; k < NP; ++k) bs_add(cnt, t[off[k]]);
                alive = bs_le(cnt, j->max_mm);
            }
            while (alive) {
                const int i = __builtin_ctzll(alive);
                alive &= alive - 1;
                const int q = g * 64 + i;
                if (q >= j->nq) break;
                const uint64_t x = w ^ j->qbits[q];
                const uint64_t d = ((x | (x >> 1)) & lmask) | winv | j->qinv[q];
                if (__builtin_popcountll(d) > j->max_mm) continue;
                if (masked < 0) {
                    const int64_t a = p < 0 ? 0 : p, b = (p + L > j->n) ? j->n : p + L;
                    masked = (a >= b) ? 1 : span_masked(&mj, a, b);
                }
                if (!masked) j->counts[q]++;
            }
        }
    }
    return NULL;
}

int oracle_screen_hamming_bitsliced(const char* seq, const int64_t* offsets, int n_contigs,
                                    const int32_t* mask_contig, const int64_t* mask_start, const int64_t* mask_end, int n_masks,
                                    const char* grna, int n, int L, int max_mm, int n_threads, uint32_t* counts_out,
                                    const int64_t* win_lo, const int64_t* win_hi) {
    if (L < 1 || L > 32 || n < 0 || n_threads < 1 || max_mm < 0 || max_mm > 31) return -2;
    const int nq = 2 * n, NP = (L + 1) / 2, S2 = 16 * NP + 1, n_groups = (nq + 63) / 64;
    uint64_t* qbits = (uint64_t*)calloc((size_t)nq + 1, 8);
    uint64_t* qinv = (uint64_t*)calloc((size_t)nq + 1, 8);
    uint8_t* qc = (uint8_t*)malloc((size_t)nq * L + 1);              /* oriented query codes, 4 = not ACGT */
    for (int i = 0; i < n; ++i)
        for (int k = 0; k < L; ++k) {
            const int c = code_of(grna[(size_t)i * L + k]);
            qc[(size_t)i * L + k] = (uint8_t)c;
            qc[(size_t)(n + i) * L + (L - 1 - k)] = (uint8_t)(c == 4 ? 4 : 3 - c);
            if (c == 4) { qinv[i] |= 1ULL << (2 * k); qinv[n + i] |= 1ULL << (2 * (L - 1 - k)); }
            else { qbits[i] |= (uint64_t)c << (2 * k); qbits[n + i] |= (uint64_t)(3 - c) << (2 * (L - 1 - k)); }
        }
    uint64_t* tab2 = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)n_groups * S2);
    for (int g = 0; g < n_groups; ++g)
        for (int k = 0; k < NP; ++k)
            for (int nib = 0; nib < 16; ++nib) {
                uint64_t wd = 0;
                for (int i = 0; i < 64; ++i) {
                    const int q = g * 64 + i;
                    int mism = 1;
                    if (q < nq) {
                        mism = qc[(size_t)q * L + 2 * k] != (nib & 3);
                        if (2 * k + 1 < L) mism |= qc[(size_t)q * L + 2 * k + 1] != ((nib >> 2) & 3);
                    }
                    wd |= (uint64_t)mism << i;
                }
                tab2[(size_t)g * S2 + 16 * k + nib] = wd;
            }
    for (int g = 0; g < n_groups; ++g) tab2[(size_t)g * S2 + 16 * NP] = ~0ULL;
    memset(counts_out, 0, sizeof(uint32_t) * (size_t)n);
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)n_threads);
    bjob_t* jobs = (bjob_t*)malloc(sizeof(bjob_t) * (size_t)n_threads);
    uint32_t* priv = (uint32_t*)calloc((size_t)n_threads * (size_t)(nq + 1), 4);
    for (int c = 0; c < n_contigs; ++c) {
        const int64_t len = offsets[c + 1] - offsets[c];
        int64_t first = -(int64_t)L + 1, last = len;
        if (win_lo != NULL && win_hi != NULL) {
            if (win_lo[c] > first) first = win_lo[c];
            if (win_hi[c] < last) last = win_hi[c];
        }
        if (last <= first) continue;
        uint8_t* codes = (uint8_t*)malloc((size_t)len + 1);
        for (int64_t i = 0; i < len; ++i) codes[i] = (uint8_t)code_of(seq[offsets[c] + i]);
        int nm = 0;
        int64_t* ms = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n_masks + 1));
        int64_t* me = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n_masks + 1));
        for (int i = 0; i < n_masks; ++i) if (mask_contig[i] == c) { ms[nm] = mask_start[i]; me[nm] = mask_end[i]; ++nm; }
        const int64_t total = last - first;
        for (int t = 0; t < n_threads; ++t) {
            bjob_t* j = &jobs[t];
            j->codes = codes; j->n = len; j->tab2 = tab2; j->qbits = qbits; j->qinv = qinv; j->nq = nq; j->n_groups = n_groups;
            j->L = L; j->NP = NP; j->max_mm = max_mm; j->ms = ms; j->me = me; j->n_masks = nm;
            j->counts = priv + (size_t)t * (size_t)(nq + 1);
            j->p0 = first + total * t / n_threads; j->p1 = first + total * (t + 1) / n_threads;
            pthread_create(&th[t], NULL, brun, j);
        }
        for (int t = 0; t < n_threads; ++t) pthread_join(th[t], NULL);
        free(codes); free(ms); free(me);
    }
    for (int t = 0; t < n_threads; ++t)
        for (int i = 0; i < n; ++i)
            counts_out[i] += priv[(size_t)t * (size_t)(nq + 1) + i] + priv[(size_t)t * (size_t)(nq + 1) + n + i];
    free(priv); free(jobs); free(th); free(qbits); free(qinv); free(qc); free(tab2);
    return 0;
}
