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
} gjob_t;

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
            for (int64_t s = 0; s < j->n; ++s)
                for (int qs = 0; qs < j->L && qs <= c.B; ++qs) { c.qs = qs; c.s = s; g_dfs(&c, qs, s, 0, 0, 0); }
            uint32_t cnt = 0;
            for (size_t b = 0; b < nbytes; ++b) cnt += (uint32_t)__builtin_popcount(ends[b]);
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
int oracle_screen_nopattern_pam(const char* seq, const int64_t* offsets, int n_contigs,
                                const int32_t* mask_contig, const int64_t* mask_start, const int64_t* mask_end, int n_masks,
                                const char* grna, int n, int L, int max_mm, int max_gap, int n_threads, uint32_t* counts_out,
                                int pamless, int n5, const int32_t* len5, const uint8_t* allow5, int max5,
                                int n3, const int32_t* len3, const uint8_t* allow3, int max3) {
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
            pthread_create(&th[t], NULL, g_run, j);
        }
        for (int t = 0; t < n_threads; ++t) pthread_join(th[t], NULL);
        free(fwd); free(rev); free(ms); free(me);
    }
    free(jobs); free(th); free(qcodes);
    return 0;
}
