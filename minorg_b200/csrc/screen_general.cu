// K4a/K5: the general screen - off-target patterns, gaps, PAM proximity - as prefilter + exact verifier.
//
// Replaces MINORg._offtarget_hits for every criterion the reference offers (MINORg.py:2015-2102):
//   prefilter  (a NECESSARY condition, evaluated for every cell)
//     * no gaps allowed : the bit-sliced Hamming scan of screen_bitsliced.cu with budget K (cold path = verifier)
//     * gaps allowed    : Myers' bit-parallel <=K-edit search (one 32-bit column per text character)
//     * no finite K     : every cell goes to the verifier (degenerate patterns; small inputs)
//   verifier   verify.cuh: exact enumeration of the alignment universe at the surviving anchors.
// K = max_mismatch + max_gap without a pattern (MINORg.py:2037-2038: unaligned positions and gaps spend one
// combined budget), or the bound the host derives from the pattern (ot_regex.OffTargetExpression.edit_bound).
#include <algorithm>
#include <chrono>
#include "common.cuh"
#include "verify.cuh"

namespace mg {

int launch_screen_bitsliced_general(const mg_genome* g, const mg_queries* q, int k, const GeneralCtx* d_ctx,
                                    int64_t wb, int64_t we, uint32_t* d_counts, cudaStream_t s);
int launch_gapped_rows(const mg_genome* g, const mg_queries* q, const GeneralCtx* d_ctx, int K, int strand, int g_begin,
                       int g_end, int64_t eb, int64_t ee, uint32_t** d_tab_cache, cudaStream_t s);
bool seedjoin_eligible(const mg_queries* q, const mg_criteria* crit);
int launch_seedjoin(const mg_genome* g, const mg_queries* q, const mg_criteria* crit, const GeneralCtx* d_ctx, int64_t wb, int64_t we,
                    uint32_t* d_counts, cudaStream_t s, std::vector<std::pair<int64_t, int64_t>>& dirty_ranges);

// ----------------------------------------------------------------------------------------- Myers <= K
constexpr int kMyersThreads = 128;     // gRNA per CTA (one per thread; a warp shares the text character)
constexpr int kMyersChunk = 8192;      // text characters per CTA (plus L+K warm-up)

__global__ void __launch_bounds__(kMyersThreads)
myers_kernel(GenomeView gv, QueriesView qv, const GeneralCtx* __restrict__ ctx, int K, int strand, int64_t eb,
             int64_t ee, uint32_t* __restrict__ counts) {
    // strand coordinates: anchor e (virtual end, exclusive) in [eb, ee); text position pos = e - 1
    __shared__ uint8_t text[kMyersChunk + 2 * MG_MAX_GRNA_LEN + 16];
    __shared__ uint32_t peq[5][kMyersThreads];
    const int L = qv.len, N = qv.n;
    const int gi = blockIdx.y * kMyersThreads + threadIdx.x;
    const int64_t p0 = eb - 1 + (int64_t)blockIdx.x * kMyersChunk;      // first reported text position
    const int64_t p1 = min(ee - 1, p0 + kMyersChunk);                   // one past the last
    if (p0 >= p1) return;
    const int warm = L + K;
    const int64_t from = p0 - warm;
    const int n_text = (int)(p1 - from);
    for (int i = threadIdx.x; i < n_text; i += kMyersThreads) text[i] = (uint8_t)tbase(gv, strand, from + i);
    {
        uint32_t e0 = 0, e1 = 0, e2 = 0, e3 = 0;
        if (gi < N) {
            const uint8_t* q = qv.codes + (size_t)gi * L;     // gRNA orientation (first N rows)
            for (int j = 0; j < L; ++j) {
                const uint32_t b = 1u << j;
                const int c = q[j];
                e0 |= c == 0 ? b : 0u; e1 |= c == 1 ? b : 0u; e2 |= c == 2 ? b : 0u; e3 |= c == 3 ? b : 0u;
            }
        }
        peq[0][threadIdx.x] = e0; peq[1][threadIdx.x] = e1; peq[2][threadIdx.x] = e2; peq[3][threadIdx.x] = e3;
        peq[4][threadIdx.x] = 0u;                            // invalid base: matches nothing
    }
    __syncthreads();
    if (gi >= N) return;
    const uint32_t top = 1u << (L - 1);
    uint32_t Pv = L >= 32 ? 0xFFFFFFFFu : ((1u << L) - 1u), Mv = 0u;
    int score = L;
    for (int i = 0; i < n_text; ++i) {
        const uint32_t Eq = peq[text[i]][threadIdx.x];
        const uint32_t Xv = Eq | Mv;
        const uint32_t Xh = (((Eq & Pv) + Pv) ^ Pv) | Eq;
        uint32_t Ph = Mv | ~(Xh | Pv);
        uint32_t Mh = Pv & Xh;
        score += (Ph & top) ? 1 : 0;
        score -= (Mh & top) ? 1 : 0;
        Ph <<= 1;
        Mh <<= 1;
        Pv = Mh | ~(Xv | Ph);
        Mv = Ph & Xv;
        if (score <= K && i >= warm) emit_candidate(*ctx, strand > 0 ? gi : N + gi, from + i + 1);
    }
}

// ----------------------------------------------------------------------------------------- exact pieces
// Pigeonhole prefilter for patterns (host proof: ot_regex.OffTargetExpression.exact_pieces): a query survives at
// window p iff it matches the subject exactly on one of the given position ranges.  Same bit-sliced tables as
// the scan kernel (OR of the mismatch words of a range == 0): pair words where two positions of the range share
// one, single-position words at an odd edge.  One launch per (piece, orientation), instantiated for its number of
// lookups T; after the first three words a thread drops the group unless some query still matches (a 6-base exact
// match: 1 word in 128), so the kernel runs at ~3.4 shared loads per (window, 32 gRNA).  Four groups per trip and 32
// warps per SM keep enough loads in flight (one group per trip with 16 warps was latency-bound at half the rate;
// looking at four words before the drop instead of three measures the same).  A gap after (before) the
// matching range moves the alignment's virtual end by at most Gp, so 2*Gp+1 anchors go to the exact verifier.
constexpr int kPieceThreads = 1024;
constexpr int kPieceLookups = 9;       // 16 positions: at most 7 pair words + 2 single words

struct PieceOp {
    int piece;                         // index of the piece in the criteria
    int minus;                         // 1: the reverse-complement queries (mirrored range)
    int n;                             // lookups
    int pair[kPieceLookups];           // 1: idx = pair index (positions 2 idx, 2 idx + 1); 0: idx = position
    int idx[kPieceLookups];
};

// cold path: the queries of group word `hit` (already restricted to the op's orientation) survive at window p
// An anchor is only queued if the window that owns it (e - L on '+', glen - e on '-') lies in [own_a, own_b): the scan
// itself covers Gp extra windows on both sides, so every anchor of the shard is reached and none of a neighbour's is.
// The tag (piece, offset) lets the verifier count an anchor that several (piece, window) pairs reach exactly once.
static __device__ __noinline__ void emit_piece_hits(const GeneralCtx* __restrict__ ctx, uint32_t hit, int q0, int N, int L,
                                                    int64_t p, int64_t glen, int piece, int64_t own_a, int64_t own_b) {
    const int Gp = ctx->Gp;
    while (hit) {
        const int i = __ffs(hit) - 1;
        hit &= hit - 1;
        const int q = q0 + i;
        if (q >= 2 * N) continue;
        const int64_t e0 = q < N ? p + L : glen - p;
        for (int d = -Gp; d <= Gp; ++d) {
            const int64_t pw = q < N ? p + d : p - d;            // window owning anchor e0 + d
            if (pw >= own_a && pw < own_b) emit_candidate(*ctx, q, e0 + d, kCandTagged | (piece << 8) | (d + Gp));
        }
    }
}

template <int T, int EE>
__global__ void __launch_bounds__(kPieceThreads, 1)
pieces_kernel(GenomeView gv, QueriesView qv, const GeneralCtx* __restrict__ ctx, PieceOp op, int groups_per_chunk, int64_t wb,
              int64_t we, int64_t own_a, int64_t own_b) {
    extern __shared__ uint32_t tab[];                      // per group: pair table, then single-position table
    const int L = qv.len, N = qv.n, NPL = (L + 1) / 2;
    const int S2 = 16 * NPL + 1, S1 = 4 * L + 1, SG = S2 + S1;
    constexpr int E = T < EE ? T : EE;                     // words looked at before the early drop
    // groups holding queries of this orientation: '+' queries are 0..N-1, '-' queries N..2N-1 (group N/32 may hold both)
    const int g_lo = op.minus ? N / 32 : 0, g_hi = op.minus ? qv.n_groups : (N + 31) / 32;
    for (int cg = g_lo; cg < g_hi; cg += groups_per_chunk) {
        const int ng = min(groups_per_chunk, g_hi - cg);
        __syncthreads();
        for (int i = threadIdx.x; i < ng * SG; i += blockDim.x) {
            const int g = i / SG, r = i - g * SG;
            tab[i] = r < S2 ? __ldg(qv.tab2 + (size_t)(cg + g) * S2 + r) : __ldg(qv.tab + (size_t)(cg + g) * S1 + (r - S2));
        }
        __syncthreads();
        for (int64_t p = wb + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < we; p += (int64_t)gridDim.x * blockDim.x) {
            uint32_t lo, hi, inv;
            load_window(gv, p, lo, hi, inv);
            const unsigned long long w = ((unsigned long long)hi << 32) | lo;
            uint32_t off[T];                               // byte offsets of this window's table words inside a group
#pragma unroll
            for (int x = 0; x < T; ++x) {
                const int k = op.idx[x];
                uint32_t word;
                if (op.pair[x]) {
                    const uint32_t nib = (uint32_t)(w >> (4 * k)) & 15u;
                    word = ((inv >> (2 * k)) & 3u) ? (uint32_t)(16 * NPL) : (uint32_t)(16 * k) + nib;
                } else {
                    const uint32_t code = (uint32_t)(w >> (2 * k)) & 3u;
                    word = (uint32_t)S2 + (((inv >> k) & 1u) ? (uint32_t)(4 * L) : (uint32_t)(4 * k) + code);
                }
                off[x] = word * 4u;
            }
            // the rest of the words of one group and, for the (rare) survivors, the hand-over to the verifier
            auto finish = [&](int g, const char* t, uint32_t acc) {
#pragma unroll
                for (int x = E; x < T; ++x) acc |= *reinterpret_cast<const uint32_t*>(t + off[x]);
                uint32_t hit = ~acc;
                if (hit == 0u) return;
                const int q0 = (cg + g) * 32;
                uint32_t plus_bits = 0u;
                if (q0 + 32 <= N) plus_bits = 0xFFFFFFFFu;
                else if (q0 < N) plus_bits = (1u << (N - q0)) - 1u;
                hit &= op.minus ? ~plus_bits : plus_bits;
                if (hit) emit_piece_hits(ctx, hit, q0, N, L, p, gv.glen, op.piece, own_a, own_b);
            };
            const char* t = reinterpret_cast<const char*>(tab);
            int g = 0;
            // U groups per trip so that their loads are in flight together: one group per trip is latency-bound
            // (load -> OR -> compare -> branch with 4 warps per scheduler: issue slots 48 % busy, ncu)
            constexpr int U = 4;
            for (; g + U <= ng; g += U, t += U * SG * 4) {
                uint32_t acc[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    acc[u] = 0u;
#pragma unroll
                    for (int x = 0; x < E; ++x) acc[u] |= *reinterpret_cast<const uint32_t*>(t + u * (SG * 4) + off[x]);
                }
                uint32_t all = acc[0];
#pragma unroll
                for (int u = 1; u < U; ++u) all &= acc[u];
                if (all == 0xFFFFFFFFu) continue;
#pragma unroll
                for (int u = 0; u < U; ++u)
                    if (acc[u] != 0xFFFFFFFFu) finish(g + u, t + u * (SG * 4), acc[u]);
            }
            for (; g < ng; ++g, t += SG * 4) {
                uint32_t acc = 0u;
#pragma unroll
                for (int x = 0; x < E; ++x) acc |= *reinterpret_cast<const uint32_t*>(t + off[x]);
                if (acc != 0xFFFFFFFFu) finish(g, t, acc);
            }
        }
    }
}

// the lookups that cover query positions [a, a + n) (of the orientation's own coordinates)
static PieceOp make_piece_op(int a, int n, int minus) {
    PieceOp op;
    memset(&op, 0, sizeof op);
    op.minus = minus;
    if (n > kPieceMax) n = kPieceMax;
    for (int j = a; j < a + n && op.n < kPieceLookups;) {
        if (j % 2 == 0 && j + 1 < a + n) { op.pair[op.n] = 1; op.idx[op.n] = j / 2; j += 2; }
        else { op.pair[op.n] = 0; op.idx[op.n] = j; j += 1; }
        ++op.n;
    }
    return op;
}

static int launch_pieces(const mg_genome* g, const mg_queries* q, const GeneralCtx* d_ctx, const PieceOp& op, int sms, int smem_max,
                         int64_t a, int64_t b, int64_t own_a, int64_t own_b, cudaStream_t s) {
    const int L = q->len, SG = 16 * ((L + 1) / 2) + 1 + 4 * L + 1;
    const int max_groups = (smem_max - 1024) / (SG * 4);
    const int gpc = q->n_groups < max_groups ? q->n_groups : max_groups;
    const size_t smem = (size_t)gpc * SG * 4;
    int64_t grid = (b - a + kPieceThreads - 1) / kPieceThreads;
    if (grid > sms) grid = sms;
    switch (op.n) {
#define MG_PIECES(TT) case TT: \
        MG_CUDA(cudaFuncSetAttribute(pieces_kernel<TT, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        pieces_kernel<TT, 3><<<(unsigned)grid, kPieceThreads, smem, s>>>(g->view(), q->view(), d_ctx, op, gpc, a, b, own_a, own_b); break;
        MG_PIECES(1) MG_PIECES(2) MG_PIECES(3) MG_PIECES(4) MG_PIECES(5) MG_PIECES(6) MG_PIECES(7) MG_PIECES(8) MG_PIECES(9)
#undef MG_PIECES
        default: set_error("piece with %d lookups", op.n); return MG_ERR_ARG;
    }
    MG_CUDA(cudaGetLastError());
    return MG_OK;
}

// ----------------------------------------------------------------------------------------- verifier pass
__global__ void __launch_bounds__(128)
verify_kernel(GenomeView gv, QueriesView qv, const GeneralCtx* __restrict__ ctx, unsigned long long n_cand,
              uint32_t* __restrict__ counts) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_cand) return;
    const Cand c = ctx->cand[i];
    const int gi = c.q < qv.n ? c.q : c.q - qv.n;
    if (c.tag & kCandTagged) {
        // exact-piece prefilter: the same anchor is reached from up to (pieces x (2 Gp + 1)) (piece, window) pairs.  It is
        // counted by the first pair in (piece, offset) order whose piece matches - every pair's window is scanned by the
        // launch that owns the anchor, so exactly one queued copy passes this test.
        const int L = qv.len, Gp = ctx->Gp;
        const int my_piece = (c.tag >> 8) & 0xFF, my_d = (c.tag & 0xFF) - Gp;
        const bool plus = c.q < qv.n;
        const uint8_t* codes = qv.codes + (size_t)c.q * L;               // oriented query (rows N.. are reverse complements)
        for (int k = 0; k <= my_piece; ++k) {
            const int pa = plus ? ctx->piece_start[k] : L - ctx->piece_start[k] - ctx->piece_len[k];
            const int pn = min(ctx->piece_len[k], kPieceMax);
            for (int d = -Gp; d <= Gp; ++d) {
                if (k == my_piece && d >= my_d) break;
                const int64_t pw = plus ? c.e - d - L : gv.glen - c.e + d;   // the window whose hit of piece k reaches anchor e with offset d
                if (pw < 0 || pw > gv.glen - L) continue;
                bool same = true;
                for (int j = pa; j < pa + pn && same; ++j) same = codes[j] < 4 && base_at(gv, pw + j) == (int)codes[j];
                if (same) return;                                         // an earlier pair owns this anchor
            }
        }
    }
    if (verify_anchor(gv, *ctx, qv.codes + (size_t)gi * qv.len, qv.packed + gi, c.q < qv.n ? 1 : -1, c.e)) atomicAdd(counts + c.q, 1u);
}

// ----------------------------------------------------------------------------------------- every cell
__global__ void __launch_bounds__(128)
allcells_kernel(GenomeView gv, QueriesView qv, const GeneralCtx* __restrict__ ctx, int strand, int64_t eb, int64_t ee,
                uint32_t* __restrict__ counts) {
    const int L = qv.len, N = qv.n;
    const int64_t n_anchor = ee - eb;
    const int64_t total = n_anchor * N;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = eb + idx % n_anchor;
        const int gi = (int)(idx / n_anchor);
        if (verify_anchor(gv, *ctx, qv.codes + (size_t)gi * L, qv.packed + gi, strand, e))
            atomicAdd(counts + (strand > 0 ? gi : N + gi), 1u);
    }
}

static void fill_side(const mg_pam_side& in, PamSideDev& out) {
    out.n_alts = in.n_alts;
    memcpy(out.len, in.len, sizeof out.len);
    memcpy(out.allow4, in.allow4, sizeof out.allow4);
}

int launch_screen_general(const mg_genome* g, const mg_queries* q, const mg_criteria* crit, const mg_pam* pam,
                          int64_t wb, int64_t we, uint32_t* d_counts, cudaStream_t s) {
    const int L = q->len;
    GeneralCtx h;
    memset(&h, 0, sizeof h);
    h.L = L; h.M = crit->max_mismatch; h.Gp = crit->max_gap; h.pamless = crit->pamless;
    h.n_units = crit->n_units; h.n_ops = crit->n_ops; h.n_grna = q->n;
    h.prune_pattern = crit->max_gap <= 1;
    for (int i = 0; i < crit->n_units; ++i) {
        const mg_pattern_unit& u = crit->units[i];
        h.units[i] = UnitDev{u.n, u.chars, u.start, u.end, u.rvs, u.add_unaligned_m, u.add_unaligned_g,
                             py_slice_mask(u.start, u.end, L, L)};
    }
    {   // validate the postfix program
        int depth = 0;
        for (int i = 0; i < crit->n_ops; ++i) {
            const int op = crit->ops[i];
            if (op >= 0) { if (op >= crit->n_units) { set_error("pattern op %d references unit %d", i, op); return MG_ERR_ARG; } ++depth; }
            else { if ((op != MG_OP_AND && op != MG_OP_OR) || depth < 2) { set_error("malformed pattern program at op %d", i); return MG_ERR_ARG; } --depth; }
            h.ops[i] = op;
        }
        if (crit->n_units > 0 && depth != 1) { set_error("malformed pattern program"); return MG_ERR_ARG; }
    }
    if (!crit->pamless) {
        MG_CHECK_ARG(pam->five.n_alts >= 0 && pam->five.n_alts <= MG_MAX_PAM_ALTS && pam->three.n_alts >= 0 && pam->three.n_alts <= MG_MAX_PAM_ALTS);
        fill_side(pam->five, h.five);
        fill_side(pam->three, h.three);
        h.five_max = pam->five_maxlen;
        h.three_max = pam->three_maxlen;
    }
    MG_CHECK_ARG(crit->n_pieces >= 0 && crit->n_pieces <= MG_MAX_PIECES);
    h.n_pieces = crit->n_units ? crit->n_pieces : 0;
    for (int k = 0; k < h.n_pieces; ++k) {
        const int a = crit->piece_start[k], n = crit->piece_len[k];
        MG_CHECK_ARG(a >= 0 && n >= 1 && a + n <= L);
        uint32_t m = 0;
        for (int j = a; j < a + n; ++j) m |= 1u << j;
        uint32_t r = 0;
        for (int j = a; j < a + n; ++j) r |= 1u << (L - 1 - j);      // the same bases in a reverse-complemented query
        h.piece_fwd[k] = m;
        h.piece_rev[k] = r;
        h.piece_start[k] = a;
        h.piece_len[k] = n;
    }
    // candidate queue between the prefilter and the verifier pass
    h.cand_cap = 64ull << 20;                                   // 1 GB of 16-byte entries
    if (const char* cap_env = getenv("MG_QUEUE_CAP")) {          // tests shrink the queue to exercise the re-cut path
        const long long v = atoll(cap_env);
        if (v >= 8) h.cand_cap = (unsigned long long)v;
    }
    MG_CUDA(dev_alloc((void**)&h.cand, sizeof(Cand) * h.cand_cap, s));
    if (dev_alloc((void**)&h.cand_count, 8, s) != cudaSuccess || cudaMemsetAsync(h.cand_count, 0, 8, s) != cudaSuccess) {
        dev_free(h.cand, s); dev_free(h.cand_count, s);
        set_error("candidate counter: %s", cudaGetErrorString(cudaGetLastError()));
        return MG_ERR_NOMEM;
    }
    int K = crit->n_units ? crit->prefilter_edits : crit->max_mismatch + crit->max_gap;
    if (K >= L) K = -1;
    const char* pf_env = getenv("MG_PREFILTER");        // "none": hand EVERY anchor to the verifier (tests: prefilter soundness)
    const bool no_prefilter = pf_env && strcmp(pf_env, "none") == 0;
    // gapped patterns: the verifier drops a candidate whose cheapest full-length alignment within the gap cap already
    // exceeds the host-proved bound, before the exhaustive walk (not in the soundness-test mode, which must not rely on it)
    h.edit_bound = (crit->n_units > 0 && crit->max_gap >= 1 && K >= 0 && !no_prefilter) ? K : -1;
    GeneralCtx* d_ctx = nullptr;
    int rc = MG_OK;
    cudaError_t e = cudaMalloc(&d_ctx, sizeof h);
    if (e != cudaSuccess) { dev_free(h.cand, s); dev_free(h.cand_count, s); set_error("criteria buffer: %s", cudaGetErrorString(e)); return MG_ERR_NOMEM; }
    e = cudaMemcpyAsync(d_ctx, &h, sizeof h, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);     // h is on this stack frame
    if (e != cudaSuccess) { cudaFree(d_ctx); dev_free(h.cand, s); dev_free(h.cand_count, s); set_error("copying criteria: %s", cudaGetErrorString(e)); return MG_ERR_CUDA; }

    // anchors (virtual ends, exclusive) owned by the window range [wb, we): e = p + L
    const bool by_pieces = !no_prefilter && h.n_pieces > 0;
    const bool by_hamming = !no_prefilter && !by_pieces && crit->max_gap == 0 && K >= 0;
    const bool by_myers = !no_prefilter && !by_pieces && !by_hamming && K >= 0;
    int dev = 0, sms = 0, smem_max = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);

    // prefilter over windows [a, b): queues survivors.  strand -1 anchors of a window [p, p+L) end, in
    // reverse-complement coordinates, at glen - p; so windows [a, b) own e' in [glen - b + 1, glen - a + 1).
    uint32_t* d_rows_tab = nullptr;                                   // tables of the bit-sliced edit-distance kernel
    const char* gapped_env = getenv("MG_GAPPED_PREFILTER");       // "myers": the one-pattern-per-thread kernel (A/B)
    const bool use_rows = !(gapped_env && strcmp(gapped_env, "myers") == 0);
    // Only the edit-distance rows kernel can also be cut by gRNA group [g0, g1) (groups of 32 gRNA): its threads own
    // long runs of text, so cutting the text would multiply the L+K warm-up characters every run starts with.
    const bool group_slices = by_myers && use_rows;
    const int n_query_groups = group_slices ? (q->n + 31) / 32 : 1;
    auto prefilter = [&](int64_t a, int64_t b, int g0, int g1) -> int {
        if (by_pieces) {
            for (int k = 0; k < h.n_pieces; ++k)
                for (int minus = 0; minus < 2; ++minus) {
                    const int pa = crit->piece_start[k], pn = crit->piece_len[k];
                    PieceOp op = make_piece_op(minus ? L - pa - pn : pa, pn, minus);
                    op.piece = k;
                    // Gp extra windows on both sides: an anchor of [a, b) may only be reachable from a neighbour's piece hit
                    const int64_t sa = std::max<int64_t>(0, a - crit->max_gap), sb = std::min<int64_t>(g->glen - L + 1, b + crit->max_gap);
                    const int r2 = launch_pieces(g, q, d_ctx, op, sms, smem_max, sa, sb, a, b, s);
                    if (r2 != MG_OK) return r2;
                }
            return MG_OK;
        }
        if (by_hamming) return launch_screen_bitsliced_general(g, q, K, d_ctx, a, b, d_counts, s);
        for (int z = 0; z < 2; ++z) {
            const int strand = z == 0 ? 1 : -1;
            const int64_t eb = z == 0 ? a + L : g->glen - b + 1, ee = z == 0 ? b + L : g->glen - a + 1;
            const int64_t n_anchor = ee - eb;
            if (n_anchor <= 0) continue;
            if (by_myers && use_rows) {
                const int r2 = launch_gapped_rows(g, q, d_ctx, K, strand, g0, g1, eb, ee, &d_rows_tab, s);
                if (r2 != MG_OK) return r2;
            } else if (by_myers) {
                dim3 grid((unsigned)((n_anchor + kMyersChunk - 1) / kMyersChunk), (unsigned)((q->n + kMyersThreads - 1) / kMyersThreads), 1);
                myers_kernel<<<grid, kMyersThreads, 0, s>>>(g->view(), q->view(), d_ctx, K, strand, eb, ee, d_counts);
            } else {
                int64_t grid = (n_anchor * q->n + 127) / 128;
                if (grid > sms * 16) grid = sms * 16;
                allcells_kernel<<<(unsigned)grid, 128, 0, s>>>(g->view(), q->view(), d_ctx, strand, eb, ee, d_counts);
            }
        }
        return MG_OK;
    };

    // Run [wb, we) x [all gRNA groups) in slices whose survivors fit the queue; an overflowing slice is re-run in as
    // many parts as its survivor count calls for (nothing of it has been verified yet, so no double counting).
    struct Slice { int64_t a, b; int g0, g1; };
    auto cut = [&](const Slice& r, int64_t parts, std::vector<Slice>& out) {      // pushes the parts in reverse order
        const int64_t ng = r.g1 - r.g0, len = r.b - r.a;
        const int64_t gparts = parts < ng ? parts : ng;                           // groups first, then text
        int64_t tparts = (parts + gparts - 1) / gparts;
        if (tparts > len) tparts = len;
        for (int64_t i = gparts - 1; i >= 0; --i)
            for (int64_t j = tparts - 1; j >= 0; --j)
                out.push_back(Slice{r.a + len * j / tparts, r.a + len * (j + 1) / tparts,
                                    (int)(r.g0 + ng * i / gparts), (int)(r.g0 + ng * (i + 1) / gparts)});
    };
    unsigned long long total_cand = 0;
    int n_slices = 0;
    const bool trace = getenv("MG_TRACE") != nullptr;      // adds a synchronisation per slice to time the two passes
    double t_prefilter = 0.0, t_verify = 0.0;
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    std::vector<Slice> todo;
    Slice rest{0, 0, 0, 0};
    bool have_rest = false;
    // Large gRNA sets under the one-gap no-pattern criterion: the seed-and-join screen (screen_seedjoin.cu) counts every anchor
    // whose neighbourhood is free of invalid bases; what is left for the prefilter + verifier below are the short "dirty" window
    // ranges around contig ends and N runs.
    bool seedjoin_done = false;
    if (!no_prefilter && by_myers && use_rows && seedjoin_eligible(q, crit)) {
        std::vector<std::pair<int64_t, int64_t>> dirty;
        const double t0 = trace ? now() : 0.0;
        const int r2 = launch_seedjoin(g, q, crit, d_ctx, wb, we, d_counts, s, dirty);
        if (r2 == MG_OK) {
            seedjoin_done = true;
            for (auto it = dirty.rbegin(); it != dirty.rend(); ++it) todo.push_back(Slice{it->first, it->second, 0, n_query_groups});
            if (trace) { cudaStreamSynchronize(s); fprintf(stderr, "[mg_screen general] seed-and-join %.1f ms, %zu dirty window ranges left to the DP path\n", (now() - t0) * 1e3, dirty.size()); }
        } else if (r2 != MG_ERR_UNSUPPORTED) {
            rc = r2;
        }
    }
    if (!seedjoin_done && rc == MG_OK) {   // a small probe slice first: its survivor density sizes the remaining slices, so that an overflow (a wasted
        // prefilter pass) only happens when the density is very uneven
        const int64_t len = we - wb, probe = len / 64;
        int64_t probe_min = 1 << 16;
        if (const char* pm = getenv("MG_PROBE_MIN")) probe_min = atoll(pm) > 0 ? atoll(pm) : probe_min;
        if (probe >= probe_min) {
            rest = Slice{wb + probe, we, 0, n_query_groups};
            have_rest = true;
            todo.push_back(Slice{wb, wb + probe, 0, n_query_groups});
        } else {
            todo.push_back(Slice{wb, we, 0, n_query_groups});
        }
    }
    while (!todo.empty() && rc == MG_OK) {
        const Slice r = todo.back();
        todo.pop_back();
        if (r.b <= r.a || r.g1 <= r.g0) continue;
        if ((e = cudaMemsetAsync(h.cand_count, 0, 8, s)) != cudaSuccess) break;      // falls through to the common clean-up
        const double t0 = trace ? now() : 0.0;
        rc = prefilter(r.a, r.b, r.g0, r.g1);
        if (rc != MG_OK) break;
        ++n_slices;
        unsigned long long n_cand = 0;
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaMemcpyAsync(&n_cand, h.cand_count, 8, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) break;
        if (trace) t_prefilter += now() - t0;
        if (n_cand > h.cand_cap) {
            if (r.b - r.a <= 1 && r.g1 - r.g0 <= 1) { set_error("candidate queue too small for a single window (%llu survivors)", n_cand); rc = MG_ERR_NOMEM; break; }
            int64_t parts = (int64_t)((n_cand + h.cand_cap / 2 - 1) / (h.cand_cap / 2));
            if (parts < 2) parts = 2;
            cut(r, parts, todo);
            continue;
        }
        total_cand += n_cand;
        if (have_rest) {         // the first slice that fitted (the probe or, if it overflowed, a part of it): cut the
            have_rest = false;   // rest into slices of ~cap/3 expected survivors, by its density per (position, group)
            const double density = (double)n_cand / ((double)(r.b - r.a) * (double)(r.g1 - r.g0));
            const double expect = density * (double)(rest.b - rest.a) * (double)(rest.g1 - rest.g0);
            cut(rest, (int64_t)(expect / ((double)h.cand_cap / 3.0)) + 1, todo);
        }
        if (n_cand > 0) {
            const double t1 = trace ? now() : 0.0;
            verify_kernel<<<(unsigned)((n_cand + 127) / 128), 128, 0, s>>>(g->view(), q->view(), d_ctx, n_cand, d_counts);
            e = cudaGetLastError();
            if (e != cudaSuccess) break;
            if (trace) { cudaStreamSynchronize(s); t_verify += now() - t1; }
        }
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);      // d_ctx and the queue must outlive the kernels
    if (trace) fprintf(stderr, "[mg_screen general] prefilter=%s K=%d survivors=%llu slices=%d prefilter %.1f ms verify %.1f ms\n",
                       by_pieces ? "pieces" : by_hamming ? "hamming" : by_myers ? (use_rows ? "edit-rows" : "myers") : "all-cells", K, total_cand, n_slices,
                       t_prefilter * 1e3, t_verify * 1e3);
    dev_free(d_rows_tab, s);
    dev_free(h.cand, s);
    dev_free(h.cand_count, s);
    cudaFree(d_ctx);
    if (e != cudaSuccess) { set_error("general screen: %s", cudaGetErrorString(e)); return MG_ERR_CUDA; }
    return rc;
}

}  // namespace mg
