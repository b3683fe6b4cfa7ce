// K5: gapped prefilter - edit distance <= K of every gRNA against every text position, bit-sliced ACROSS gRNAs.
//
// Necessary condition for the reference's gapped criteria (MINORg.py:2018-2044 with ot_gap >= 2, or a pattern
// whose bound the host derived): an alignment with mismatches + gap characters + unaligned positions <= K extends
// to a full-length alignment of the gRNA ending at the same diagonal with unit-cost edit distance <= K.
//
// Formulation.  Myers/Hyyro's bit-vector recurrences keep the DP column of ONE pattern in a machine word and
// resolve the vertical dependency with an addition.  Here the word holds the SAME DP cell of 32 different
// gRNAs (the query groups of the scan kernel), the rows of the column live in registers (vertical deltas
// +1 / -1 as two bit-planes) and the vertical dependency is simply the order of the unrolled row loop.  Per
// cell (Hyyro 2001), with nD0 = ~D0 ("the cell does NOT equal its diagonal neighbour"):
//        nD0 = mis & ~Mv & ~Mh_in
//        Mh  = Pv & ~nD0                  Ph  = Mv | (nD0 & ~Pv)          (horizontal delta, handed to the next row)
//        Mv' = Ph_in & ~nD0               Pv' = Mh_in | (nD0 & ~Ph_in)    (vertical delta, kept for the next character)
// = 5 LOP3 per 32 cells, all on the ALU pipe, which is what bounds the kernel (measured: 150 ALU-pipe
// instructions per character and group = 300 issue cycles per scheduler, the LSU and the FMA pipe idle).
//
// Pipe balancing.  Two of the five are rewritten as carry-free INTEGER arithmetic, which runs on the otherwise
// idle FMA pipe as IMAD.  With X = Mv - Pv (the signed vertical delta, kept as a third word per row) and
// Dh = Ph - Mh (the signed horizontal delta, handed down next to Ph and Mh):
//        Dh_out = nD0 + X                 Ph_out = Dh_out + Mh_out
//        X'     = Dh_in - nD0             Pv'    = Mv' - X'
// Each identity holds bit by bit with a per-bit result in {0, 1} (nD0 = 1 forces Mv = Mh_in = 0 and D0 = 0, so
// Ph = 1 - Pv and Pv' = 1 - Ph_in; nD0 = 0 gives Ph = Mv, Mh = Pv, Pv' = Mh_in, Mv' = Ph_in), so no carry or borrow
// survives in the final word although the intermediate X and Dh are signed per bit.  tests/test_host_mirror.py
// checks the identities exhaustively.  The multipliers +1 / -1 arrive as kernel arguments so that the
// compiler cannot fold the IMADs back into ALU adds.  Per row: 3 LOP3 (ALU) + 4 IMAD (FMA); per character and group
// 82 ALU + 82 FMA + 5 LSU instructions = ~170 issue slots, all three limits within a few percent of each other.
// Measured (B200, L = 20, K = 2, 4096 gRNA x 135 Mb): 204 cycles per warp-character = 5.65 Tcell/s; five-LOP3 form
// with the same loads 232 cycles = 4.96; converting only part of the rows (10..16 of 20) gains nothing (4.9-5.1).
//
// The mismatch words of one character are contiguous (tab[g][code][row]) and fetched with 128-bit shared loads:
// the same LSU time as 32-bit loads (measured, mg_microbench) but a quarter of the issue slots.
// The bottom-row score is a bit-sliced up/down counter biased so that "score <= K" is the complement of its top bit.
// A thread owns a run of consecutive text characters (plus L+K warm-up characters: a match of cost <= K spans at
// most L+K characters) and walks it once per query group; survivors are queued for the exact verifier.
#include <algorithm>

#include "common.cuh"
#include "verify.cuh"

namespace mg {

// threads per CTA (one CTA per SM): the DP column of a group is 3L registers, so long gRNAs get 255 registers/thread
template <int L> struct RowsThreads { static constexpr int value = L <= 24 ? 512 : 256; };
constexpr int kRowsCodes = 8;                          // table columns per row: A C G T + four all-ones (invalid base)
constexpr int kRowsQueue = 6144;                       // survivor entries (8 B) staged per CTA between two drains (48 KB)

__device__ __forceinline__ uint32_t lop3_and_nn(uint32_t a, uint32_t b, uint32_t c) {       // a & ~b & ~c
    uint32_t r; asm("lop3.b32 %0, %1, %2, %3, 0x10;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
}
__device__ __forceinline__ uint32_t lop3_a_or_b_andn(uint32_t a, uint32_t b, uint32_t c) {  // a | (b & ~c)
    uint32_t r; asm("lop3.b32 %0, %1, %2, %3, 0xF4;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
}
__device__ __forceinline__ uint32_t imad(uint32_t a, uint32_t m, uint32_t c) {               // a * m + c on the FMA pipe
    uint32_t r; asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(m), "r"(c)); return r;
}

// 16 consecutive characters of the strand's text as a 2-bit code stream (character k at bits 2k) and their invalid
// bits pre-shifted to bit 2 (character k at bit k+2), so that  (w & 3) | (iv4 & 4)  is the table column.
// Strand -1 reads the forward words back to front: glen is a multiple of 16, so strand block sb is forward block
// glen/16-1-sb with its 2-bit fields reversed and complemented.
__device__ __forceinline__ void load_text_block(const GenomeView& gv, int strand, int64_t sb, uint32_t& w, uint32_t& iv4) {
    const int64_t fb = strand > 0 ? sb : (gv.glen >> 4) - 1 - sb;
    uint32_t x = __ldg(gv.bases + fb);
    uint32_t v = (__ldg(gv.invalid + (fb >> 1)) >> ((unsigned)(fb & 1) * 16u)) & 0xFFFFu;
    if (strand < 0) {
        x = __brev(x);
        x = ~(((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1));
        v = __brev(v) >> 16;
    }
    w = x;
    iv4 = v << 2;
}

template <int L, int NF>                              // rows [0, NF) use the arithmetic form, the rest five LOP3
struct RowsState {
    static constexpr bool FMA = NF > 0;
    static constexpr int NB = 6;                       // score counter bits (scores 0..32 + bias)
    uint32_t Pv[L], Mv[L], X[NF > 0 ? NF : 1], sc[NB];

    __device__ __forceinline__ void init(int K) {
#pragma unroll
        for (int i = 0; i < L; ++i) { Pv[i] = 0xFFFFFFFFu; Mv[i] = 0u; if (i < NF) X[i] = 1u; }   // X = Mv - Pv = 0 - (-1)
        // score counter bias: c = score + 2^(NB-1) - K - 1, so score <= K  <=>  top bit of c is clear
        const uint32_t start = (uint32_t)(L + (1 << (NB - 1)) - K - 1);
#pragma unroll
        for (int b = 0; b < NB; ++b) sc[b] = ((start >> b) & 1u) ? 0xFFFFFFFFu : 0u;
    }

    // one text character: mis[i] = mismatch word of row i.  Returns the word of queries whose score is <= K here.
    __device__ __forceinline__ uint32_t step(const uint32_t (&mis)[(L + 3) / 4 * 4], uint32_t one, uint32_t minus_one) {
        uint32_t Ph = 0u, Mh = 0u, Dh = 0u;                // horizontal delta entering row 0: D[0][j] = 0 for all j
#pragma unroll
        for (int i = 0; i < L; ++i) {
            const uint32_t nD0 = lop3_and_nn(mis[i], Mv[i], Mh);
            const uint32_t Mh_out = Pv[i] & ~nD0;
            const uint32_t Mv_new = Ph & ~nD0;
            if (i < NF) {
                const uint32_t Dh_out = imad(nD0, one, X[i]);
                const uint32_t Ph_out = imad(Mh_out, one, Dh_out);
                X[i] = imad(nD0, minus_one, Dh);
                Pv[i] = imad(X[i], minus_one, Mv_new);
                Dh = Dh_out;
                Ph = Ph_out;
            } else {
                const uint32_t Ph_out = lop3_a_or_b_andn(Mv[i], nD0, Pv[i]);
                Pv[i] = lop3_a_or_b_andn(Mh, nD0, Ph);
                Ph = Ph_out;
            }
            Mv[i] = Mv_new;
            Mh = Mh_out;
        }
        // score += Ph - Mh (bottom row), bit-sliced ripple
        uint32_t carry = Ph | Mh;
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const uint32_t old = sc[b];
            sc[b] = old ^ carry;
            carry &= old ^ Mh;                             // increment ripples through 1s, decrement through 0s
        }
        return ~sc[NB - 1];
    }
};

template <int L, int NF>
__global__ void __launch_bounds__(RowsThreads<L>::value, 1)
gapped_rows_kernel(GenomeView gv, QueriesView qv, const uint32_t* __restrict__ tabr, const GeneralCtx* __restrict__ ctx,
                   int K, int strand, int groups_per_chunk, int groups_per_cta, int g_begin, int g_end, int64_t eb, int64_t ee, int run,
                   uint32_t one, uint32_t minus_one) {
    constexpr int LP = (L + 3) / 4 * 4;                // rows padded to whole 128-bit loads
    constexpr int SG = kRowsCodes * LP;                // words per group: [code][row]
    constexpr int kRowsThreads = RowsThreads<L>::value;
    extern __shared__ __align__(16) uint32_t tab[];
    const uint32_t tab_sh = (uint32_t)__cvta_generic_to_shared(tab);
    // Survivors are staged in a per-CTA shared-memory queue (one ATOMS per survivor instead of a global atomic round
    // trip that stalls the warp for ~600 cycles) and drained to the verifier's global queue once per gRNA group with ONE
    // global atomic per CTA and coalesced stores.  An entry is (oriented query, anchor - first text position of the CTA).
    uint2* const cq = reinterpret_cast<uint2*>(tab + (size_t)groups_per_chunk * SG);
    __shared__ unsigned int cq_count;
    __shared__ unsigned long long cq_base;
    const int N = qv.n;
    const int warm = L + K;
    // anchors e in [eb, ee) <-> text positions pos = e - 1 in [eb - 1, ee - 1)
    const int64_t p_first = eb - 1 + ((int64_t)blockIdx.x * kRowsThreads + threadIdx.x) * run;
    const int64_t p_last = min(ee - 1, p_first + run);
    const bool active = p_first < p_last;
    const int64_t p_begin = max((int64_t)0, p_first - warm);       // the initial column is exact at the start of the text
    const int64_t e_base = eb - 1 + (int64_t)blockIdx.x * kRowsThreads * run - warm;   // <= every anchor this CTA reports
    if (threadIdx.x == 0) cq_count = 0u;
    // a short text range with many gRNA groups is cut over blockIdx.y by GROUP range (groups_per_cta each)
    g_begin += (int)blockIdx.y * groups_per_cta;
    g_end = min(g_end, g_begin + groups_per_cta);

    for (int cg = g_begin; cg < g_end; cg += groups_per_chunk) {
        const int ng = min(groups_per_chunk, g_end - cg);
        __syncthreads();
        for (int i = threadIdx.x; i < ng * SG; i += kRowsThreads) tab[i] = __ldg(tabr + (size_t)cg * SG + i);
        __syncthreads();
        for (int g = 0; g < ng; ++g) {
            const uint32_t t = tab_sh + (uint32_t)g * (SG * 4);     // shared-window byte address of this group's table
            const int q0 = (cg + g) * 32;
            // only the gRNA-orientation queries (index < N) take part: the strand is selected by the text
            const uint32_t qmask = (q0 + 32 <= N) ? 0xFFFFFFFFu : ((q0 < N) ? ((1u << (N - q0)) - 1u) : 0u);
            RowsState<L, NF> st;
            st.init(K);
            int64_t pos = active ? p_begin : p_last;
            while (pos < p_last) {
                // characters up to the end of this 16-character block; the warm-up part reports nothing
                const int k0 = (int)(pos & 15);
                const bool warmup = pos < p_first;
                const int64_t stop = warmup ? p_first : p_last;
                const int n = (int)min((int64_t)(16 - k0), stop - pos);
                uint32_t w, iv4;
                load_text_block(gv, strand, pos >> 4, w, iv4);
                w >>= 2 * k0;
                iv4 >>= k0;
                const uint32_t live = warmup ? 0u : qmask;
                for (int k = 0; k < n; ++k) {
                    const uint32_t code = (w & 3u) | (iv4 & 4u);
                    w >>= 2;
                    iv4 >>= 1;
                    const uint32_t row = t + code * (LP * 4);
                    uint32_t mis[LP];
#pragma unroll
                    for (int v = 0; v < LP / 4; ++v)
                        asm("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                            : "=r"(mis[4 * v]), "=r"(mis[4 * v + 1]), "=r"(mis[4 * v + 2]), "=r"(mis[4 * v + 3]) : "r"(row + 16 * v));
                    uint32_t h = st.step(mis, one, minus_one) & live;
                    while (h) {
                        const int i = __ffs(h) - 1;
                        h &= h - 1;
                        const int q = strand > 0 ? q0 + i : N + q0 + i;
                        const unsigned int slot = atomicAdd(&cq_count, 1u);
                        if (slot < (unsigned int)kRowsQueue) cq[slot] = make_uint2((uint32_t)q, (uint32_t)(pos + k + 1 - e_base));
                        else emit_candidate(*ctx, q, pos + k + 1);          // queue full: the slow direct path
                    }
                }
                pos += n;
            }
            // drain (every thread of the CTA takes part; the threads walk equally long runs, so they arrive together)
            __syncthreads();
            const unsigned int n_q = min(cq_count, (unsigned int)kRowsQueue);
            __syncthreads();                                        // everyone has read the count before it is reset
            if (n_q) {
                if (threadIdx.x == 0) { cq_base = atomicAdd(ctx->cand_count, (unsigned long long)n_q); cq_count = 0u; }
                __syncthreads();
                const unsigned long long base = cq_base;
                for (unsigned int i = threadIdx.x; i < n_q; i += kRowsThreads) {
                    const uint2 c = cq[i];
                    if (base + i < ctx->cand_cap) ctx->cand[base + i] = Cand{(int32_t)c.x, 0, e_base + (int64_t)c.y};
                }
                __syncthreads();                                    // the queue is free again
            }
        }
    }
}

template <int L>
static int launch_rows(const mg_genome* g, const mg_queries* q, const uint32_t* d_tab, const GeneralCtx* d_ctx, int K,
                       int strand, int g_begin, int g_end, int64_t eb, int64_t ee, cudaStream_t s) {
    int dev = 0, sms = 0, smem_max = 0;
    MG_CUDA(cudaGetDevice(&dev));
    MG_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    MG_CUDA(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    const int SG = kRowsCodes * ((L + 3) / 4 * 4);
    constexpr int kRowsThreads = RowsThreads<L>::value;
    const int max_groups = (smem_max - 1024 - kRowsQueue * 8) / (SG * 4);
    const int n_groups = g_end - g_begin;
    static const bool alu_only = [] { const char* e = getenv("MG_ROWS"); return e && strcmp(e, "alu") == 0; }();   // A/B: all five ops as LOP3
    auto kern = alu_only ? gapped_rows_kernel<L, 0> : gapped_rows_kernel<L, L>;
    const int64_t n_anchor = ee - eb;
    // characters per thread: long enough to amortise the L+K warm-up, short enough to fill the machine; whole waves of
    // CTAs (one CTA per SM), and a multiple of 16 so that every lane of a warp sits at the same offset inside its
    // 16-character block
    const int64_t per_wave = (int64_t)sms * kRowsThreads;
    const int64_t waves = (n_anchor + per_wave * 4096 - 1) / (per_wave * 4096);
    int64_t run = (n_anchor + per_wave * waves - 1) / (per_wave * waves);
    if (run < 64) run = 64;
    run = (run + 15) / 16 * 16;
    // A short text range (the dirty ranges the seed-and-join screen leaves to this path: ~10^2 windows around a contig end or
    // an N run) cannot fill the machine with text: shorter runs, and the gRNA groups are cut over blockIdx.y instead.
    int64_t threads = (n_anchor + run - 1) / run;
    unsigned grid = (unsigned)((threads + kRowsThreads - 1) / kRowsThreads);
    int gy = 1;
    if ((int)grid * 2 <= sms && n_groups > 32) {
        if (threads < 64) { run = 16; threads = (n_anchor + run - 1) / run; grid = (unsigned)((threads + kRowsThreads - 1) / kRowsThreads); }
        gy = std::min((n_groups + 15) / 16, (4 * sms + (int)grid - 1) / (int)grid);
        if (gy < 1) gy = 1;
        if (gy > 65535) gy = 65535;
    }
    const int groups_per_cta = (n_groups + gy - 1) / gy;
    const int gpc = groups_per_cta < max_groups ? groups_per_cta : max_groups;
    const size_t smem = (size_t)gpc * SG * 4 + (size_t)kRowsQueue * 8;       // tables, then the survivor queue
    MG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<dim3(grid, (unsigned)gy, 1), kRowsThreads, smem, s>>>(g->view(), q->view(), d_tab, d_ctx, K, strand, gpc, groups_per_cta, g_begin,
                                                                   g_end, eb, ee, (int)run, 1u, 0xFFFFFFFFu);
    MG_CUDA(cudaGetLastError());
    return MG_OK;
}

// tabr[g][code][row] (code stride = rows padded to a multiple of 4): bit k = 1 iff gRNA 32g+k does NOT have base `code`
// at position `row` (or is padding); codes 4..7 (invalid genome base) and the padding rows are all ones.
__global__ void rows_table_kernel(const uint8_t* __restrict__ codes, int n, int len, int n_groups, uint32_t* __restrict__ tabr) {
    const int lp = (len + 3) / 4 * 4, stride = kRowsCodes * lp;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_groups * stride) return;
    const int g = idx / stride, slot = idx % stride;
    const int b = slot / lp, i = slot % lp;
    uint32_t w = 0xFFFFFFFFu;
    if (b < 4 && i < len) {
        w = 0;
        for (int k = 0; k < 32; ++k) {
            const int qi = g * 32 + k;
            const bool mism = (qi >= n) || (codes[(size_t)qi * len + i] != b);
            w |= (mism ? 1u : 0u) << k;
        }
    }
    tabr[idx] = w;
}

int launch_gapped_rows(const mg_genome* g, const mg_queries* q, const GeneralCtx* d_ctx, int K, int strand, int g_begin,
                       int g_end, int64_t eb, int64_t ee, uint32_t** d_tab_cache, cudaStream_t s) {
    const int n_groups_used = (q->n + 31) / 32;
    if (g_end > n_groups_used) g_end = n_groups_used;
    if (q->n == 0 || ee <= eb || g_begin >= g_end) return MG_OK;
    if (*d_tab_cache == nullptr) {
        const int words = n_groups_used * kRowsCodes * ((q->len + 3) / 4 * 4);
        MG_CUDA(dev_alloc((void**)d_tab_cache, (size_t)words * 4 + 16, s));
        rows_table_kernel<<<(words + 255) / 256, 256, 0, s>>>(q->d_codes, q->n, q->len, n_groups_used, *d_tab_cache);
        MG_CUDA(cudaGetLastError());
    }
    switch (q->len) {
#define MG_ROWS(LL) case LL: return launch_rows<LL>(g, q, *d_tab_cache, d_ctx, K, strand, g_begin, g_end, eb, ee, s);
        MG_ROWS(1) MG_ROWS(2) MG_ROWS(3) MG_ROWS(4) MG_ROWS(5) MG_ROWS(6) MG_ROWS(7) MG_ROWS(8) MG_ROWS(9) MG_ROWS(10)
        MG_ROWS(11) MG_ROWS(12) MG_ROWS(13) MG_ROWS(14) MG_ROWS(15) MG_ROWS(16) MG_ROWS(17) MG_ROWS(18) MG_ROWS(19)
        MG_ROWS(20) MG_ROWS(21) MG_ROWS(22) MG_ROWS(23) MG_ROWS(24) MG_ROWS(25) MG_ROWS(26) MG_ROWS(27) MG_ROWS(28)
        MG_ROWS(29) MG_ROWS(30) MG_ROWS(31) MG_ROWS(32)
#undef MG_ROWS
        default: set_error("gRNA length %d unsupported (1..32)", q->len); return MG_ERR_UNSUPPORTED;
    }
}

}  // namespace mg
