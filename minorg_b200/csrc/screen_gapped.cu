// K5: gapped prefilter - edit distance <= K of every gRNA against every text position, bit-sliced ACROSS gRNAs.
//
// Necessary condition for the reference's gapped criteria (MINORg.py:2018-2044 with ot_gap >= 2, or a pattern
// whose bound the host derived): an alignment with mismatches + gap characters + unaligned positions <= K extends
// to a full-length alignment of the gRNA ending at the same diagonal with unit-cost edit distance <= K.
//
// Formulation.  Myers/Hyyro's bit-vector recurrences keep the DP column of ONE pattern in a machine word and
// resolve the vertical dependency with an addition.  Here the word holds the SAME DP cell of 32 different
// gRNAs (the query groups of the scan kernel), the rows of the column live in 2L registers (vertical deltas
// +1 / -1 as two bit-planes) and the vertical dependency is simply the order of the unrolled row loop.  Per
// cell (Hyyro 2001):   D0 = Eq | Mv | Mh_in          (cell equals its diagonal neighbour)
//                      Ph = Mv | ~(D0 | Pv)   Mh = Pv & D0          (horizontal delta, handed to the next row)
//                      Pv' = Mh_in | ~(D0 | Ph_in)   Mv' = Ph_in & D0   (vertical delta, kept for the next character)
// = 5 LOP3 + one shared-memory load (the mismatch word T[g][5i + base], shared with nothing else) per 32 cells of
// the DP, i.e. ~5.8 lane-instructions per (gRNA, text character) at L = 20 including the score counter, against
// ~27 for the one-pattern-per-thread Myers kernel it replaces.  The bottom-row score is a bit-sliced up/down
// counter biased so that "score <= K" is the complement of its top bit.
// A thread owns a run of consecutive text characters (plus L+K warm-up characters: a match of cost <= K spans at
// most L+K characters) and walks it once per query group; survivors are queued for the exact verifier.
#include "common.cuh"
#include "verify.cuh"

namespace mg {

constexpr int kRowsThreads = 512;

__device__ __forceinline__ uint32_t lop3_d0(uint32_t mis, uint32_t mv, uint32_t mh) {       // ~mis | mv | mh
    uint32_t r; asm("lop3.b32 %0, %1, %2, %3, 0xFD;" : "=r"(r) : "r"(mv), "r"(mh), "r"(mis)); return r;
}
__device__ __forceinline__ uint32_t lop3_a_or_nor(uint32_t a, uint32_t b, uint32_t c) {      // a | ~(b | c)
    uint32_t r; asm("lop3.b32 %0, %1, %2, %3, 0xF1;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
}

// Strand-aware sequential reader of base codes (0..3, 4 = invalid): position pos of the strand's text.  Keeps
// the current 16-base word (and its invalid bits) in registers, so a character costs a few ALU ops, not two loads.
struct TextReader {
    const GenomeView& gv;
    int strand;
    int64_t blk;
    uint32_t word, inv16;
    __device__ TextReader(const GenomeView& g, int s) : gv(g), strand(s), blk(-2), word(0u), inv16(0xFFFFu) {}
    __device__ __forceinline__ int at(int64_t pos) {
        const int64_t f = strand > 0 ? pos : gv.glen - 1 - pos;           // forward-genome coordinate
        if (f < 0 || f >= gv.glen) return kCodeInvalid;
        const int64_t b = f >> 4;
        if (b != blk) {
            blk = b;
            word = __ldg(gv.bases + b);
            inv16 = (__ldg(gv.invalid + (b >> 1)) >> ((unsigned)(b & 1) * 16u)) & 0xFFFFu;
        }
        const unsigned k = (unsigned)(f & 15);
        if ((inv16 >> k) & 1u) return kCodeInvalid;
        const int c = (int)((word >> (2u * k)) & 3u);
        return strand > 0 ? c : 3 - c;
    }
};

template <int L>
__global__ void __launch_bounds__(kRowsThreads, 1)
gapped_rows_kernel(GenomeView gv, QueriesView qv, const uint32_t* __restrict__ tab5, const GeneralCtx* __restrict__ ctx,
                   int K, int strand, int groups_per_chunk, int n_groups_used, int64_t eb, int64_t ee, int run) {
    constexpr int S5 = 5 * L;                          // words per group: [position][A, C, G, T, invalid]
    constexpr int NB = 6;                              // score counter bits (scores 0..32 + bias)
    extern __shared__ uint32_t tab[];
    const int N = qv.n;
    const int warm = L + K;
    // anchors e in [eb, ee) <-> text positions pos = e - 1 in [eb - 1, ee - 1)
    const int64_t p_first = eb - 1 + ((int64_t)blockIdx.x * kRowsThreads + threadIdx.x) * run;
    const int64_t p_last = min(ee - 1, p_first + run);
    const bool active = p_first < p_last;
    TextReader text(gv, strand);
    // score counter bias: c = score + 2^(NB-1) - K - 1, so score <= K  <=>  top bit of c is clear
    const uint32_t init = (uint32_t)(L + (1 << (NB - 1)) - K - 1);

    for (int cg = 0; cg < n_groups_used; cg += groups_per_chunk) {
        const int ng = min(groups_per_chunk, n_groups_used - cg);
        __syncthreads();
        for (int i = threadIdx.x; i < ng * S5; i += kRowsThreads) tab[i] = __ldg(tab5 + (size_t)cg * S5 + i);
        __syncthreads();
        if (!active) continue;
        for (int g = 0; g < ng; ++g) {
            const uint32_t* t = tab + (size_t)g * S5;
            const int q0 = (cg + g) * 32;
            // only the gRNA-orientation queries (index < N) take part: the strand is selected by the text
            const uint32_t qmask = (q0 + 32 <= N) ? 0xFFFFFFFFu : ((q0 < N) ? ((1u << (N - q0)) - 1u) : 0u);
            uint32_t Pv[L], Mv[L], sc[NB];
#pragma unroll
            for (int i = 0; i < L; ++i) { Pv[i] = 0xFFFFFFFFu; Mv[i] = 0u; }
#pragma unroll
            for (int b = 0; b < NB; ++b) sc[b] = ((init >> b) & 1u) ? 0xFFFFFFFFu : 0u;
            for (int64_t pos = p_first - warm; pos < p_last; ++pos) {
                const int c = text.at(pos);
                const uint32_t* row = t + c;
                uint32_t Ph = 0u, Mh = 0u;                 // horizontal delta entering row 0: D[0][j] = 0 for all j
#pragma unroll
                for (int i = 0; i < L; ++i) {
                    const uint32_t mis = row[5 * i];
                    const uint32_t D0 = lop3_d0(mis, Mv[i], Mh);
                    const uint32_t Ph_out = lop3_a_or_nor(Mv[i], D0, Pv[i]);
                    const uint32_t Mh_out = Pv[i] & D0;
                    Pv[i] = lop3_a_or_nor(Mh, D0, Ph);
                    Mv[i] = Ph & D0;
                    Ph = Ph_out;
                    Mh = Mh_out;
                }
                // score += Ph - Mh (bottom row), bit-sliced ripple
                uint32_t carry = Ph | Mh;
#pragma unroll
                for (int b = 0; b < NB; ++b) {
                    const uint32_t old = sc[b];
                    sc[b] = old ^ carry;
                    carry &= old ^ Mh;                     // increment ripples through 1s, decrement through 0s
                }
                const uint32_t hit = ~sc[NB - 1] & qmask;
                if (hit != 0u && pos >= p_first) {
                    uint32_t h = hit;
                    while (h) {
                        const int i = __ffs(h) - 1;
                        h &= h - 1;
                        emit_candidate(*ctx, strand > 0 ? q0 + i : N + q0 + i, pos + 1);
                    }
                }
            }
        }
    }
}

template <int L>
static int launch_rows(const mg_genome* g, const mg_queries* q, const uint32_t* d_tab5, const GeneralCtx* d_ctx, int K,
                       int strand, int64_t eb, int64_t ee, cudaStream_t s) {
    int dev = 0, sms = 0, smem_max = 0;
    MG_CUDA(cudaGetDevice(&dev));
    MG_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    MG_CUDA(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    const int n_groups_used = (q->n + 31) / 32;
    const int S5 = 5 * L;
    const int max_groups = (smem_max - 1024) / (S5 * 4);
    const int gpc = n_groups_used < max_groups ? n_groups_used : max_groups;
    const size_t smem = (size_t)gpc * S5 * 4;
    auto kern = gapped_rows_kernel<L>;
    MG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t n_anchor = ee - eb;
    // characters per thread: long enough to amortise the L+K warm-up, short enough to fill the machine
    int64_t run = (n_anchor + (int64_t)sms * kRowsThreads - 1) / ((int64_t)sms * kRowsThreads);
    if (run < 64) run = 64;
    if (run > 4096) run = 4096;
    const int64_t threads = (n_anchor + run - 1) / run;
    const unsigned grid = (unsigned)((threads + kRowsThreads - 1) / kRowsThreads);
    kern<<<grid, kRowsThreads, smem, s>>>(g->view(), q->view(), d_tab5, d_ctx, K, strand, gpc, n_groups_used, eb, ee, (int)run);
    MG_CUDA(cudaGetLastError());
    return MG_OK;
}

// tab5[g][5i + b]: like the scan tables but with an explicit all-ones column for invalid bases.
__global__ void rows_table_kernel(const uint8_t* __restrict__ codes, int n, int len, int n_groups, uint32_t* __restrict__ tab5) {
    const int stride = 5 * len;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_groups * stride) return;
    const int g = idx / stride, slot = idx % stride;
    const int i = slot / 5, b = slot % 5;
    uint32_t w = 0xFFFFFFFFu;
    if (b < 4) {
        w = 0;
        for (int k = 0; k < 32; ++k) {
            const int qi = g * 32 + k;
            const bool mism = (qi >= n) || (codes[(size_t)qi * len + i] != b);
            w |= (mism ? 1u : 0u) << k;
        }
    }
    tab5[idx] = w;
}

int launch_gapped_rows(const mg_genome* g, const mg_queries* q, const GeneralCtx* d_ctx, int K, int strand, int64_t eb,
                       int64_t ee, uint32_t** d_tab5_cache, cudaStream_t s) {
    if (q->n == 0 || ee <= eb) return MG_OK;
    const int n_groups_used = (q->n + 31) / 32;
    if (*d_tab5_cache == nullptr) {
        const int words = n_groups_used * 5 * q->len;
        MG_CUDA(dev_alloc((void**)d_tab5_cache, (size_t)words * 4 + 16, s));
        rows_table_kernel<<<(words + 255) / 256, 256, 0, s>>>(q->d_codes, q->n, q->len, n_groups_used, *d_tab5_cache);
        MG_CUDA(cudaGetLastError());
    }
    switch (q->len) {
#define MG_ROWS(LL) case LL: return launch_rows<LL>(g, q, *d_tab5_cache, d_ctx, K, strand, eb, ee, s);
        MG_ROWS(1) MG_ROWS(2) MG_ROWS(3) MG_ROWS(4) MG_ROWS(5) MG_ROWS(6) MG_ROWS(7) MG_ROWS(8) MG_ROWS(9) MG_ROWS(10)
        MG_ROWS(11) MG_ROWS(12) MG_ROWS(13) MG_ROWS(14) MG_ROWS(15) MG_ROWS(16) MG_ROWS(17) MG_ROWS(18) MG_ROWS(19)
        MG_ROWS(20) MG_ROWS(21) MG_ROWS(22) MG_ROWS(23) MG_ROWS(24) MG_ROWS(25) MG_ROWS(26) MG_ROWS(27) MG_ROWS(28)
        MG_ROWS(29) MG_ROWS(30) MG_ROWS(31) MG_ROWS(32)
#undef MG_ROWS
        default: set_error("gRNA length %d unsupported (1..32)", q->len); return MG_ERR_UNSUPPORTED;
    }
}

}  // namespace mg
