// K4: bit-sliced ungapped mismatch scan (the hot kernel).
//
// Replaces, for the reference's default criteria (no --ot-pattern, ot_gap <= 1, ot_pamless), the whole of
// MINORg._offtarget_hits: `blastn -task blastn-short` + is_offtarget_hit per HSP (MINORg.py:2123-2144,
// criterion MINORg.py:2018-2044, position check MINORg.py:1994-2000).
//
// Formulation.  Queries (gRNAs in both orientations) are grouped 32 to a 32-bit word, bit i = query i of
// the group.  For gRNA position j and genome base b the table word T[g][4j+b] has bit i set iff query i
// does NOT have base b at position j.  A thread owns ONE genome window p; the mismatch word of the group
// against that window at position j is T[g][4j + base(p+j)] - one conflict-free shared-memory load (the
// 32 lanes of a warp hit at most 4 consecutive banks + broadcast).  The L words are summed per bit with a
// carry-save adder tree of LOP3 full adders (csa_gen.cuh) and compared with the mismatch budget, giving
// 32 verdicts per ~(L loads + 1.9 L logic ops): ~2 lane-instructions per (gRNA, window, strand) cell,
// all on the full-rate ALU/LSU pipes, against ~10 for a per-cell XOR/POPC.  Hits are rare (4e-7 per cell
// at <=4 mismatches on random sequence), so the mask lookup and the per-gRNA counting live in a cold path.
//
// The kernel is bound by the LSU (L x 4 B per lane per 32 cells from shared memory at 128 B/clk/SM) and by
// the ALU pipe (~39 LOP3 per 32 cells at L=20) at about the same level; HBM traffic is the packed genome
// once per query chunk (0.375 B/base) and is three orders of magnitude below the HBM roofline.
#include "common.cuh"
#include "csa_gen.cuh"
#include "verify.cuh"

namespace mg {

template <int NB>
__device__ __forceinline__ uint32_t le_const_impl(const uint32_t (&b)[NB], int K) {
    // bit i of the result = (count_i <= K); with K a compile-time constant this folds to ~NB/2 LOP3.
    uint32_t lt = 0u, eq = 0xFFFFFFFFu;
#pragma unroll
    for (int k = NB - 1; k >= 0; --k) {
        if ((K >> k) & 1) { lt |= eq & ~b[k]; eq &= b[k]; }
        else              { eq &= ~b[k]; }
    }
    return lt | eq;
}

template <int NB>
__device__ __forceinline__ uint32_t le_runtime(const uint32_t (&b)[NB], int K) {
    uint32_t lt = 0u, eq = 0xFFFFFFFFu;
#pragma unroll
    for (int k = NB - 1; k >= 0; --k) {
        const uint32_t kb = 0u - (uint32_t)((K >> k) & 1);      // warp-uniform all-ones / zero
        lt |= eq & ~b[k] & kb;
        eq &= ~(b[k] ^ kb);
    }
    return lt | eq;
}

// Cold path: some query of group `group` is within budget at window p.
__device__ __noinline__ void report_hits(const GenomeView& gv, int n_queries, int L, uint32_t hit, int group,
                                         int64_t p, uint32_t* __restrict__ counts) {
    const int c = contig_of(gv, p);
    if (c < 0) return;
    const int64_t a = max(p, gv.cstart[c]), b = min(p + L, gv.cend[c]);   // on-contig span of the window
    if (a >= b) return;
    if (gv.n_masks && span_masked(gv, a, b)) return;                      // on-target (MINORg.py:1994-2000)
    while (hit) {
        const int i = __ffs(hit) - 1;
        hit &= hit - 1;
        const int q = group * 32 + i;
        if (q < n_queries) atomicAdd(counts + q, 1u);
    }
}

// Cold path of the general screen: the Hamming budget is only a prefilter; each survivor is queued as an anchor
// for the exact verifier pass (verify.cuh, screen_general.cu).  Query q < N is gRNA q on the '+' strand, q >= N its reverse complement, i.e.
// gRNA q-N on the '-' strand, whose window [p, p+L) ends at glen - p in reverse-complement coordinates.
__device__ __noinline__ void verify_hits(const GenomeView& gv, const QueriesView& qv, const GeneralCtx* ctx, uint32_t hit,
                                         int group, int64_t p) {
    while (hit) {
        const int i = __ffs(hit) - 1;
        hit &= hit - 1;
        const int q = group * 32 + i;
        if (q >= 2 * qv.n) continue;
        emit_candidate(*ctx, q, q < qv.n ? p + qv.len : gv.glen - p);
    }
}

template <int L, int K, int THREADS, bool GENERAL, bool SPLIT>
__global__ void __launch_bounds__(THREADS, 1)
screen_bitsliced_kernel(GenomeView gv, QueriesView qv, int kthr, int groups_per_chunk, int64_t wb, int64_t we,
                        uint32_t* __restrict__ counts, const GeneralCtx* __restrict__ gctx,
                        unsigned long long* __restrict__ stats) {
    constexpr int S = 4 * L + 1;                     // table words per group
    constexpr int NB = mgcsa::CountBits<L>::value;
    constexpr int T1 = SPLIT ? mgcsa::SplitOf<L, K>::value : 0;   // positions summed before the early-out test
    constexpr int NB1 = mgcsa::CountBits<(T1 > 0 ? T1 : 1)>::value;
    extern __shared__ uint32_t tab[];

    // contiguous, warp-aligned window range of this CTA
    const int64_t nwin = we - wb;
    int64_t per = (nwin + gridDim.x - 1) / gridDim.x;
    per = (per + 31) / 32 * 32;
    const int64_t w0 = wb + (int64_t)blockIdx.x * per;
    const int64_t w1 = min(we, w0 + per);
    unsigned int n_steps = 0, n_stage2 = 0;       // per-warp tallies of (warp, group) steps / second stages

    for (int cg = 0; cg < qv.n_groups; cg += groups_per_chunk) {
        const int ng = min(groups_per_chunk, qv.n_groups - cg);
        __syncthreads();
        {   // stage this chunk of the query tables: coalesced 32-bit copies (the table is L2-resident)
            const uint32_t* src = qv.tab + (size_t)cg * S;
            for (int i = threadIdx.x; i < ng * S; i += THREADS) tab[i] = __ldg(src + i);
        }
        __syncthreads();

        for (int64_t pbase = w0; pbase < w1; pbase += THREADS) {
            const int64_t p = pbase + threadIdx.x;
            const bool active = p < w1;
            uint32_t off[L];
            {
                uint32_t lo = 0, hi = 0, inv = 0xFFFFFFFFu;
                if (active) load_window(gv, p, lo, hi, inv);
#pragma unroll
                for (int j = 0; j < L; ++j) {
                    const uint32_t code = (j < 16) ? ((lo >> (2 * j)) & 3u) : ((hi >> (2 * (j - 16))) & 3u);
                    off[j] = (((inv >> j) & 1u) ? (uint32_t)(4 * L) : (uint32_t)(4 * j) + code) * 4u;   // byte offset
                }
            }
            const char* t = reinterpret_cast<const char*>(tab);
#pragma unroll 1
            for (int g = 0; g < ng; ++g, t += S * 4) {
                uint32_t hit;
                if constexpr (T1 > 0) {
                    // Stage 1: the first T1 positions.  The count can only grow, so a query already over budget
                    // is decided; on random sequence a whole warp (32 windows x 32 queries) is over budget ~90 %
                    // of the time at L=20, K=4, and the remaining L-T1 loads and their adders are skipped.
                    // Every cell is still decided exactly - no verdict is skipped, only arithmetic whose result
                    // cannot change it.
                    uint32_t x1[T1];
#pragma unroll
                    for (int j = 0; j < T1; ++j) x1[j] = *reinterpret_cast<const uint32_t*>(t + off[j]);
                    uint32_t part[NB1];
                    mgcsa::csa_count<T1>(x1, part);             // the exact partial count is reused by the second stage
                    const uint32_t alive = (K >= 0) ? le_const_impl<NB1>(part, K) : le_runtime<NB1>(part, kthr);
                    ++n_steps;
                    if (!__any_sync(0xFFFFFFFFu, alive != 0u)) continue;
                    ++n_stage2;
                    uint32_t x2[L - T1];
#pragma unroll
                    for (int j = T1; j < L; ++j) x2[j - T1] = *reinterpret_cast<const uint32_t*>(t + off[j]);
                    uint32_t bits[NB];
                    mgcsa::csa_count_rest<L, T1>(part, x2, bits);
                    hit = (K >= 0) ? le_const_impl<NB>(bits, K) : le_runtime<NB>(bits, kthr);
                } else {
                    uint32_t x[L];
#pragma unroll
                    for (int j = 0; j < L; ++j) x[j] = *reinterpret_cast<const uint32_t*>(t + off[j]);
                    uint32_t bits[NB];
                    mgcsa::csa_count<L>(x, bits);
                    hit = (K >= 0) ? le_const_impl<NB>(bits, K) : le_runtime<NB>(bits, kthr);
                }
                if (hit) {
                    if (GENERAL) verify_hits(gv, qv, gctx, hit, cg + g, p);
                    else report_hits(gv, 2 * qv.n, L, hit, cg + g, p, counts);
                }
            }
        }
    }
    if (stats != nullptr && (threadIdx.x & 31) == 0) {
        atomicAdd(stats, (unsigned long long)n_steps);
        atomicAdd(stats + 1, (unsigned long long)n_stage2);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Pair-word variant of the scan (the default): stage 1 looks up, per PAIR of gRNA positions, the word "query i
// mismatches at least one of the two positions" - one 32-bit load per pair (16 table entries = 16 banks, still
// conflict-free) instead of one per position.  The number of set pair words never exceeds the number of
// mismatches, so a cell whose pair count is over budget is decided; on uniform sequence a pair word is set with
// probability 15/16 and, at L = 20, K = 4, a whole (warp x 32 queries) step survives all 10 pairs only ~1 % of the
// time.  Survivors take the exact path: L per-position words (from the L2-resident table in global memory) and
// the full adder tree.  10.2 loads per step instead of 15.5 (single-word early-out) or 20 (plain).
template <int L, int K, int THREADS, bool GENERAL>
__global__ void __launch_bounds__(THREADS, 1)
screen_pairs_kernel(GenomeView gv, QueriesView qv, int kthr, int groups_per_chunk, int64_t wb, int64_t we,
                    uint32_t* __restrict__ counts, const GeneralCtx* __restrict__ gctx,
                    unsigned long long* __restrict__ stats) {
    constexpr int NP = (L + 1) / 2;                  // position pairs (an odd last position pairs with "always matches")
    constexpr int T = mgcsa::PairSplit<L, K>::value; // pairs tested in stage 1
    constexpr int S2 = 16 * NP + 1;                  // pair-table words per group in global memory
    constexpr int SS = 16 * T + 1;                   // ... and in shared memory (only the pairs stage 1 uses)
    constexpr int S1 = 4 * L + 1;
    constexpr int NBT = mgcsa::CountBits<T>::value;
    constexpr int NB = mgcsa::CountBits<L>::value;
    extern __shared__ uint32_t tab[];

    const int64_t nwin = we - wb;
    int64_t per = (nwin + gridDim.x - 1) / gridDim.x;
    per = (per + 31) / 32 * 32;
    const int64_t w0 = wb + (int64_t)blockIdx.x * per;
    const int64_t w1 = min(we, w0 + per);
    unsigned int n_steps = 0, n_stage2 = 0;

    for (int cg = 0; cg < qv.n_groups; cg += groups_per_chunk) {
        const int ng = min(groups_per_chunk, qv.n_groups - cg);
        __syncthreads();
        for (int i = threadIdx.x; i < ng * SS; i += THREADS) {
            const int g = i / SS, slot = i % SS;
            tab[i] = __ldg(qv.tab2 + (size_t)(cg + g) * S2 + (slot < 16 * T ? slot : 16 * NP));
        }
        __syncthreads();

        for (int64_t pbase = w0; pbase < w1; pbase += THREADS) {
            const int64_t p = pbase + threadIdx.x;
            const bool active = p < w1;
            uint32_t lo = 0, hi = 0, inv = 0xFFFFFFFFu;
            if (active) load_window(gv, p, lo, hi, inv);
            uint32_t off[T];                          // byte offsets of this window's pair words
#pragma unroll
            for (int k = 0; k < T; ++k) {
                const uint32_t nib = (k < 8) ? ((lo >> (4 * k)) & 15u) : ((hi >> (4 * (k - 8))) & 15u);
                const uint32_t bad = (inv >> (2 * k)) & ((2 * k + 1 < L) ? 3u : 1u);
                off[k] = (bad ? (uint32_t)(16 * T) : (uint32_t)(16 * k) + nib) * 4u;
            }
            const char* t = reinterpret_cast<const char*>(tab);
#pragma unroll 1
            for (int g = 0; g < ng; ++g, t += SS * 4) {
                uint32_t x[T];
#pragma unroll
                for (int k = 0; k < T; ++k) x[k] = *reinterpret_cast<const uint32_t*>(t + off[k]);
                uint32_t alive;
                if constexpr (mgcsa::HasCountLe<T, K>::value) {
                    alive = mgcsa::count_le<T, K>(x);           // generated threshold network (gen_csa.py): 15 LOP3 for 4-of-10
                } else {
                    uint32_t part[NBT];
                    mgcsa::csa_count<T>(x, part);
                    alive = (K >= 0) ? le_const_impl<NBT>(part, K) : le_runtime<NBT>(part, kthr);
                }
                ++n_steps;
                if (!__any_sync(0xFFFFFFFFu, alive != 0u)) continue;
                ++n_stage2;
                // exact count for the whole warp (warp-uniform branch): per-position words from global memory
                const uint32_t* t1 = qv.tab + (size_t)(cg + g) * S1;
                uint32_t y[L];
#pragma unroll
                for (int j = 0; j < L; ++j) {
                    const uint32_t code = (j < 16) ? ((lo >> (2 * j)) & 3u) : ((hi >> (2 * (j - 16))) & 3u);
                    y[j] = __ldg(t1 + (((inv >> j) & 1u) ? (uint32_t)(4 * L) : (uint32_t)(4 * j) + code));
                }
                uint32_t bits[NB];
                mgcsa::csa_count<L>(y, bits);
                const uint32_t hit = ((K >= 0) ? le_const_impl<NB>(bits, K) : le_runtime<NB>(bits, kthr)) & alive;
                if (hit) {
                    if (GENERAL) verify_hits(gv, qv, gctx, hit, cg + g, p);
                    else report_hits(gv, 2 * qv.n, L, hit, cg + g, p, counts);
                }
            }
        }
    }
    if (stats != nullptr && (threadIdx.x & 31) == 0) {
        atomicAdd(stats, (unsigned long long)n_steps);
        atomicAdd(stats + 1, (unsigned long long)n_stage2);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Comparison kernel (MG_SCAN=popc): the per-cell XOR/POPC formulation BASELINE.json's north star starts from - one
// thread per window, every query compared with  x = w ^ q;  d = (x | x >> 1) & 0x5555...;  mm = popc(d).  Same
// results as the bit-sliced scan (tests run both); kept to put a measured number next to the design choice:
// ~12 integer instructions and two quarter-rate POPCs per cell against ~0.8 lane-instructions for the pair scan.
__global__ void __launch_bounds__(1024, 1)
screen_popc_kernel(GenomeView gv, QueriesView qv, const uint2* __restrict__ qbits, const uint2* __restrict__ qinv, int K,
                   int queries_per_chunk, int64_t wb, int64_t we, uint32_t* __restrict__ counts) {
    extern __shared__ uint2 sq[];                 // [chunk] packed query, then [chunk] its invalid-position mask
    const int L = qv.len, nq = 2 * qv.n;
    const unsigned long long even = 0x5555555555555555ull;
    const unsigned long long lmask = L >= 32 ? even : (((1ull << (2 * L)) - 1ull) & even);
    for (int c0 = 0; c0 < nq; c0 += queries_per_chunk) {
        const int m = min(queries_per_chunk, nq - c0);
        __syncthreads();
        for (int i = threadIdx.x; i < m; i += blockDim.x) { sq[i] = qbits[c0 + i]; sq[queries_per_chunk + i] = qinv[c0 + i]; }
        __syncthreads();
        for (int64_t p = wb + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < we; p += (int64_t)gridDim.x * blockDim.x) {
            uint32_t lo, hi, inv;
            load_window(gv, p, lo, hi, inv);
            const unsigned long long w = ((unsigned long long)hi << 32) | lo;
            unsigned long long winv = 0;          // invalid window positions spread to the even bits
            for (int j = 0; j < L; ++j) winv |= (unsigned long long)((inv >> j) & 1u) << (2 * j);
            for (int i = 0; i < m; ++i) {
                const uint2 qb = sq[i], qi = sq[queries_per_chunk + i];
                const unsigned long long q = ((unsigned long long)qb.y << 32) | qb.x;
                const unsigned long long x = w ^ q;
                const unsigned long long d = ((x | (x >> 1)) & lmask) | winv | (((unsigned long long)qi.y << 32) | qi.x);
                if (__popcll(d) <= K) report_hits(gv, nq, L, 1u << ((c0 + i) & 31), (c0 + i) >> 5, p, counts);
            }
        }
    }
}

// packs query codes into 2-bit words + invalid masks for the comparison kernel
__global__ void popc_pack_kernel(const uint8_t* __restrict__ codes, int nq, int len, uint2* __restrict__ qbits, uint2* __restrict__ qinv) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    unsigned long long b = 0, iv = 0;
    for (int j = 0; j < len; ++j) {
        const int c = codes[(size_t)q * len + j];
        if (c >= 4) iv |= 1ull << (2 * j); else b |= (unsigned long long)c << (2 * j);
    }
    qbits[q] = make_uint2((uint32_t)b, (uint32_t)(b >> 32));
    qinv[q] = make_uint2((uint32_t)iv, (uint32_t)(iv >> 32));
}

static int launch_popc(const mg_genome* g, const mg_queries* q, int k, int64_t wb, int64_t we, uint32_t* d_counts, cudaStream_t s) {
    const int nq = 2 * q->n;
    uint2 *d_b = nullptr, *d_i = nullptr;
    MG_CUDA(dev_alloc((void**)&d_b, sizeof(uint2) * (size_t)nq, s));
    MG_CUDA(dev_alloc((void**)&d_i, sizeof(uint2) * (size_t)nq, s));
    popc_pack_kernel<<<(nq + 255) / 256, 256, 0, s>>>(q->d_codes, nq, q->len, d_b, d_i);
    int dev = 0, sms = 0;
    MG_CUDA(cudaGetDevice(&dev));
    MG_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int per_chunk = nq < 12288 ? nq : 12288;                         // 16 B per query in shared memory
    const size_t smem = (size_t)per_chunk * 16;
    MG_CUDA(cudaFuncSetAttribute(screen_popc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t grid = (we - wb + 1023) / 1024;
    if (grid > sms) grid = sms;
    screen_popc_kernel<<<(unsigned)grid, 1024, smem, s>>>(g->view(), q->view(), d_b, d_i, k, per_chunk, wb, we, d_counts);
    MG_CUDA(cudaGetLastError());
    dev_free(d_b, s);
    dev_free(d_i, s);
    return MG_OK;
}

// [0] (warp, query-group) steps, [1] steps that needed the second stage - of the most recent launch
// One buffer per device (a process may screen on several devices; the counters must live on the launching one).
constexpr int kMaxDevices = 64;
static unsigned long long* g_stats_dev[kMaxDevices] = {nullptr};

static int stats_buffer(unsigned long long** out) {
    int dev = 0;
    MG_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= kMaxDevices) { *out = nullptr; return MG_OK; }     // no statistics beyond 64 devices
    if (!g_stats_dev[dev]) MG_CUDA(cudaMalloc(&g_stats_dev[dev], 16));
    *out = g_stats_dev[dev];
    return MG_OK;
}

int screen_stats(unsigned long long out[2]) {
    out[0] = out[1] = 0;
    int dev = 0;
    MG_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= kMaxDevices || !g_stats_dev[dev]) return MG_OK;
    MG_CUDA(cudaMemcpy(out, g_stats_dev[dev], 16, cudaMemcpyDeviceToHost));
    return MG_OK;
}

template <int L, int K, bool GENERAL = false>
static int launch_one(const mg_genome* g, const mg_queries* q, int kthr, int64_t wb, int64_t we, uint32_t* d_counts,
                      cudaStream_t s, const GeneralCtx* gctx = nullptr) {
    constexpr int THREADS = 1024;
    const char* scan_env = getenv("MG_SCAN");            // "singles": the per-position early-out scan (A/B)
    const bool pairs = L >= 8 && !(scan_env && strcmp(scan_env, "singles") == 0);
    const int S = pairs ? 16 * mgcsa::PairSplit<L, K>::value + 1 : 4 * L + 1;
    int dev = 0, sms = 0, smem_max = 0;
    MG_CUDA(cudaGetDevice(&dev));
    MG_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    MG_CUDA(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    const int max_groups = (smem_max - 1024) / (S * 4);
    const int gpc = q->n_groups < max_groups ? q->n_groups : max_groups;
    const size_t smem = (size_t)gpc * S * 4;
    auto kern = pairs ? screen_pairs_kernel<L, K, THREADS, GENERAL> : screen_bitsliced_kernel<L, K, THREADS, GENERAL, true>;
    MG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t nwin = we - wb;
    int64_t grid = (nwin + THREADS - 1) / THREADS;
    if (grid > sms) grid = sms;             // persistent: one CTA per SM, each owns a contiguous genome slice
    if (grid < 1) grid = 1;
    unsigned long long* stats = nullptr;
    { const int rc = stats_buffer(&stats); if (rc != MG_OK) return rc; }
    if (stats) MG_CUDA(cudaMemsetAsync(stats, 0, 16, s));
    kern<<<(unsigned)grid, THREADS, smem, s>>>(g->view(), q->view(), kthr, gpc, wb, we, d_counts, gctx, stats);
    MG_CUDA(cudaGetLastError());
    return MG_OK;
}

template <int L>
static int dispatch_k(const mg_genome* g, const mg_queries* q, int k, int64_t wb, int64_t we, uint32_t* c, cudaStream_t s) {
    switch (k) {
        case 0: return launch_one<L, 0>(g, q, k, wb, we, c, s);
        case 1: return launch_one<L, 1>(g, q, k, wb, we, c, s);
        case 2: return launch_one<L, 2>(g, q, k, wb, we, c, s);
        case 3: return launch_one<L, 3>(g, q, k, wb, we, c, s);
        case 4: return launch_one<L, 4>(g, q, k, wb, we, c, s);
        case 5: return launch_one<L, 5>(g, q, k, wb, we, c, s);
        default: return launch_one<L, -1>(g, q, k, wb, we, c, s);
    }
}

int launch_screen_bitsliced(const mg_genome* g, const mg_queries* q, int max_mismatch, int64_t wb, int64_t we,
                            uint32_t* d_counts, cudaStream_t s) {
    if (q->n == 0 || we <= wb) return MG_OK;
    const int k = max_mismatch;
    {
        const char* scan_env = getenv("MG_SCAN");
        if (scan_env && strcmp(scan_env, "popc") == 0) return launch_popc(g, q, k, wb, we, d_counts, s);
    }
    switch (q->len) {
#define MG_CASE_FULL(LL) case LL: return dispatch_k<LL>(g, q, k, wb, we, d_counts, s);
#define MG_CASE_RT(LL) case LL: return launch_one<LL, -1>(g, q, k, wb, we, d_counts, s);
        MG_CASE_FULL(20) MG_CASE_FULL(23)
        MG_CASE_RT(1) MG_CASE_RT(2) MG_CASE_RT(3) MG_CASE_RT(4) MG_CASE_RT(5) MG_CASE_RT(6) MG_CASE_RT(7) MG_CASE_RT(8)
        MG_CASE_RT(9) MG_CASE_RT(10) MG_CASE_RT(11) MG_CASE_RT(12) MG_CASE_RT(13) MG_CASE_RT(14) MG_CASE_RT(15)
        MG_CASE_RT(16) MG_CASE_RT(17) MG_CASE_RT(18) MG_CASE_RT(19) MG_CASE_RT(21) MG_CASE_RT(22) MG_CASE_RT(24)
        MG_CASE_RT(25) MG_CASE_RT(26) MG_CASE_RT(27) MG_CASE_RT(28) MG_CASE_RT(29) MG_CASE_RT(30) MG_CASE_RT(31)
        MG_CASE_RT(32)
#undef MG_CASE_FULL
#undef MG_CASE_RT
        default: set_error("gRNA length %d unsupported (1..32)", q->len); return MG_ERR_UNSUPPORTED;
    }
}


// Same scan as a prefilter for the general screen (runtime budget, cold path = exact verifier).
int launch_screen_bitsliced_general(const mg_genome* g, const mg_queries* q, int k, const GeneralCtx* d_ctx,
                                    int64_t wb, int64_t we, uint32_t* d_counts, cudaStream_t s) {
    if (q->n == 0 || we <= wb) return MG_OK;
    switch (q->len) {
#define MG_CASE_GEN(LL) case LL: return launch_one<LL, -1, true>(g, q, k, wb, we, d_counts, s, d_ctx);
#define MG_CASE_GEN_K(LL) case LL: switch (k) {                                                   \
            case 0: return launch_one<LL, 0, true>(g, q, k, wb, we, d_counts, s, d_ctx);            \
            case 1: return launch_one<LL, 1, true>(g, q, k, wb, we, d_counts, s, d_ctx);            \
            case 2: return launch_one<LL, 2, true>(g, q, k, wb, we, d_counts, s, d_ctx);            \
            case 3: return launch_one<LL, 3, true>(g, q, k, wb, we, d_counts, s, d_ctx);            \
            case 4: return launch_one<LL, 4, true>(g, q, k, wb, we, d_counts, s, d_ctx);            \
            case 5: return launch_one<LL, 5, true>(g, q, k, wb, we, d_counts, s, d_ctx);            \
            default: return launch_one<LL, -1, true>(g, q, k, wb, we, d_counts, s, d_ctx); }
        MG_CASE_GEN(1) MG_CASE_GEN(2) MG_CASE_GEN(3) MG_CASE_GEN(4) MG_CASE_GEN(5) MG_CASE_GEN(6) MG_CASE_GEN(7) MG_CASE_GEN(8)
        MG_CASE_GEN(9) MG_CASE_GEN(10) MG_CASE_GEN(11) MG_CASE_GEN(12) MG_CASE_GEN(13) MG_CASE_GEN(14) MG_CASE_GEN(15)
        MG_CASE_GEN(16) MG_CASE_GEN(17) MG_CASE_GEN(18) MG_CASE_GEN(19) MG_CASE_GEN_K(20) MG_CASE_GEN(21) MG_CASE_GEN(22)
        MG_CASE_GEN_K(23) MG_CASE_GEN(24) MG_CASE_GEN(25) MG_CASE_GEN(26) MG_CASE_GEN(27) MG_CASE_GEN(28) MG_CASE_GEN(29)
        MG_CASE_GEN(30) MG_CASE_GEN(31) MG_CASE_GEN(32)
#undef MG_CASE_GEN
#undef MG_CASE_GEN_K
        default: set_error("gRNA length %d unsupported (1..32)", q->len); return MG_ERR_UNSUPPORTED;
    }
}

}  // namespace mg
