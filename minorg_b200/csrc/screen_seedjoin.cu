// K7: seed-and-join screen for the no-pattern criteria with AT MOST ONE GAP on LARGE gRNA sets (the north star's job: 10^6
// gRNA, <= 4 mismatches / <= 1 gap; also plain Hamming <= M, see "Without gaps" below).  Replaces, for those criteria, the
// bit-sliced edit-distance DP of screen_gapped.cu, which has to touch every (gRNA, text position, DP row): 182 issue slots per
// 32 cells whatever the data - and, for sets of >= 10^5 gRNA, the pair-word scan of screen_bitsliced.cu.
//
// Criterion (MINORg.py:2018-2044 with ot_gap = 2, no pattern, PAM-less): an anchor (gRNA, strand, virtual end e) is
// problematic iff some alignment ending there has  mm <= M, gaps <= 1, mm + gaps + unaligned <= M + 1.  For a window whose
// neighbourhood holds only valid bases (everything else goes to the DP path, see "dirty" below) that is equivalent to
//     ungapped : Hamming h <= M, or h == M + 1 with a mismatch at an end position (trim it: mm = M, unaligned = 1), or
//     deletion : min over k = 1..19 of  #mm(query[0,k) on diagonal w-1) + #mm(query[k,20) on diagonal w)      <= M, or
//     insertion: min over k = 1..18 of  #mm(query[0,k) on diagonal w+1) + #mm(query[k+1,20) on diagonal w)    <= M
// (with a gap, trimming a position costs what the mismatch it removes cost, so the full-length alignment decides).
//
// Pigeonhole.  Cut the gRNA into halves H1 = [0,10), H2 = [10,20), t = floor((M+1)/2), g = M - t - 1.  Every alignment that
// satisfies the criterion has a gap-free half with <= t mismatches, or its gapped half has <= g mismatches:
//     R1  H2 on the final diagonal w            with <= t mismatches      R2  H1 on w       with <= t
//     R3  H1 on w-1 (a deletion follows)        with <= t                 R4  H1 on w+1 (an insertion follows) with <= t
//     R5  H1 with a deletion before k = 1..9    and <= g mismatches       R6  H1 with an insertion at k = 1..9, <= g
//     R7  H2 with a deletion before k = 11..19  and <= g                  R8  H2 with an insertion at k = 10..18, <= g
// So instead of asking every gRNA about every position, every text position asks hash tables of the gRNA SET (by H1, by H2,
// direct-addressed by the 20-bit 10-mer; by H1 / H2 without one of their positions, 18-bit 9-mer) for the gRNA whose half
// equals one of the strings the text can produce: the Hamming-<=t neighbours of the 10-mer at p (R1-R4: 436 strings for
// t = 2), the 11-mer at p with one inner base removed and <= g substitutions (R5, R7: 279), the 9-mer at p with <= g
// substitutions against the tables that lack a position (R6, R8: 28 - the index, not the text, enumerates the inserted base):
// 743 directory probes per position and strand, independent of the number of gRNA; at 10^6 gRNA ~0.95 entries per 10-mer
// bucket and ~34 per 9-mer bucket, i.e. 3.15e-3 of the cells are looked at (counted by the kernel: mg_last_seedjoin_stats),
// each with ~30 instructions (two or three diagonal mismatch masks from bit planes held in registers, one route-specific
// necessary test), and 4.2 % of those with the exact evaluation above.
//
// Exactly-once counting.  An anchor can be reached through several routes and variants (from different threads).  It is
// counted by the route that its CANONICAL alignment (ungapped if it fits, else the deletion with the smallest k, else the
// insertion with the smallest k) implies: its gap-free half if that has <= t mismatches (H2 first), else its gapped half
// with that k.  That is a function of (gRNA, anchor) only, so every thread that reaches the anchor agrees on who counts it.
//
// Dirty windows.  A window whose bases [w-3, w+23) touch a 32-base block with an invalid base (N, IUPAC, contig padding,
// i.e. also everything that hangs off a contig end), or whose bases [w-3, w+24) meet a masked (on-target) interval - there
// "problematic" also depends on spans - is left to the exhaustive kernels (DP prefilter + exact verifier; pair-word scan without
// gaps): the host collects those blocks and the intervals, merges them into window ranges (launch_seedjoin), hands the ranges
// to the scan kernel, which skips exactly those windows, and to the caller, which screens them.  Every anchor is owned by
// exactly one side.
//
// Without gaps (crit->max_gap == 0, Hamming <= M <= 5): routes R1 / R2 only, t = floor(M/2), stage 1 is the exact test.
//
// tests/test_seedjoin_logic.py restates this scheme in Python and checks it against the exhaustive C oracle on the CPU.
#include <algorithm>
#include <mutex>
#include <type_traits>
#include <vector>

#include "common.cuh"
#include "verify.cuh"

namespace mg {

constexpr int kSjL = 20;
constexpr uint32_t kSjKeys = 1u << 20;               // 4^10 half keys
constexpr uint32_t kH1M = 0x3FFu, kH2M = 0xFFC00u, kFull = 0xFFFFFu;

// Tables: 0 = gRNA by H1, 1 = gRNA by H2 (4^10 keys, one entry per gRNA); 2 / 3 = gRNA by H1 / H2 WITHOUT one of its positions
// (4^9 keys, nine entries per gRNA: H1 without position k = 1..9, H2 without its position k = 0..8, i.e. query position 10 + k):
// an insertion in the gRNA is found by asking for the text's 9-mer instead of trying every inserted base at every position.
// Entry (8 B): [19:0] low-bit plane of the 20 query bases, [39:20] high-bit plane, [43:40] k (position inside the half),
// [63:44] gRNA index (< 2^20).
// Directory: ONE 8-byte word per key for a pair of tables (the same key is always looked up in both).  Tables 0 / 1: each
// half = first entry << 12 | entry count.  Tables 2 / 3 share one entry array in which a key's table-2 entries are followed
// by its table-3 entries: x = first entry, y = table-2 count | table-3 count << 16.  (A gRNA set with a fuller bucket -
// thousands of gRNA sharing a half - stays with the DP path.)
constexpr int kSjTables = 4;
constexpr uint32_t kSjKeys9 = 1u << 18;
constexpr int kSjCountBits01 = 12, kSjCountBits23 = 12;
__host__ __device__ constexpr uint32_t sj_table_keys(int t) { return t < 2 ? kSjKeys : kSjKeys9; }

struct SeedIndex {
    uint2* d_dir01 = nullptr;                        // [4^10] directory of tables 0 / 1
    uint2* d_dir23 = nullptr;                        // [4^9]  directory of tables 2 / 3
    uint2* d_entries0 = nullptr;                     // [n]
    uint2* d_entries1 = nullptr;                     // [n]
    uint2* d_entries23 = nullptr;                    // [18n] per key: table 2's entries, then table 3's
    int n = 0;
    bool usable = false;                             // false: a gRNA with a non-ACGT base or an overfull bucket (the set stays with the DP path)
    cudaStream_t stream = nullptr;
};

struct SeedIndexView {
    const uint2* dir01;
    const uint2* dir23;
    const uint2* entries0;
    const uint2* entries1;
    const uint2* entries23;
};

void seed_index_free(SeedIndex* ix) {
    if (!ix) return;
    dev_free(ix->d_dir01, ix->stream); dev_free(ix->d_dir23, ix->stream);
    dev_free(ix->d_entries0, ix->stream); dev_free(ix->d_entries1, ix->stream); dev_free(ix->d_entries23, ix->stream);
    delete ix;
}

// ------------------------------------------------------------------------------------------------ index
__device__ __forceinline__ void sj_keys(const uint4 p, uint32_t& k1, uint32_t& k2) {
    k1 = p.x & kFull;                                 // positions 0..9, 2 bits each
    k2 = ((p.x >> 20) | (p.y << 12)) & kFull;         // positions 10..19
}
__device__ __forceinline__ uint32_t sj_remove(uint32_t key10, int k) {      // the 10-mer without its position k (18 bits)
    const uint32_t low = (1u << (2 * k)) - 1u;
    return (key10 & low) | ((key10 >> (2 * k + 2)) << (2 * k));
}
__device__ __forceinline__ uint2 sj_entry(const uint4 p, int q, int k) {
    const uint32_t qlo = compress_even(p.x, p.y) & kFull, qhi = compress_even(p.x >> 1, p.y >> 1) & kFull;   // bit j = low / high code bit of position j
    return make_uint2(qlo | (qhi << 20), (qhi >> 12) | ((uint32_t)k << 8) | ((uint32_t)q << 12));
}

// fill == 0: count the entries of every key (cur = counters); fill == 1: cur = running cursors, write the entries
__global__ void sj_index_kernel(const uint4* __restrict__ packed, int n, int fill, uint32_t* __restrict__ cur0, uint32_t* __restrict__ cur1,
                                uint32_t* __restrict__ cur2, uint32_t* __restrict__ cur3, uint2* __restrict__ e0, uint2* __restrict__ e1,
                                uint2* __restrict__ e2, uint2* __restrict__ e3, unsigned int* __restrict__ n_invalid) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    const uint4 p = packed[q];
    if (p.z & kFull) { if (!fill) atomicAdd(n_invalid, 1u); return; }
    uint32_t k1, k2;
    sj_keys(p, k1, k2);
    uint32_t s = atomicAdd(cur0 + k1, 1u);
    if (fill) e0[s] = sj_entry(p, q, 0);
    s = atomicAdd(cur1 + k2, 1u);
    if (fill) e1[s] = sj_entry(p, q, 0);
    for (int k = 1; k <= 9; ++k) {                    // H1 without position k: an insertion at query position k
        s = atomicAdd(cur2 + sj_remove(k1, k), 1u);
        if (fill) e2[s] = sj_entry(p, q, k);
    }
    for (int k = 0; k <= 8; ++k) {                    // H2 without its position k: an insertion at query position 10 + k
        s = atomicAdd(cur3 + sj_remove(k2, k), 1u);
        if (fill) e3[s] = sj_entry(p, q, k);      // 4-bit field: the position inside H2
    }
}

static int seed_index_build(const mg_queries* q, cudaStream_t s, SeedIndex** out) {
    SeedIndex* ix = new SeedIndex();
    ix->n = q->n;
    ix->stream = s;
    *out = ix;
    uint32_t* d_cur[kSjTables] = {nullptr, nullptr, nullptr, nullptr};
    unsigned int* d_bad = nullptr;
    std::vector<uint32_t> h((size_t)kSjKeys);
    std::vector<uint2> dir((size_t)kSjKeys);
    const size_t n1 = (size_t)std::max(1, q->n);
    cudaError_t e = cudaSuccess;
    do {
        for (int t = 0; t < kSjTables && e == cudaSuccess; ++t) {
            const size_t kb = sizeof(uint32_t) * (size_t)sj_table_keys(t);
            e = dev_alloc((void**)&d_cur[t], kb, s);
            if (e == cudaSuccess) e = cudaMemsetAsync(d_cur[t], 0, kb, s);
        }
        if (e != cudaSuccess) break;
        if ((e = dev_alloc((void**)&ix->d_dir01, sizeof(uint2) * kSjKeys, s)) != cudaSuccess) break;
        if ((e = dev_alloc((void**)&ix->d_dir23, sizeof(uint2) * kSjKeys9, s)) != cudaSuccess) break;
        if ((e = dev_alloc((void**)&ix->d_entries0, sizeof(uint2) * n1, s)) != cudaSuccess) break;
        if ((e = dev_alloc((void**)&ix->d_entries1, sizeof(uint2) * n1, s)) != cudaSuccess) break;
        if ((e = dev_alloc((void**)&ix->d_entries23, sizeof(uint2) * n1 * 18, s)) != cudaSuccess) break;
        if ((e = dev_alloc((void**)&d_bad, 4, s)) != cudaSuccess) break;
        if ((e = cudaMemsetAsync(d_bad, 0, 4, s)) != cudaSuccess) break;
        const int grid = (q->n + 255) / 256;
        if (grid) sj_index_kernel<<<grid, 256, 0, s>>>(q->d_packed, q->n, 0, d_cur[0], d_cur[1], d_cur[2], d_cur[3], nullptr, nullptr, nullptr, nullptr, d_bad);
        unsigned int bad = 0;
        if ((e = cudaMemcpyAsync(&bad, d_bad, 4, cudaMemcpyDeviceToHost, s)) != cudaSuccess) break;
        bool fits = true;
        std::vector<uint32_t> h3;
        for (int t = 0; t < kSjTables && e == cudaSuccess; ++t) {
            // exclusive scan on the host: <= 4 MB per table, once per gRNA set
            const size_t nk = (size_t)sj_table_keys(t), kb = sizeof(uint32_t) * nk;
            if ((e = cudaMemcpyAsync(h.data(), d_cur[t], kb, cudaMemcpyDeviceToHost, s)) != cudaSuccess) break;
            if ((e = cudaStreamSynchronize(s)) != cudaSuccess) break;
            if (t < 2) {
                uint32_t run = 0;
                for (size_t k = 0; k < nk; ++k) {
                    const uint32_t c = h[k];
                    if (c >= (1u << kSjCountBits01)) fits = false;
                    const uint32_t word = c ? (run << kSjCountBits01) | c : 0u;
                    if (t & 1) dir[k].y = word; else dir[k].x = word;
                    h[k] = run;
                    run += c;
                }
                if ((e = cudaMemcpyAsync(d_cur[t], h.data(), kb, cudaMemcpyHostToDevice, s)) != cudaSuccess) break;
                if (t == 1) e = cudaMemcpyAsync(ix->d_dir01, dir.data(), sizeof(uint2) * nk, cudaMemcpyHostToDevice, s);
            } else if (t == 2) {
                h3.assign(h.begin(), h.begin() + nk);            // table 2's counts wait for table 3's
                continue;
            } else {
                uint32_t run = 0;
                for (size_t k = 0; k < nk; ++k) {
                    const uint32_t c2 = h3[k], c3 = h[k];
                    if (c2 >= (1u << kSjCountBits23) || c3 >= (1u << kSjCountBits23)) fits = false;
                    dir[k] = make_uint2(run, c2 | (c3 << 16));
                    h3[k] = run;                                 // cursor of table 2
                    h[k] = run + c2;                             // cursor of table 3
                    run += c2 + c3;
                }
                if ((e = cudaMemcpyAsync(d_cur[2], h3.data(), kb, cudaMemcpyHostToDevice, s)) != cudaSuccess) break;
                if ((e = cudaMemcpyAsync(d_cur[3], h.data(), kb, cudaMemcpyHostToDevice, s)) != cudaSuccess) break;
                e = cudaMemcpyAsync(ix->d_dir23, dir.data(), sizeof(uint2) * nk, cudaMemcpyHostToDevice, s);
            }
            if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        }
        if (e != cudaSuccess) break;
        ix->usable = bad == 0 && fits;
        if (ix->usable && grid) sj_index_kernel<<<grid, 256, 0, s>>>(q->d_packed, q->n, 1, d_cur[0], d_cur[1], d_cur[2], d_cur[3], ix->d_entries0,
                                                                     ix->d_entries1, ix->d_entries23, ix->d_entries23, d_bad);
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    } while (0);
    for (int t = 0; t < kSjTables; ++t) dev_free(d_cur[t], s);
    dev_free(d_bad, s);
    if (e != cudaSuccess) { set_error("seed index: %s", cudaGetErrorString(e)); return MG_ERR_CUDA; }
    return MG_OK;
}

// ------------------------------------------------------------------------------------------------ scan
struct SjParams {
    int N, M, t, g;            // gRNA count, mismatch budget, pigeonhole thresholds
    int gapless;               // 1: the criterion allows no gap (Hamming <= M): routes R1 / R2 only
    int strand;
    int64_t wb, we;            // forward window starts owned by this call
};

struct SjText {                        // a text position's context, kept in its lane's registers
    uint32_t tlo0, tlo1, thi0, thi1;   // bit planes of the 64 strand bases from p - 12 (bit i of word 0 = base p - 12 + i)
    uint32_t okmask;                   // per candidate window wi = 0..12 (w = p - 11 + wi): owned by this call, clean and away from masks
};

// the context of lane src (any permutation of lanes; every lane of the warp must call)
__device__ __forceinline__ SjText sj_text_from(const SjText& T, int src) {
    SjText r;
    r.tlo0 = __shfl_sync(0xFFFFFFFFu, T.tlo0, src); r.tlo1 = __shfl_sync(0xFFFFFFFFu, T.tlo1, src);
    r.thi0 = __shfl_sync(0xFFFFFFFFu, T.thi0, src); r.thi1 = __shfl_sync(0xFFFFFFFFu, T.thi1, src);
    r.okmask = __shfl_sync(0xFFFFFFFFu, T.okmask, src);
    return r;
}

// 32 bases of the strand's text from strand position x0, as 2-bit codes (base i at bits 2i of hi:lo)
__device__ __forceinline__ void sj_load32(const GenomeView& gv, int strand, int64_t x0, uint32_t& lo, uint32_t& hi) {
    uint32_t inv;
    if (strand > 0) { load_window(gv, x0, lo, hi, inv); return; }
    uint32_t flo, fhi;
    load_window(gv, gv.glen - x0 - 32, flo, fhi, inv);
    auto rev2 = [](uint32_t v) { v = __brev(v); return ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1); };
    lo = ~rev2(fhi);
    hi = ~rev2(flo);
}

// forward window start of the strand window w (20 bases from strand position w)
__device__ __forceinline__ int64_t sj_fwd_window(const GenomeView& gv, int strand, int64_t w) { return strand > 0 ? w : gv.glen - w - kSjL; }

// w lies in one of the sorted, disjoint window ranges [r[2i], r[2i+1]) the host keeps for the exhaustive kernels
__device__ __forceinline__ bool sj_delegated(const int64_t* __restrict__ r, int n, int64_t w) {
    int lo = 0, hi = n - 1, ans = -1;
    while (lo <= hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(r + 2 * mid) <= w) { ans = mid; lo = mid + 1; } else hi = mid - 1;
    }
    return ans >= 0 && w < __ldg(r + 2 * ans + 1);
}

__device__ __forceinline__ uint32_t sj_diag(const SjText& T, uint32_t qlo, uint32_t qhi, int o) {
    const uint32_t tl = __funnelshift_r(T.tlo0, T.tlo1, (unsigned)o), th = __funnelshift_r(T.thi0, T.thi1, (unsigned)o);
    return ((tl ^ qlo) | (th ^ qhi)) & kFull;        // bit j: query position j mismatches the text at offset o + j
}

// The canonical alignment of a clean, unmasked anchor and the route it implies (0: the anchor is not problematic).
__device__ __noinline__ int sj_canonical(uint32_t D0, uint32_t DM, uint32_t DP, int M, int t, bool gapless, int& kq) {
    const int h = __popc(D0), c2 = __popc(D0 & kH2M);
    kq = 0;
    if (gapless) return h <= M ? (c2 <= t ? 1 : 2) : 0;          // no gap budget: a trimmed end costs what its mismatch cost
    if (h <= M || (h == M + 1 && (D0 & 0x80001u))) return c2 <= t ? 1 : 2;
    int a = 0, b = h;                                 // deletion before k: #mm(DM, [0,k)) + #mm(D0, [k,20))
    for (int k = 1; k <= 19; ++k) {
        a += (int)((DM >> (k - 1)) & 1u);
        b -= (int)((D0 >> (k - 1)) & 1u);
        if (a + b <= M) {
            kq = k;
            if (k <= 9) return c2 <= t ? 1 : 5;
            if (k == 10) return c2 <= t ? 1 : 3;
            return __popc(DM & kH1M) <= t ? 3 : 7;
        }
    }
    a = 0; b = h - (int)(D0 & 1u);                    // insertion at k: #mm(DP, [0,k)) + #mm(D0, [k+1,20))
    for (int k = 1; k <= 18; ++k) {
        a += (int)((DP >> (k - 1)) & 1u);
        b -= (int)((D0 >> k) & 1u);
        if (a + b <= M) {
            kq = k;
            if (k <= 9) return c2 <= t ? 1 : 6;
            return __popc(DP & kH1M) <= t ? 4 : 8;
        }
    }
    return 0;
}

// Stage 2 (the few per cent of the candidates that pass the cheap necessary test of stage 1, again 32 at a time): the exact
// decision and the ownership rule.  Every lane of the warp must call (the text context comes by shuffle).
// record: x, y = index entry; z = [3:0] wi, [7:4] route, [12:8] k_probe, [17:13] lane that owns the text context
__device__ __noinline__ void sj_stage2(int M, int t, bool gapless, SjText T, bool act, uint4 rec, uint32_t* __restrict__ counts) {
    const int wi = (int)(rec.z & 15u), r = (int)((rec.z >> 4) & 15u), k_probe = (int)((rec.z >> 8) & 31u), src = (int)((rec.z >> 13) & 31u);
    const uint32_t qlo = rec.x & kFull, qhi = ((rec.x >> 20) | (rec.y << 12)) & kFull, id = rec.y >> 12;
    const SjText Ts = sj_text_from(T, src);
    if (!act) return;
    const int o = wi + 1;
    const uint32_t D0 = sj_diag(Ts, qlo, qhi, o), DM = sj_diag(Ts, qlo, qhi, o - 1), DP = sj_diag(Ts, qlo, qhi, o + 1);
    int kq = 0;
    const int owner = sj_canonical(D0, DM, DP, M, t, gapless, kq);
    if (owner != r || (r >= 5 && kq != k_probe)) return;
    atomicAdd(counts + id, 1u);
}

// Variant table (the same for every text position): [19:0] substitution mask on the key, [21:20] class (0: the 10-mer at p and
// its Hamming neighbours -> tables 0 / 1; 1: the 11-mer at p without its base k, substituted -> tables 0 / 1; 2: the 9-mer at
// p, substituted -> tables 2 / 3), [25:22] k.  One loop: the 32 lanes of a warp work on the same variant of 32 text positions.
static std::vector<uint32_t> sj_variants(int t, int g) {
    std::vector<uint32_t> v;
    auto push = [&](uint32_t mask, uint32_t cls, uint32_t k) { v.push_back(mask | (cls << 20) | (k << 22)); };
    push(0, 0, 0);
    if (t >= 1)
        for (int i = 0; i < 10; ++i)
            for (uint32_t x = 1; x <= 3; ++x) {
                push(x << (2 * i), 0, 0);
                if (t >= 2)
                    for (int j = i + 1; j < 10; ++j)
                        for (uint32_t y = 1; y <= 3; ++y) push((x << (2 * i)) | (y << (2 * j)), 0, 0);
            }
    if (g < 0) return v;
    for (uint32_t k = 1; k <= 9; ++k) {                     // deletion: the 11-mer at p without its base k
        push(0, 1, k);
        if (g >= 1)
            for (int s = 0; s < 10; ++s)
                for (uint32_t x = 1; x <= 3; ++x) push(x << (2 * s), 1, k);
    }
    push(0, 2, 0);                                           // insertion: the 9-mer at p (the index knows which base was left out)
    if (g >= 1)
        for (int s = 0; s < 9; ++s)
            for (uint32_t x = 1; x <= 3; ++x) push(x << (2 * s), 2, 0);
    return v;
}

constexpr int kSjWarps = 8;                                  // warps per CTA
constexpr int kSjQueue = 160;                                // candidate records per warp and queue (4 B each): < 32 left over + 32 lanes x kSjRound
constexpr int kSjRound = 4;                                  // records a lane may append per table and round
constexpr int kSjSurvivors = 64;                             // stage-1 survivors per warp (16 B each): < 32 left over + 32
constexpr int kSjItemsPerLane = 4;                           // class 2: items (32 entries) a lane lists per round
constexpr int kSjItems = 32 * kSjItemsPerLane;               // list capacity per warp


// One warp = 32 consecutive text positions, one lane each; the context of a position (64 bases as bit planes) stays in its
// lane's registers and travels by shuffle.
//
// Classes 0 / 1 (tables 0 / 1, ~1 entry per bucket at 10^6 gRNA).  The buckets of the 32 lanes hold 0, 1, 2 ... entries, and
// evaluating them where they were found would leave most lanes idle, so the lanes append (entry number, lane, k) records to
// a per-warp queue in shared memory - one queue per table, so the 32 candidates of a batch take the SAME route - and the
// queue is evaluated 32 records at a time, one per lane.  A popped batch first only ISSUES its 32 entry loads and is
// evaluated one variant later, so neither the directory load (issued one variant ahead) nor the entry load is waited for.
// Class 2 (tables 2 / 3, ~34 entries per bucket at 10^6 gRNA): the warp walks the buckets of one lane after the other, all
// lanes reading consecutive entries (coalesced), the lane's context broadcast by shuffle; no queue.
// The few per cent that pass stage 1 go to a survivor queue and are decided 32 at a time as well (sj_stage2).
// PHASE 0: classes 0 / 1 (queues), PHASE 1: class 2 (bucket walks) - two launches, so that each keeps only its own state in
// registers (64 registers, 4 CTAs per SM, no spills).
// Q2 (phase 0 only): class 2 goes through the queues as well (short 9-mer buckets: gRNA sets below kSjWalkMinN).
template <int PHASE, bool Q2>
__global__ void __launch_bounds__(32 * kSjWarps, 4)
sj_scan_kernel(GenomeView gv, SeedIndexView ix, SjParams prm, const uint32_t* __restrict__ variants, int n01, int n_queued, int n_variants,
               int64_t p_begin, int64_t p_end, const int64_t* __restrict__ delegated, int n_delegated, uint32_t* __restrict__ counts,
               unsigned long long* __restrict__ stats) {
    // PHASE 0: two queues of 4-byte records; PHASE 1: the item list and the cp.async ring
    constexpr int kBufBytes = PHASE == 0 ? 2 * kSjQueue * 4 : kSjItems * 16 + 4 * 32 * 8;
    __shared__ __align__(16) unsigned char s_buf[kSjWarps][kBufBytes];
    __shared__ uint4 s_surv[kSjWarps][kSjSurvivors];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t* q0 = reinterpret_cast<uint32_t*>(s_buf[warp]);
    uint32_t* q1 = q0 + kSjQueue;
    uint4* qs = s_surv[warp];
    const int64_t p = p_begin + ((int64_t)blockIdx.x * kSjWarps + warp) * 32 + lane;
    const int M = prm.M, t = prm.t;
    const bool gapless = prm.gapless != 0;
    uint32_t* cnt = counts + (prm.strand > 0 ? 0 : prm.N);
    SjText T;
    T.okmask = 0u;
    T.tlo0 = T.tlo1 = T.thi0 = T.thi1 = 0u;
    uint32_t T10 = 0u, U11 = 0u, S9 = 0u;
    if (p < p_end) {
        for (int wi = 0; wi <= 12; ++wi) {
            const int64_t wf = sj_fwd_window(gv, prm.strand, p - 11 + wi);
            if (wf < prm.wb || wf >= prm.we || sj_delegated(delegated, n_delegated, wf)) continue;
            T.okmask |= 1u << wi;
        }
    }
    if (!__any_sync(0xFFFFFFFFu, T.okmask != 0u)) return;          // the whole warp sits in padding / outside the shard
    if (T.okmask) {
        uint32_t lo0, hi0, lo1, hi1;
        sj_load32(gv, prm.strand, p - 12, lo0, hi0);
        sj_load32(gv, prm.strand, p + 20, lo1, hi1);
        T.tlo0 = compress_even(lo0, hi0); T.thi0 = compress_even(lo0 >> 1, hi0 >> 1);
        T.tlo1 = compress_even(lo1, hi1); T.thi1 = compress_even(lo1 >> 1, hi1 >> 1);
        const unsigned long long W = (((unsigned long long)hi0 << 32) | lo0) >> 24;   // 2-bit codes from base p on (20 bases)
        T10 = (uint32_t)W & kFull; U11 = (uint32_t)W & 0x3FFFFFu; S9 = (uint32_t)W & 0x3FFFFu;
    }
    const bool live = T.okmask != 0u;
    int qlen0 = 0, qlen1 = 0, n_surv = 0;                           // warp-uniform
    unsigned int st_cand = 0u, st_surv = 0u;                        // candidates looked at / stage-1 survivors (mg_last_seedjoin_stats)
    int cur_cls = 0;                                                // class of the records in the queues
    int pend0 = 0, pend1 = 0;                                       // records of the popped batch whose entry loads are in flight
    uint32_t prec0 = 0u, prec1 = 0u;
    uint2 pent0 = make_uint2(0u, 0u), pent1 = make_uint2(0u, 0u);

    // stage-1 survivors, compacted; a full batch goes through stage 2 at once
    auto survive = [&](bool pass, const uint2 e, int wi, int r, int kp, int src) {
        const unsigned m = __ballot_sync(0xFFFFFFFFu, pass);
        if (m == 0u) return;
        if (pass) qs[n_surv + __popc(m & ((1u << lane) - 1u))] = make_uint4(e.x, e.y, (uint32_t)wi | ((uint32_t)r << 4) | ((uint32_t)kp << 8) | ((uint32_t)src << 13), 0u);
        n_surv += __popc(m);
        st_surv += __popc(m);
        if (n_surv >= 32) {
            __syncwarp();
            sj_stage2(M, t, gapless, T, true, qs[n_surv - 32 + lane], cnt);
            n_surv -= 32;
            __syncwarp();
        }
    };

    // Stage 1 of the pending batch of queue j: cheap NECESSARY conditions for "the anchor is problematic and its canonical
    // alignment implies this route (and k)".  An insertion skips one query position, so bounds that use "mismatches on both
    // diagonals" get one unit of slack.
    auto eval = [&](auto J) {
        constexpr int j = decltype(J)::value;
        const bool act = lane < (j ? pend1 : pend0);
        const uint32_t rec = j ? prec1 : prec0;
        const uint2 e = j ? pent1 : pent0;
        const int src = (int)(rec >> 27), kv = (int)((rec >> 20) & 15u);
        const SjText Ts = sj_text_from(T, src);
        const uint32_t ok = act ? Ts.okmask : 0u;
        const uint32_t qlo = e.x & kFull, qhi = ((e.x >> 20) | (e.y << 12)) & kFull;
        if (cur_cls == 0 && j == 0) {
            // the text's 10-mer at p is H1: on the final diagonal (R2, w = p), on w - 1 (R3, w = p + 1), on w + 1 (R4, w = p - 1)
            const uint32_t d11 = sj_diag(Ts, qlo, qhi, 11), d12 = sj_diag(Ts, qlo, qhi, 12), d13 = sj_diag(Ts, qlo, qhi, 13);
            const int c1 = __popc(d12 & kH1M);
            const bool p2 = ((ok >> 11) & 1u) && __popc(d12) <= (gapless ? M : M + 1);           // R2 only owns ungapped alignments
            const bool p3 = !gapless && ((ok >> 12) & 1u) && c1 + __popc(d12 & d13 & kH2M) <= M;
            const bool p4 = !gapless && ((ok >> 10) & 1u) && c1 + __popc(d12 & d11 & kH2M) <= M + 1;
            if (__any_sync(0xFFFFFFFFu, p2 || p3 || p4)) {
                survive(p2, e, 11, 2, 0, src);
                survive(p3, e, 12, 3, 0, src);
                survive(p4, e, 10, 4, 0, src);
            }
        } else if (cur_cls == 0) {
            // the 10-mer at p is H2 on the final diagonal (R1): w = p - 10
            const uint32_t DM = sj_diag(Ts, qlo, qhi, 1), D0 = sj_diag(Ts, qlo, qhi, 2), DP = sj_diag(Ts, qlo, qhi, 3);
            const bool p1 = ((ok >> 1) & 1u) && (gapless ? __popc(D0) <= M
                                                         : min(__popc(DM & D0 & kH1M), __popc(DP & D0 & kH1M)) + __popc(D0 & kH2M) <= M + 1);
            survive(p1, e, 1, 1, 0, src);
        } else if (Q2 && cur_cls == 2) {
            // the 9-mer at p is H1 without its position ke (R6: w = p - 1) or H2 without its position ke (R8: w + 1 = p - 10), see
            // the bucket walk below; here for short buckets (small gRNA sets)
            const int ke = (int)((e.y >> 8) & 15u), wi = j ? 0 : 10, kp = j ? 10 + ke : ke;
            const uint32_t lowk = (1u << kp) - 1u;
            const bool p68 = ((ok >> wi) & 1u) && __popc(sj_diag(Ts, qlo, qhi, wi + 2) & lowk) + __popc(sj_diag(Ts, qlo, qhi, wi + 1) & ~(2u * lowk + 1u)) <= M;
            survive(p68, e, wi, j ? 8 : 6, kp, src);
        } else if (j == 0) {
            // a deletion before query position k of H1 (R5): H1 starts at p on w - 1, w = p + 1; the deletion must fit
            const uint32_t lowk = (1u << kv) - 1u;
            const bool p5 = ((ok >> 12) & 1u) && __popc(sj_diag(Ts, qlo, qhi, 12) & lowk) + __popc(sj_diag(Ts, qlo, qhi, 13) & ~lowk) <= M;
            survive(p5, e, 12, 5, kv, src);
        } else {
            // a deletion before query position 10 + k (R7): H2's first base is at p on w - 1, w - 1 = p - 10
            const uint32_t lowk = (1u << (10 + kv)) - 1u;
            const bool p7 = ((ok >> 2) & 1u) && __popc(sj_diag(Ts, qlo, qhi, 2) & lowk) + __popc(sj_diag(Ts, qlo, qhi, 3) & ~lowk) <= M;
            survive(p7, e, 2, 7, 10 + kv, src);
        }
        if (j) pend1 = 0; else pend0 = 0;
    };
    // take up to 32 records off queue j and issue their entry loads
    auto pop = [&](auto J) {
        constexpr int j = decltype(J)::value;
        int& qlen = j ? qlen1 : qlen0;
        const int take = qlen < 32 ? qlen : 32;
        const uint32_t rec = lane < take ? (j ? q1 : q0)[qlen - take + lane] : 0u;
        const uint2* table = Q2 && cur_cls == 2 ? ix.entries23 : (j ? ix.entries1 : ix.entries0);
        const uint2 e = lane < take ? __ldg(table + (rec & (Q2 && cur_cls == 2 ? 0x7FFFFFFu : kFull))) : make_uint2(0u, 0u);
        if (j) { prec1 = rec; pent1 = e; pend1 = take; } else { prec0 = rec; pent0 = e; pend0 = take; }
        qlen -= take;
        __syncwarp();
    };
    auto service = [&](auto J) {                                     // keep the queue below one batch
        constexpr int j = decltype(J)::value;
        while ((j ? qlen1 : qlen0) >= 32) { if (j ? pend1 : pend0) eval(J); pop(J); }
    };
    auto flush = [&](auto J) {
        constexpr int j = decltype(J)::value;
        while ((j ? qlen1 : qlen0) > 0 || (j ? pend1 : pend0) > 0) { if (j ? pend1 : pend0) eval(J); if ((j ? qlen1 : qlen0) > 0) pop(J); }
    };
    const std::integral_constant<int, 0> J0;
    const std::integral_constant<int, 1> J1;

    // directory words of variant v (classes 0 / 1) for this lane: issues the load
    auto dir01 = [&](int v) -> uint2 {
        if (v >= n_queued) return make_uint2(0u, 0u);
        const uint32_t d = __ldg(variants + v);
        const uint32_t mask = d & kFull;
        if (Q2 && ((d >> 20) & 3u) == 2u) {                         // class 2 through the queues (short buckets)
            return live ? __ldg(ix.dir23 + (S9 ^ mask)) : make_uint2(0u, 0u);     // x = first entry, y = count 2 | count 3 << 16
        }
        uint32_t key = T10 ^ mask;
        bool valid = live;
        if ((d >> 20) & 3u) {
            // the 11-mer at p without its base k; two equal neighbours give the same string whichever is removed: the first one
            // (smaller k) stands for both
            const int k = (int)((d >> 22) & 15u);
            valid = valid && ((U11 >> (2 * k)) & 3u) != ((U11 >> (2 * k - 2)) & 3u);
            const uint32_t low = (1u << (2 * k)) - 1u;
            key = ((U11 & low) | ((U11 >> (2 * k + 2)) << (2 * k))) ^ mask;
        }
        return valid ? __ldg(ix.dir01 + key) : make_uint2(0u, 0u);
    };

    uint2 Bn = PHASE == 0 ? dir01(0) : make_uint2(0u, 0u);
    for (int v = 0; PHASE == 0 && v < n_queued; ++v) {
        const uint32_t d = __ldg(variants + v);
        const int cls = (int)((d >> 20) & 3u);
        const uint2 B = Bn;
        Bn = dir01(v + 1);                                          // in flight while this variant is queued and evaluated
        if (cls != cur_cls) { flush(J0); flush(J1); cur_cls = cls; }   // a queue only ever holds one route family
        if (pend0) eval(J0);
        if (pend1) eval(J1);
        if (!__any_sync(0xFFFFFFFFu, (Q2 && cls == 2 ? B.y : (B.x | B.y)) != 0u)) continue;  // no lane found anything: the common case for small gRNA sets
        const uint32_t tag = ((uint32_t)lane << 27) | (((d >> 22) & 15u) << 20);     // record: [31:27] lane, [23:20] k, entry number below
        uint32_t a0 = B.x >> kSjCountBits01, n0 = B.x & ((1u << kSjCountBits01) - 1u);
        uint32_t a1 = B.y >> kSjCountBits01, n1 = B.y & ((1u << kSjCountBits01) - 1u);
        if (Q2 && cls == 2) { a0 = B.x; n0 = B.y & 0xFFFFu; a1 = B.x + n0; n1 = B.y >> 16; }
        do {
            // every lane appends up to kSjRound records per table; one scan for both counts (16 bits each)
            const int h0 = (int)min((uint32_t)kSjRound, n0), h1 = (int)min((uint32_t)kSjRound, n1);
            int off = h0 | (h1 << 16);
#pragma unroll
            for (int dd = 1; dd < 32; dd <<= 1) { const int o = __shfl_up_sync(0xFFFFFFFFu, off, dd); if (lane >= dd) off += o; }
            const int total = __shfl_sync(0xFFFFFFFFu, off, 31);
            if (total == 0) break;
            off -= h0 | (h1 << 16);
            const int o0 = qlen0 + (off & 0xFFFF), o1 = qlen1 + (off >> 16);
#pragma unroll
            for (int j = 0; j < kSjRound; ++j) {
                if (j < h0) q0[o0 + j] = (a0 + j) | tag;
                if (j < h1) q1[o1 + j] = (a1 + j) | tag;
            }
            a0 += h0; n0 -= h0; a1 += h1; n1 -= h1;
            qlen0 += total & 0xFFFF;
            qlen1 += total >> 16;
            st_cand += (total & 0xFFFF) + (total >> 16);
            __syncwarp();
            service(J0);
            service(J1);
        } while (__any_sync(0xFFFFFFFFu, (n0 | n1) != 0u));
    }
    if (PHASE == 0) {
        flush(J0);
        flush(J1);
    }

    // class 2: the text's 9-mer at p is H1 without its position ke (R6: an insertion at query position ke, H1 starts at p on
    // w + 1, w = p - 1) or H2 without its position ke (R8: insertion at 10 + ke, w + 1 = p - 10); the insertion must fit
    auto dir23 = [&](int v) -> uint2 {
        if (v >= n_variants || !live) return make_uint2(0u, 0u);
        return __ldg(ix.dir23 + (S9 ^ (__ldg(variants + v) & kFull)));
    };
    // Work items = (lane whose bucket is walked, 32 consecutive entries of it).  Per variant the lanes write their items into a
    // list in shared memory (one scan); the warp then runs down the list, all lanes reading consecutive entries (coalesced).
    // The entries travel global -> shared by cp.async three items ahead, each lane reading back only its own slot, so no
    // register is tied up by a load in flight and nothing has to be rotated.
    if (PHASE == 1) {
        uint4* items = reinterpret_cast<uint4*>(q0);                // kSjItems x 16 B: x = first entry, y = n2 | nn << 16, z = lane | base << 8
        uint2* ring = reinterpret_cast<uint2*>(items + kSjItems);   // 4 slots x 32 lanes
        uint2 Cn = dir23(n_queued);
        int t_src = -1;
        SjText Ts = T;
        for (int v = n_queued; v < n_variants; ++v) {
            const uint2 C = Cn;
            Cn = dir23(v + 1);
            uint32_t done = 0u;                                     // entries of this lane's bucket already listed
            const uint32_t n2 = C.y & 0xFFFFu, nn = n2 + (C.y >> 16);
            for (;;) {
                // every lane lists up to kSjItemsPerLane items of its bucket per round
                const uint32_t left = nn - done;
                const int c = (int)min((uint32_t)kSjItemsPerLane, (left + 31u) >> 5);
                int off = c;
#pragma unroll
                for (int dd = 1; dd < 32; dd <<= 1) { const int o = __shfl_up_sync(0xFFFFFFFFu, off, dd); if (lane >= dd) off += o; }
                const int n_items = __shfl_sync(0xFFFFFFFFu, off, 31);
                if (n_items == 0) break;
                off -= c;
#pragma unroll
                for (int j = 0; j < kSjItemsPerLane; ++j)
                    if (j < c) items[off + j] = make_uint4(C.x, C.y, (uint32_t)lane | ((done + 32u * j) << 8), 0u);
                done += 32u * c;
                st_cand += __reduce_add_sync(0xFFFFFFFFu, min(left, 32u * c));
                __syncwarp();
                auto issue = [&](int it) {
                    if (it < n_items) {
                        const uint4 d = items[it];
                        const uint32_t i = (d.z >> 8) + lane;
                        if (i < (d.y & 0xFFFFu) + (d.y >> 16)) {
                            const unsigned dst = (unsigned)__cvta_generic_to_shared(ring + (it & 3) * 32 + lane);
                            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(ix.entries23 + d.x + i) : "memory");
                        }
                    }
                    asm volatile("cp.async.commit_group;" ::: "memory");
                };
                issue(0); issue(1); issue(2);
                for (int it = 0; it < n_items; ++it) {
                    asm volatile("cp.async.wait_group 2;" ::: "memory");
                    const uint4 d = items[it];
                    const uint2 e = ring[(it & 3) * 32 + lane];
                    const int src = (int)(d.z & 31u);
                    if (src != t_src) { Ts = sj_text_from(T, src); t_src = src; }
                    const uint32_t i = (d.z >> 8) + lane, in2 = d.y & 0xFFFFu, inn = in2 + (d.y >> 16);
                    const bool t3 = i >= in2;
                    const int ke = (int)((e.y >> 8) & 15u), wi = t3 ? 0 : 10, kp = t3 ? 10 + ke : ke;
                    const uint32_t qlo = e.x & kFull, qhi = ((e.x >> 20) | (e.y << 12)) & kFull, lowk = (1u << kp) - 1u;
                    const bool pass = i < inn && ((Ts.okmask >> wi) & 1u) &&
                                      __popc(sj_diag(Ts, qlo, qhi, wi + 2) & lowk) + __popc(sj_diag(Ts, qlo, qhi, wi + 1) & ~(2u * lowk + 1u)) <= M;
                    survive(pass, e, wi, t3 ? 8 : 6, kp, src);
                    issue(it + 3);                                   // slot (it + 3) & 3 was read one item ago
                }
                asm volatile("cp.async.wait_group 0;" ::: "memory");
                __syncwarp();
            }
        }
    }
    __syncwarp();
    sj_stage2(M, t, gapless, T, lane < n_surv, lane < n_surv ? qs[lane] : make_uint4(0u, 0u, 0u, 0u), cnt);
    if (lane == 0) { atomicAdd(stats, (unsigned long long)st_cand); atomicAdd(stats + 1, (unsigned long long)st_surv); }
}

// 32-base blocks of [b0, b1) that hold an invalid base
__global__ void sj_dirty_blocks_kernel(GenomeView gv, int64_t b0, int64_t b1, unsigned int cap, unsigned int* __restrict__ n_out,
                                       int64_t* __restrict__ out) {
    const int64_t b = b0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= b1) return;
    if (__ldg(gv.invalid + b) != 0u) {
        const unsigned int slot = atomicAdd(n_out, 1u);
        if (slot < cap) out[slot] = b;
    }
}

// [0] candidates (index entries compared with the text), [1] stage-1 survivors of the launches since the last reset, per device
constexpr int kSjMaxDevices = 64;
constexpr int64_t kSjMinInterior = 8192;             // windows
constexpr long long kSjWalkMinN = 400000;            // gRNA: from here the 9-mer buckets are long enough for the cooperative walk
constexpr int64_t kSjMergeGap = 1024;                // windows
constexpr size_t kSjMaxRanges = 4096;
static unsigned long long* g_sj_stats[kSjMaxDevices] = {nullptr};

static int sj_stats_buffer(unsigned long long** out) {
    int dev = 0;
    MG_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= kSjMaxDevices) { set_error("device index %d out of range", dev); return MG_ERR_ARG; }
    if (!g_sj_stats[dev]) MG_CUDA(cudaMalloc(&g_sj_stats[dev], 16));
    *out = g_sj_stats[dev];
    return MG_OK;
}

int seedjoin_stats(uint64_t* out) {
    int dev = 0;
    out[0] = out[1] = 0;
    MG_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= kSjMaxDevices || !g_sj_stats[dev]) return MG_OK;
    MG_CUDA(cudaMemcpy(out, g_sj_stats[dev], 16, cudaMemcpyDeviceToHost));
    return MG_OK;
}

// Can this screen use the seed-and-join path?  (criterion class, gRNA length, set size; MG_GAPPED=dp|seedjoin overrides the size rule)
bool seedjoin_eligible(const mg_queries* q, const mg_criteria* crit) {
    if (crit->n_units != 0 || crit->max_gap > 1 || !crit->pamless || q->len != kSjL) return false;
    const bool gapless = crit->max_gap == 0;
    if (crit->max_mismatch < 0 || crit->max_mismatch > (gapless ? 5 : 4)) return false;          // half-neighbourhoods of radius <= 2
    if (q->n < 1 || q->n > (1 << 20)) return false;                                              // 20-bit gRNA index in an entry
    const char* env = getenv("MG_SEEDJOIN");           // on | off: whatever the set size (MG_GAPPED=seedjoin | dp is the older spelling)
    const char* old = getenv("MG_GAPPED");
    if ((env && strcmp(env, "off") == 0) || (old && strcmp(old, "dp") == 0)) return false;
    if ((env && strcmp(env, "on") == 0) || (old && strcmp(old, "seedjoin") == 0)) return true;
    // Below these sizes the exhaustive kernels are faster: the probes per position (743 with a gap, 436 without) do not depend
    // on the number of gRNA.  One B200, 135 Mb: DP 53 us, pair-word scan 9.6 us, seed-and-join ~1.5 s / N + 4.5 us (gap) or
    // ~0.9 s / N + 1.2 us (no gap) per gRNA.
    long long min_n = gapless ? 100000 : 25000;
    if (const char* m = getenv("MG_SEEDJOIN_MIN")) min_n = atoll(m);
    return q->n >= min_n;
}

// Counts every CLEAN anchor owned by the windows [wb, we) and returns the DIRTY window ranges the caller must hand to the DP
// path.  Returns MG_ERR_UNSUPPORTED (nothing counted) if the index cannot represent the gRNA set or the genome is too ragged.
int launch_seedjoin(const mg_genome* g, const mg_queries* q, const mg_criteria* crit, const GeneralCtx* d_ctx, int64_t wb, int64_t we,
                    uint32_t* d_counts, cudaStream_t s, std::vector<std::pair<int64_t, int64_t>>& dirty_ranges) {
    dirty_ranges.clear();
    {
        // built once per query handle, whichever thread screens with it first
        static std::mutex build_lock;
        std::lock_guard<std::mutex> guard(build_lock);
        if (q->sj == nullptr) {
            SeedIndex* ix = nullptr;
            const int rc = seed_index_build(q, s, &ix);
            if (rc != MG_OK) { seed_index_free(ix); return rc; }
            q->sj = ix;
        }
    }
    const SeedIndex* ix = q->sj;
    if (!ix->usable) return MG_ERR_UNSUPPORTED;
    unsigned long long* d_stats = nullptr;
    { const int rc = sj_stats_buffer(&d_stats); if (rc != MG_OK) return rc; }
    MG_CUDA(cudaMemsetAsync(d_stats, 0, 16, s));
    const GenomeView gv = g->view();
    // dirty 32-base blocks around the owned windows -> window ranges for the DP path
    const int64_t b0 = std::max<int64_t>(0, (wb - 3) >> 5), b1 = std::min<int64_t>(g->glen >> 5, ((we + 22) >> 5) + 1);
    const unsigned int cap = 1u << 20;
    unsigned int* d_n = nullptr;
    int64_t* d_blocks = nullptr;
    MG_CUDA(dev_alloc((void**)&d_n, 4, s));
    cudaError_t e = dev_alloc((void**)&d_blocks, sizeof(int64_t) * cap, s);
    if (e != cudaSuccess) { dev_free(d_n, s); set_error("seed join: %s", cudaGetErrorString(e)); return MG_ERR_NOMEM; }
    unsigned int n_dirty = 0;
    std::vector<int64_t> blocks;
    e = cudaMemsetAsync(d_n, 0, 4, s);
    if (e == cudaSuccess && b1 > b0) {
        sj_dirty_blocks_kernel<<<(unsigned)((b1 - b0 + 255) / 256), 256, 0, s>>>(gv, b0, b1, cap, d_n, d_blocks);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(&n_dirty, d_n, 4, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e == cudaSuccess && n_dirty <= cap && n_dirty) {
        blocks.resize(n_dirty);
        e = cudaMemcpyAsync(blocks.data(), d_blocks, sizeof(int64_t) * n_dirty, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    }
    dev_free(d_n, s); dev_free(d_blocks, s);
    if (e != cudaSuccess) { set_error("seed join: %s", cudaGetErrorString(e)); return MG_ERR_CUDA; }
    if (n_dirty > cap) return MG_ERR_UNSUPPORTED;      // a very ragged subject: the DP path takes all of it
    std::vector<std::pair<int64_t, int64_t>> raw, interior;
    for (const int64_t b : blocks) raw.emplace_back(32 * b - 22, 32 * b + 35);    // windows w with [w-3, w+23) meeting block b
    if (g->n_masks > 0) {                              // windows w with [w-3, w+24) meeting a masked interval (sj_mask_near)
        std::vector<int64_t> ms(g->n_masks), me(g->n_masks);
        e = cudaMemcpyAsync(ms.data(), g->d_mstart, sizeof(int64_t) * g->n_masks, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(me.data(), g->d_mend_pm, sizeof(int64_t) * g->n_masks, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) { set_error("seed join: %s", cudaGetErrorString(e)); return MG_ERR_CUDA; }
        // the running maximum of the ends gives the same union as the ends themselves (the interval that set it starts earlier)
        for (int i = 0; i < g->n_masks; ++i) raw.emplace_back(ms[i] - 23, me[i] + 3);
        // Windows whose [w-3, w+24) lies INSIDE one masked interval (span_masked: the last interval starting at or before w-3
        // has a running-maximum end >= w+24) hold no countable anchor at all - every alignment there (<= 1 gap) is on-target:
        // neither side screens them.  Only the ~50 windows around each interval edge are left to the exhaustive kernels.
        for (int i = 0; i < g->n_masks; ++i) {
            const int64_t a = ms[i] + 3, z = std::min<int64_t>(i + 1 < g->n_masks ? ms[i + 1] + 3 : INT64_MAX, me[i] - 23);
            if (z - a >= kSjMinInterior) {          // (a short interior costs the exhaustive kernels less than a second range does)
                if (!interior.empty() && a <= interior.back().second) interior.back().second = std::max(interior.back().second, z);
                else interior.emplace_back(a, z);
            }
        }
    }
    // Merge (ranges closer than kSjMergeGap windows become one: every range is a separate launch of the exhaustive kernels, the
    // windows in between cost them next to nothing) and clip to the call's windows.  The scan kernel skips exactly these ranges.
    std::sort(raw.begin(), raw.end());
    int64_t n_delegated_windows = 0;
    for (const auto& r : raw) {
        if (!dirty_ranges.empty() && r.first <= dirty_ranges.back().second + kSjMergeGap) dirty_ranges.back().second = std::max(dirty_ranges.back().second, r.second);
        else dirty_ranges.emplace_back(r.first, r.second);
    }
    {
        size_t k = 0;
        for (auto& r : dirty_ranges) {
            r.first = std::max(wb, r.first); r.second = std::min(we, r.second);
            if (r.first < r.second) { n_delegated_windows += r.second - r.first; dirty_ranges[k++] = r; }
        }
        dirty_ranges.resize(k);
    }
    // the scan kernel skips `skip` = the merged ranges; the caller screens them minus the interiors of the masked intervals
    const std::vector<std::pair<int64_t, int64_t>> skip = dirty_ranges;
    if (!interior.empty()) {
        std::vector<std::pair<int64_t, int64_t>> rest;
        size_t j = 0;
        n_delegated_windows = 0;
        for (const auto& r : skip) {
            int64_t a = r.first;
            while (j < interior.size() && interior[j].second <= a) ++j;
            for (size_t k = j; k < interior.size() && interior[k].first < r.second; ++k) {
                if (interior[k].first > a) rest.emplace_back(a, interior[k].first);
                a = std::max(a, interior[k].second);
            }
            if (a < r.second) rest.emplace_back(a, r.second);
        }
        for (const auto& r : rest) n_delegated_windows += r.second - r.first;
        dirty_ranges.swap(rest);
    }
    // a subject that is mostly N runs / interval edges, or cut into very many pieces: the exhaustive kernels take all of it
    if (dirty_ranges.size() > kSjMaxRanges || skip.size() > kSjMaxRanges || 2 * n_delegated_windows > we - wb) { dirty_ranges.clear(); return MG_ERR_UNSUPPORTED; }
    // (the >= 64 invalid bases before the first and after the last contig make the windows at both ends of the coordinate
    // space dirty through their blocks, which is what sj_dirty assumes for wf < 3 and wf + 23 > glen)
    SjParams prm;
    prm.N = q->n; prm.M = crit->max_mismatch; prm.gapless = crit->max_gap == 0 ? 1 : 0;
    prm.t = prm.gapless ? prm.M / 2 : (prm.M + 1) / 2;                 // a half with <= t mismatches always exists
    prm.g = prm.gapless ? -1 : prm.M - prm.t - 1;
    prm.wb = wb; prm.we = we;
    SeedIndexView iv;
    iv.dir01 = ix->d_dir01; iv.dir23 = ix->d_dir23; iv.entries0 = ix->d_entries0; iv.entries1 = ix->d_entries1;
    iv.entries23 = ix->d_entries23;
    const std::vector<uint32_t> vars = sj_variants(prm.t, prm.g);
    int n01 = 0;                                       // classes 0 / 1 come first
    while (n01 < (int)vars.size() && ((vars[n01] >> 20) & 3u) != 2u) ++n01;
    uint32_t* d_vars = nullptr;
    int64_t* d_ranges = nullptr;
    std::vector<int64_t> flat;
    for (const auto& r : skip) { flat.push_back(r.first); flat.push_back(r.second); }
    MG_CUDA(dev_alloc((void**)&d_vars, sizeof(uint32_t) * vars.size(), s));
    e = dev_alloc((void**)&d_ranges, sizeof(int64_t) * std::max<size_t>(2, flat.size()), s);
    if (e != cudaSuccess) { dev_free(d_vars, s); set_error("seed join: %s", cudaGetErrorString(e)); return MG_ERR_NOMEM; }
    e = cudaMemcpyAsync(d_vars, vars.data(), sizeof(uint32_t) * vars.size(), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess && !flat.empty()) e = cudaMemcpyAsync(d_ranges, flat.data(), sizeof(int64_t) * flat.size(), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);              // vars and flat live on this stack frame
    if (e != cudaSuccess) { dev_free(d_vars, s); dev_free(d_ranges, s); set_error("seed join: %s", cudaGetErrorString(e)); return MG_ERR_CUDA; }
    // The 9-mer buckets hold ~N / 29 000 entries per table: long ones are walked by the whole warp (phase 1), short ones would
    // leave most lanes idle there and go through the queues of phase 0 like the 10-mer buckets (MG_SJ_WALK_MIN overrides)
    long long walk_min = kSjWalkMinN;
    if (const char* w = getenv("MG_SJ_WALK_MIN")) walk_min = atoll(w);
    const int n_queued = q->n < walk_min ? (int)vars.size() : n01;
    for (int z = 0; z < 2 && e == cudaSuccess; ++z) {
        prm.strand = z == 0 ? 1 : -1;
        // strand windows of the owned forward windows, then the positions that can reach them (w - 1 .. w + 11)
        const int64_t ws0 = z == 0 ? wb : g->glen - we - kSjL + 1, ws1 = z == 0 ? we : g->glen - wb - kSjL + 1;
        int64_t p0 = std::max<int64_t>(32, ws0 - 2), p1 = std::min<int64_t>(g->glen - 64, ws1 + 12);
        if (p1 <= p0) continue;
        const unsigned grid = (unsigned)((p1 - p0 + 32 * kSjWarps - 1) / (32 * kSjWarps));
        if (n_queued > n01)
            sj_scan_kernel<0, true><<<grid, 32 * kSjWarps, 0, s>>>(gv, iv, prm, d_vars, n01, n_queued, (int)vars.size(), p0, p1, d_ranges, (int)skip.size(), d_counts, d_stats);
        else
            sj_scan_kernel<0, false><<<grid, 32 * kSjWarps, 0, s>>>(gv, iv, prm, d_vars, n01, n_queued, (int)vars.size(), p0, p1, d_ranges, (int)skip.size(), d_counts, d_stats);
        if (n_queued < (int)vars.size())
            sj_scan_kernel<1, false><<<grid, 32 * kSjWarps, 0, s>>>(gv, iv, prm, d_vars, n01, n_queued, (int)vars.size(), p0, p1, d_ranges, (int)skip.size(), d_counts, d_stats);
        e = cudaGetLastError();
        if (e != cudaSuccess) break;
    }
    dev_free(d_vars, s);
    dev_free(d_ranges, s);
    if (e != cudaSuccess) { set_error("seed join: %s", cudaGetErrorString(e)); return MG_ERR_CUDA; }
    return MG_OK;
}

}  // namespace mg
