// Shared declarations of libminorg_b200 (sm_100a).  See include/minorg_b200.h for the C ABI.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/minorg_b200.h"

namespace mg {

void set_error(const char* fmt, ...);
// Stream-ordered allocations from the device's default memory pool (kept warm: cudaFree of a 0.4 GB plane
// costs hundreds of ms, a pool free costs microseconds - matters for the host-buffer end-to-end call).
cudaError_t dev_alloc(void** p, size_t bytes, cudaStream_t s);
void dev_free(void* p, cudaStream_t s);

#define MG_CUDA(call)                                                                        \
    do {                                                                                     \
        cudaError_t e__ = (call);                                                            \
        if (e__ != cudaSuccess) {                                                            \
            mg::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return MG_ERR_CUDA;                                                              \
        }                                                                                    \
    } while (0)

#define MG_CHECK_ARG(cond)                                                    \
    do {                                                                      \
        if (!(cond)) {                                                        \
            mg::set_error("%s:%d: invalid argument: %s", __FILE__, __LINE__, #cond); \
            return MG_ERR_ARG;                                                \
        }                                                                     \
    } while (0)

constexpr int kPad = 64;        // invalid bases between contigs (>= L + gaps + PAM flank)
constexpr int kAlign = 32;      // contig global starts are multiples of this (one invalid-plane word)
constexpr int kCodeInvalid = 4;
constexpr int kMinHalo = 256;   // smallest halo of a chunk genome: gRNA + gaps + PAM flank + the 48-base window loads

// What the kernels see of a genome.
struct GenomeView {
    const uint32_t* bases;    // 2 bits/base, 16 bases per word, base i at bits [2(i%16), 2(i%16)+1]
    const uint32_t* invalid;  // 1 bit/base, 32 bases per word
    int64_t glen;             // global length in bases (multiple of 32; buffers hold 4 spare words)
    const int64_t* cstart;    // [n_contigs] global start (sorted ascending)
    const int64_t* cend;      // [n_contigs] global end
    int n_contigs;
    const int64_t* mstart;    // [n_masks] global mask starts, ascending
    const int64_t* mend_pm;   // [n_masks] running maximum of global mask ends
    int n_masks;
};

struct QueriesView {
    const uint8_t* codes;     // [2N][L] base codes 0..3, 4 = invalid; query N+i = revcomp of gRNA i
    const uint32_t* tab;      // [n_groups][4L+1] bit-sliced mismatch words (+ one all-ones word)
    const uint32_t* tab2;     // [n_groups][16*ceil(L/2)+1] pair words: some base of the position PAIR mismatches
    const uint4* packed;      // [2N] {2-bit codes of positions 0..15, of 16..31, bit j = position j is not ACGT, 0}
    int n;                    // N
    int len;                  // L
    int n_groups;             // ceil(2N/32)
};

__device__ __forceinline__ uint32_t ld_bases(const uint32_t* p) { return __ldg(p); }

// Bases [p, p+32) as 64 bits (lo = bases 0..15) and their invalid bits.
__device__ __forceinline__ void load_window(const GenomeView& g, int64_t p, uint32_t& lo, uint32_t& hi, uint32_t& inv) {
    const int64_t wi = p >> 4;
    const unsigned sh = (unsigned)(p & 15) * 2u;
    const uint32_t b0 = __ldg(g.bases + wi), b1 = __ldg(g.bases + wi + 1), b2 = __ldg(g.bases + wi + 2);
    lo = __funnelshift_r(b0, b1, sh);
    hi = __funnelshift_r(b1, b2, sh);
    const int64_t ii = p >> 5;
    inv = __funnelshift_r(__ldg(g.invalid + ii), __ldg(g.invalid + ii + 1), (unsigned)(p & 31));
}

__device__ __forceinline__ int base_at(const GenomeView& g, int64_t p) {   // 0..3 or 4
    if (p < 0 || p >= g.glen) return kCodeInvalid;
    if ((__ldg(g.invalid + (p >> 5)) >> (p & 31)) & 1u) return kCodeInvalid;
    return (__ldg(g.bases + (p >> 4)) >> ((p & 15) * 2)) & 3u;
}

// Index of the contig whose [cstart - kPad/2, cend + kPad/2) neighbourhood holds p, or -1.
__device__ __forceinline__ int contig_of(const GenomeView& g, int64_t p) {
    int lo = 0, hi = g.n_contigs - 1, ans = -1;
    while (lo <= hi) {
        const int mid = (lo + hi) >> 1;
        if (g.cstart[mid] - kPad / 2 <= p) { ans = mid; lo = mid + 1; } else hi = mid - 1;
    }
    return ans;
}

// True iff [a,b) (global) lies inside a single masked interval.
__device__ __forceinline__ bool span_masked(const GenomeView& g, int64_t a, int64_t b) {
    int lo = 0, hi = g.n_masks - 1, ans = -1;
    while (lo <= hi) {
        const int mid = (lo + hi) >> 1;
        if (g.mstart[mid] <= a) { ans = mid; lo = mid + 1; } else hi = mid - 1;
    }
    return ans >= 0 && g.mend_pm[ans] >= b;
}

}  // namespace mg

namespace mg {
struct SeedIndex;                         // half-gRNA hash index of the seed-and-join screen (screen_seedjoin.cu), built on first use
void seed_index_free(SeedIndex* ix);
}

// Concrete handle types behind the opaque C names.
struct mg_genome {
    int device = 0;
    int n_contigs = 0;
    std::vector<int64_t> clen, cstart, cend;   // host copies
    int64_t glen = 0;                           // global length (bases)
    int64_t total = 0;
    uint32_t* d_bases = nullptr;
    uint32_t* d_invalid = nullptr;
    int64_t* d_cstart = nullptr;
    int64_t* d_cend = nullptr;
    int64_t* d_mstart = nullptr;
    int64_t* d_mend_pm = nullptr;
    int n_masks = 0;
    size_t bases_words = 0, invalid_words = 0;
    // A CHUNK of a subject (mg_genome_from_*_chunk): the coordinate space is the whole subject's, but the planes only
    // hold the global range [lo, hi) = the owned window range [own_begin, own_end) plus a halo on both sides.  The
    // kernels index the planes with global positions through pointers biased by lo (multiple of 128 bases).
    int64_t lo = 0, hi = 0;                     // stored global range (whole genome: [0, glen))
    int64_t own_begin = 0, own_end = 0;         // window starts this handle owns (whole genome: [0, glen))
    int64_t halo = 0;
    bool is_chunk = false;
    mutable cudaStream_t last_stream = nullptr; // stream of the most recent screen: the planes are freed in its order
    mg::GenomeView view() const {
        mg::GenomeView v;
        v.bases = d_bases - (lo >> 4); v.invalid = d_invalid - (lo >> 5); v.glen = glen; v.cstart = d_cstart; v.cend = d_cend;
        v.n_contigs = n_contigs; v.mstart = d_mstart; v.mend_pm = d_mend_pm; v.n_masks = n_masks;
        return v;
    }
};

struct mg_queries {
    int n = 0, len = 0, n_groups = 0;
    uint8_t* d_codes = nullptr;
    uint32_t* d_tab = nullptr;
    uint32_t* d_tab2 = nullptr;
    uint4* d_packed = nullptr;
    cudaStream_t stream = nullptr;          // the creating stream; the tables are freed in its order
    mutable mg::SeedIndex* sj = nullptr;    // lazily built by launch_seedjoin
    mg::QueriesView view() const {
        mg::QueriesView v;
        v.codes = d_codes; v.tab = d_tab; v.tab2 = d_tab2; v.packed = d_packed; v.n = n; v.len = len; v.n_groups = n_groups;
        return v;
    }
};

namespace mg {
// kernels / launchers implemented across the .cu files
int launch_pack(const char* d_ascii, const std::vector<int64_t>& offsets, mg_genome* g, cudaStream_t s);
int launch_build_queries(const char* d_ascii, mg_queries* q, cudaStream_t s);
int launch_screen_bitsliced(const mg_genome* g, const mg_queries* q, int max_mismatch, int64_t wb, int64_t we,
                            uint32_t* d_counts, cudaStream_t s);
int launch_screen_general(const mg_genome* g, const mg_queries* q, const mg_criteria* crit, const mg_pam* pam,
                          int64_t wb, int64_t we, uint32_t* d_counts, cudaStream_t s);
}  // namespace mg
