// C-ABI glue: screen dispatch, host-buffer convenience path, integer-pipe microbenchmarks.
#include <ctime>

#include <algorithm>
#include <string>
#include <vector>
#include "common.cuh"

using namespace mg;

namespace mg {

int screen_stats(unsigned long long out[2]);

// ---------------------------------------------------------------- integer-pipe microbenchmarks
// Denominators for the scan kernel's roofline (MEASURED_PEAKS.json only has HBM and bf16 numbers).
__global__ void __launch_bounds__(1024, 1) ubench_lop3(uint32_t* out, int iters) {
    uint32_t a = threadIdx.x * 2654435761u, b = blockIdx.x * 40503u + 1u, c = a ^ 0x9E3779B9u, d = b + 77u;
    uint32_t e = a + 1, f = b + 2, g = c + 3, h = d + 4;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {   // 8 independent LOP3 chains x 8 = 64 LOP3 per iteration
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(b), "r"(c));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(b) : "r"(c), "r"(d));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(c) : "r"(d), "r"(e));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(d) : "r"(e), "r"(f));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(e) : "r"(f), "r"(g));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(f) : "r"(g), "r"(h));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(g) : "r"(h), "r"(a));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(h) : "r"(a), "r"(b));
        }
    }
    if ((a ^ b ^ c ^ d ^ e ^ f ^ g ^ h) == 0x12345u) out[0] = a;
}

// IMAD (FMA pipe; the gapped rows kernel moves 4 of its 7 word operations per DP cell there) alone (MIX = 0) and
// interleaved 1:1 with LOP3 (MIX = 1: the two pipes together = the issue-slot ceiling of the rows kernel).
template <int MIX>
__global__ void __launch_bounds__(1024, 1) ubench_imad(uint32_t* out, int iters, uint32_t one) {
    uint32_t a = threadIdx.x * 2654435761u, b = blockIdx.x * 40503u + 1u, c = a ^ 0x9E3779B9u, d = b + 77u;
    uint32_t e = a + 1, f = b + 2, g = c + 3, h = d + 4;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a) : "r"(one), "r"(b));
            if (MIX) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(e) : "r"(f), "r"(g));
            asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b) : "r"(one), "r"(c));
            if (MIX) asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(f) : "r"(g), "r"(h));
            asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(c) : "r"(one), "r"(d));
            if (MIX) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(g) : "r"(h), "r"(e));
            asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(d) : "r"(one), "r"(a));
            if (MIX) asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(h) : "r"(e), "r"(f));
        }
    }
    if ((a ^ b ^ c ^ d ^ e ^ f ^ g ^ h) == 0x12345u) out[0] = a;
}

__global__ void __launch_bounds__(1024, 1) ubench_lds(uint32_t* out, int iters) {
    __shared__ uint32_t sm[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i * 2654435761u;
    __syncthreads();
    uint32_t acc = 0;
    uint32_t idx = (threadIdx.x & 3u) + 4u * (threadIdx.x >> 5);    // 4 consecutive banks per warp, like the scan
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 32; ++u) {
            uint32_t v;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"((uint32_t)__cvta_generic_to_shared(sm + ((idx + 4u * u) & 4095u))));
            acc ^= v;
        }
        idx += 128u;
    }
    if (acc == 0x12345u) out[0] = acc;
}

// Same access pattern as the scan (every lane of a warp reads one of 4 adjacent table entries), with 8- and
// 16-byte entries: does a broadcast-heavy wide load cost one wavefront or several?
template <int WORDS>
__global__ void __launch_bounds__(1024, 1) ubench_lds_wide(uint32_t* out, int iters) {
    __shared__ __align__(16) uint32_t sm[8192];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = i * 2654435761u;
    __syncthreads();
    uint32_t acc = 0;
    uint32_t idx = ((threadIdx.x * 7u) & 3u) + 4u * (threadIdx.x >> 5);   // entry index: 4 adjacent entries per warp
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            const uint32_t e = (idx + 4u * u) & (8192u / WORDS - 1u);
            const uint32_t addr = (uint32_t)__cvta_generic_to_shared(sm + e * WORDS);
            if (WORDS == 2) {
                uint32_t a, b;
                asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "r"(addr));
                acc ^= a ^ b;
            } else {
                uint32_t a, b, c, d;
                asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(addr));
                acc ^= a ^ b ^ c ^ d;
            }
        }
        idx += 64u;
    }
    if (acc == 0x12345u) out[0] = acc;
}

// Indexed constant-bank loads with the scan's pattern (4 distinct adjacent words per warp): a pipe of its own?
__constant__ uint32_t c_probe[8192];
__global__ void __launch_bounds__(1024, 1) ubench_ldc(uint32_t* out, int iters) {
    uint32_t acc = 0;
    uint32_t idx = ((threadIdx.x * 7u) & 3u) + 4u * (threadIdx.x >> 5);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) acc ^= c_probe[(idx + 4u * u) & 8191u];
        idx += 64u;
    }
    if (acc == 0x12345u) out[0] = acc;
}

// Warp shuffles as a lookup path: alone (MIX = 0) and interleaved 1:1 with shared-memory loads (MIX = 1) - do the two
// share one pipe?
template <int MIX>
__global__ void __launch_bounds__(1024, 1) ubench_shfl(uint32_t* out, int iters) {
    __shared__ uint32_t sm[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i * 2654435761u;
    __syncthreads();
    uint32_t acc = 0, v = threadIdx.x * 40503u;
    uint32_t idx = (threadIdx.x & 3u) + 4u * (threadIdx.x >> 5);
    int src = (threadIdx.x * 7) & 31;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            acc ^= __shfl_sync(0xFFFFFFFFu, v + u, (src + u) & 31);
            if (MIX) {
                uint32_t w;
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w) : "r"((uint32_t)__cvta_generic_to_shared(sm + ((idx + 4u * u) & 4095u))));
                acc ^= w;
            }
        }
        idx += 64u;
        v += acc;
    }
    if (acc == 0x12345u) out[0] = acc;
}

}  // namespace mg

namespace mg {
struct GeneralCtx;
int seedjoin_stats(uint64_t* out);
bool seedjoin_eligible(const mg_queries* q, const mg_criteria* crit);
int launch_seedjoin(const mg_genome* g, const mg_queries* q, const mg_criteria* crit, const GeneralCtx* d_ctx, int64_t wb, int64_t we,
                    uint32_t* d_counts, cudaStream_t s, std::vector<std::pair<int64_t, int64_t>>& dirty_ranges);
}

extern "C" {

int mg_last_screen_stats(uint64_t* out) {
    MG_CHECK_ARG(out != nullptr);
    unsigned long long tmp[2];
    const int rc = screen_stats(tmp);
    out[0] = tmp[0]; out[1] = tmp[1];
    return rc;
}

int mg_last_seedjoin_stats(uint64_t* out) {
    MG_CHECK_ARG(out != nullptr);
    return mg::seedjoin_stats(out);
}

int mg_microbench(int which, double* gops, void* stream) {
    MG_CHECK_ARG(gops != nullptr && which >= 0 && which <= 8);
    cudaStream_t s = (cudaStream_t)stream;
    int dev = 0, sms = 0;
    MG_CUDA(cudaGetDevice(&dev));
    MG_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    uint32_t* d = nullptr;
    MG_CUDA(cudaMalloc(&d, 64));
    cudaEvent_t e0, e1;
    MG_CUDA(cudaEventCreate(&e0));
    MG_CUDA(cudaEventCreate(&e1));
    const int iters = 4096;
    double best = 0;
    for (int rep = 0; rep < 4; ++rep) {
        MG_CUDA(cudaEventRecord(e0, s));
        if (which == 0) ubench_lop3<<<sms, 1024, 0, s>>>(d, iters);
        else if (which == 1) ubench_lds<<<sms, 1024, 0, s>>>(d, iters);
        else if (which == 2) ubench_lds_wide<2><<<sms, 1024, 0, s>>>(d, iters);
        else if (which == 3) ubench_lds_wide<4><<<sms, 1024, 0, s>>>(d, iters);
        else if (which == 4) ubench_ldc<<<sms, 1024, 0, s>>>(d, iters);
        else if (which == 5) ubench_shfl<0><<<sms, 1024, 0, s>>>(d, iters);
        else if (which == 6) ubench_shfl<1><<<sms, 1024, 0, s>>>(d, iters);
        else if (which == 7) ubench_imad<0><<<sms, 1024, 0, s>>>(d, iters, 1u);
        else ubench_imad<1><<<sms, 1024, 0, s>>>(d, iters, 1u);
        MG_CUDA(cudaEventRecord(e1, s));
        MG_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        MG_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        const double ops = (double)sms * 1024.0 * iters * (which == 0 ? 64.0 : which == 1 ? 32.0 : which == 6 ? 32.0 : which == 7 ? 32.0 : which == 8 ? 64.0 : 16.0);   // INSTRUCTIONS per lane and iteration
        const double r = ops / (ms * 1e-3) / 1e9;
        if (rep > 0 && r > best) best = r;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
    *gops = best;
    return MG_OK;
}

int mg_screen(const mg_genome* g, const mg_queries* q, const mg_criteria* crit, const mg_pam* pam,
              int64_t win_begin, int64_t win_end, uint32_t* d_counts, void* stream) {
    MG_CHECK_ARG(g && q && crit && d_counts);
    MG_CHECK_ARG(crit->max_mismatch >= 0 && crit->max_gap >= 0 && crit->max_gap <= MG_MAX_GAPS);
    MG_CHECK_ARG(crit->pamless || pam != nullptr);
    MG_CHECK_ARG(crit->n_units >= 0 && crit->n_units <= MG_MAX_PATTERN_UNITS && crit->n_ops >= 0 && crit->n_ops <= MG_MAX_PATTERN_OPS);
    const int L = q->len;
    // every window start that can hold an alignment: from the first base of padding before contig 0
    const int64_t last = g->glen - L;            // buffers carry spare words beyond glen
    if (win_begin < 0) win_begin = 0;
    if (win_end < 0 || win_end > last + 1) win_end = last + 1;
    if (g->is_chunk) {                           // a chunk only owns (and only stores the bases of) its own window range
        win_begin = std::max(win_begin, g->own_begin);
        win_end = std::min(win_end, g->own_end);
    }
    if (win_end <= win_begin || q->n == 0) return MG_OK;
    cudaStream_t s = (cudaStream_t)stream;
    g->last_stream = s;
    const bool fast = crit->n_units == 0 && crit->max_gap == 0 && crit->pamless;
    if (fast) {
        if (crit->max_mismatch >= L) { set_error("mismatch budget %d >= gRNA length %d", crit->max_mismatch, L); return MG_ERR_ARG; }
        if (mg::seedjoin_eligible(q, crit)) {
            // a large gRNA set: the seed-and-join screen counts every window whose surroundings are clean and away from masks;
            // the pair-word scan takes the rest as window ranges (same counts either way)
            std::vector<std::pair<int64_t, int64_t>> dirty;
            const int r2 = mg::launch_seedjoin(g, q, crit, nullptr, win_begin, win_end, d_counts, s, dirty);
            if (r2 == MG_OK) {
                for (const auto& r : dirty) {
                    const int r3 = launch_screen_bitsliced(g, q, crit->max_mismatch, r.first, r.second, d_counts, s);
                    if (r3 != MG_OK) return r3;
                }
                return MG_OK;
            }
            if (r2 != MG_ERR_UNSUPPORTED) return r2;
        }
        return launch_screen_bitsliced(g, q, crit->max_mismatch, win_begin, win_end, d_counts, s);
    }
    return launch_screen_general(g, q, crit, pam, win_begin, win_end, d_counts, s);
}

int mg_screen_to_host(const mg_genome* g, const mg_queries* q, const mg_criteria* crit, const mg_pam* pam,
                      int64_t win_begin, int64_t win_end, uint32_t* counts_out, void* stream) {
    MG_CHECK_ARG(g && q && crit);
    const size_t n2 = 2 * (size_t)q->n;
    if (n2 == 0) return MG_OK;
    MG_CHECK_ARG(counts_out != nullptr);
    cudaStream_t s = (cudaStream_t)stream;
    uint32_t* d_counts = nullptr;
    MG_CUDA(dev_alloc((void**)&d_counts, sizeof(uint32_t) * n2, s));
    cudaError_t e = cudaMemsetAsync(d_counts, 0, sizeof(uint32_t) * n2, s);
    int rc = e == cudaSuccess ? mg_screen(g, q, crit, pam, win_begin, win_end, d_counts, stream) : MG_ERR_CUDA;
    if (rc == MG_OK) {
        e = cudaMemcpyAsync(counts_out, d_counts, sizeof(uint32_t) * n2, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) { set_error("screen failed: %s", cudaGetErrorString(e)); rc = MG_ERR_CUDA; }
    } else if (e != cudaSuccess) {
        set_error("clearing the counts failed: %s", cudaGetErrorString(e));
    }
    dev_free(d_counts, s);
    return rc;
}

int mg_screen_host(const char* seq, const int64_t* offsets, int n_contigs,
                   const int32_t* mask_contig, const int64_t* mask_start, const int64_t* mask_end, int n_masks,
                   const char* grna, int n, int len, const mg_criteria* crit, const mg_pam* pam,
                   uint32_t* counts_out, void* stream) {
    MG_CHECK_ARG(counts_out != nullptr || n == 0);
    MG_CHECK_ARG(offsets != nullptr && n_contigs >= 0 && n_masks >= 0);
    cudaStream_t s = (cudaStream_t)stream;
    uint32_t* d_counts = nullptr;
    const bool trace = getenv("MG_TRACE") != nullptr;
    auto now = [] { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; };
    const double t0 = now();
    // Per-thread cache across calls: a pipeline screens the same gRNA set against one subject after another
    // (MINORg.filter_background: reference, query and background files), so the query tables, the copy stream and the
    // count scratch of the previous call are kept; the gRNA bytes are compared, not trusted by pointer.
    struct HostCache {
        int device = -1, n = 0, len = 0;
        std::string blob;
        mg_queries* q = nullptr;
        cudaStream_t cs = nullptr;
        ~HostCache() {}                      // released with the context at process exit
    };
    static thread_local HostCache cache;
    int dev = 0;
    MG_CUDA(cudaGetDevice(&dev));
    int rc = MG_OK;
    const size_t gbytes = (size_t)std::max(n, 0) * (size_t)std::max(len, 0);
    const bool hit = cache.q != nullptr && cache.device == dev && cache.n == n && cache.len == len &&
                     cache.blob.size() == gbytes && (gbytes == 0 || memcmp(cache.blob.data(), grna, gbytes) == 0);
    if (!hit) {
        if (cache.q) { mg_queries_free(cache.q); cache.q = nullptr; }
        rc = mg_queries_create(grna, n, len, stream, &cache.q);
        if (rc == MG_OK) { cache.device = dev; cache.n = n; cache.len = len; cache.blob.assign(grna ? grna : "", gbytes); }
        else cache.q = nullptr;
    }
    mg_queries* q = cache.q;
    if (rc == MG_OK && n > 0) {
        if (dev_alloc((void**)&d_counts, sizeof(uint32_t) * 2 * (size_t)n, s) != cudaSuccess) { set_error("allocating the counts failed"); rc = MG_ERR_NOMEM; }
        else cudaMemsetAsync(d_counts, 0, sizeof(uint32_t) * 2 * (size_t)n, s);
    }
    // Contigs are independent (no window spans two), so the subject goes through in batches of whole contigs: batch i+1
    // is copied and packed on a second stream while batch i is screened on the caller's - the H2D copy of a large genome
    // hides behind the screen.  Counts accumulate in d_counts.
    const int64_t total = n_contigs ? offsets[n_contigs] - offsets[0] : 0;
    const int64_t batch_bases = std::max<int64_t>(total / 8, (int64_t)32 << 20);
    std::vector<mg_genome*> batches;
    if (rc == MG_OK && cache.cs == nullptr && cudaStreamCreateWithFlags(&cache.cs, cudaStreamNonBlocking) != cudaSuccess) { cache.cs = nullptr; set_error("cudaStreamCreate failed"); rc = MG_ERR_CUDA; }
    cudaStream_t cs = cache.cs;
    double t_copy = 0.0;
    for (int c0 = 0; rc == MG_OK && (c0 < n_contigs || (n_contigs == 0 && batches.empty())); ) {
        int c1 = c0;
        while (c1 < n_contigs && (c1 == c0 || offsets[c1 + 1] - offsets[c0] <= batch_bases)) ++c1;
        const double tc = now();
        mg_genome* g = nullptr;
        rc = mg_genome_from_ascii(seq, offsets + c0, c1 - c0, cs, &g);          // returns when this batch is resident
        if (rc != MG_OK) break;
        batches.push_back(g);
        std::vector<int32_t> mc;
        std::vector<int64_t> ms, me;
        for (int i = 0; i < n_masks; ++i)
            if (mask_contig[i] >= c0 && mask_contig[i] < c1) { mc.push_back(mask_contig[i] - c0); ms.push_back(mask_start[i]); me.push_back(mask_end[i]); }
        rc = mg_genome_set_masks(g, mc.data(), ms.data(), me.data(), (int)mc.size(), cs);
        if (rc == MG_OK && cudaStreamSynchronize(cs) != cudaSuccess) { set_error("genome batch failed: %s", cudaGetErrorString(cudaGetLastError())); rc = MG_ERR_CUDA; }
        t_copy += now() - tc;
        if (rc == MG_OK && n > 0) rc = mg_screen(g, q, crit, pam, 0, -1, d_counts, stream);
        c0 = c1;
        if (n_contigs == 0) break;
    }
    if (rc == MG_OK && n > 0) {
        std::vector<uint32_t> tmp(2 * (size_t)n);
        cudaError_t e = cudaMemcpyAsync(tmp.data(), d_counts, sizeof(uint32_t) * tmp.size(), cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) { set_error("screen failed: %s", cudaGetErrorString(e)); rc = MG_ERR_CUDA; }
        else for (int i = 0; i < n; ++i) counts_out[i] = tmp[i] + tmp[(size_t)n + i];
    } else {
        cudaStreamSynchronize(s);
    }
    const double t3 = now();
    dev_free(d_counts, s);
    for (mg_genome* g : batches) mg_genome_free(g);
    if (trace) fprintf(stderr, "[mg_screen_host] %d batches; host time in H2D+pack+masks %.1f ms (overlaps the screen), total %.1f ms, free %.1f ms\n",
                       (int)batches.size(), t_copy, t3 - t0, now() - t3);
    return rc;
}

}  // extern "C"
