// K1: ASCII -> 2-bit + invalid plane packer, genome handle, mask interval table.
// Stands in for the subject side of the reference's blastn call and IndexedFasta (index.py:150-211).
#include <algorithm>
#include <cstdarg>

#include "common.cuh"

extern "C" void mg_genome_free(mg_genome* g);

namespace mg {

static thread_local std::string g_err;
void set_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
}
const char* last_error() { return g_err.c_str(); }

cudaError_t dev_alloc(void** p, size_t bytes, cudaStream_t s) {
    static thread_local int configured_for = -1;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (configured_for != dev) {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            unsigned long long keep = ~0ull;      // never trim: later calls reuse the planes
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        configured_for = dev;
    }
    return cudaMallocAsync(p, bytes, s);
}

void dev_free(void* p, cudaStream_t s) {
    if (p) cudaFreeAsync(p, s);
}

// One thread packs 32 consecutive bases of one contig: two 2-bit words + one invalid word.
// HBM-bound: reads 1 B/base (two 16-B vector loads per thread), writes 0.375 B/base.
__global__ void __launch_bounds__(256)
pack_kernel(const char* __restrict__ ascii, int64_t src_off, int64_t n, int64_t gstart,
            uint32_t* __restrict__ bases, uint32_t* __restrict__ invalid) {
    const int64_t blk = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // 32-base block index in this contig
    const int64_t first = blk * 32;
    if (first >= n) return;
    const char* src = ascii + src_off + first;
    uint32_t lo = 0, hi = 0, inv = 0;
    unsigned char c[32];
    const int64_t left = n - first;
    if (left >= 32 && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
        const uint4 v0 = __ldg(reinterpret_cast<const uint4*>(src));
        const uint4 v1 = __ldg(reinterpret_cast<const uint4*>(src) + 1);
        memcpy(c, &v0, 16);
        memcpy(c + 16, &v1, 16);
    } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) c[i] = (i < left) ? (unsigned char)src[i] : (unsigned char)'N';
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        // A/a=0 C/c=1 G/g=2 T/t=3, anything else invalid.  (ch>>1)&3 maps A->0 C->1 G->3 T->2; fix G/T.
        const unsigned ch = c[i] & 0xDFu;   // fold case
        const bool ok = (ch == 'A') | (ch == 'C') | (ch == 'G') | (ch == 'T');
        unsigned code = (ch >> 1) & 3u;
        code ^= (code >> 1);                // 0,1,3,2 -> 0,1,2,3
        code = ok ? code : 0u;
        if (i < 16) lo |= code << (2 * i); else hi |= code << (2 * (i - 16));
        inv |= (ok ? 0u : 1u) << i;
    }
    const int64_t gb = gstart + first;      // multiple of 32
    bases[(gb >> 4)] = lo;
    bases[(gb >> 4) + 1] = hi;
    invalid[gb >> 5] = inv;
}

__global__ void decode_kernel(GenomeView g, int64_t gfrom, int64_t n, char* out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = base_at(g, gfrom + i);
    out[i] = c == kCodeInvalid ? 'N' : "ACGT"[c];
}

// The part of contig c (contig coordinates, start a multiple of 32) that lies inside the stored range of g.
static bool stored_part(const mg_genome* g, int c, int64_t& a, int64_t& b) {
    a = std::max<int64_t>(0, g->lo - g->cstart[c]);          // lo and cstart are multiples of 32
    b = std::min<int64_t>(g->clen[c], g->hi - g->cstart[c]);
    return a < b;
}

// offsets[c] = position in d_ascii of the first STORED base of contig c (= the contig's first base for a whole genome).
int launch_pack(const char* d_ascii, const std::vector<int64_t>& offsets, mg_genome* g, cudaStream_t s) {
    // padding and tails: bases = 0, invalid = 1 everywhere first
    MG_CUDA(cudaMemsetAsync(g->d_bases, 0, g->bases_words * 4, s));
    MG_CUDA(cudaMemsetAsync(g->d_invalid, 0xFF, g->invalid_words * 4, s));
    const GenomeView v = g->view();
    for (int c = 0; c < g->n_contigs; ++c) {
        int64_t a, b;
        if (!stored_part(g, c, a, b)) continue;
        const int64_t n = b - a;
        const int64_t blocks32 = (n + 31) / 32;
        const unsigned grid = (unsigned)((blocks32 + 255) / 256);
        pack_kernel<<<grid, 256, 0, s>>>(d_ascii, offsets[c], n, g->cstart[c] + a, const_cast<uint32_t*>(v.bases), const_cast<uint32_t*>(v.invalid));
    }
    MG_CUDA(cudaGetLastError());
    return MG_OK;
}

static int64_t layout(const int64_t* offsets, int n_contigs, mg_genome* g) {
    int64_t pos = kPad;
    for (int c = 0; c < n_contigs; ++c) {
        const int64_t n = offsets[c + 1] - offsets[c];
        if (n < 0) return -1 - c;
        pos = (pos + kAlign - 1) / kAlign * kAlign;
        if (g) { g->clen.push_back(n); g->cstart.push_back(pos); g->cend.push_back(pos + n); g->total += n; }
        pos += n + kPad;
    }
    return (pos + 127) / 128 * 128;
}

// chunk_begin < 0: the whole subject.  Otherwise only [chunk_begin - halo, chunk_end + halo) is stored.
static int genome_alloc(const int64_t* offsets, int n_contigs, cudaStream_t s, mg_genome** out,
                        int64_t chunk_begin = -1, int64_t chunk_end = -1, int64_t halo = 0) {
    MG_CHECK_ARG(offsets != nullptr && n_contigs >= 0 && out != nullptr);
    mg_genome* g = new mg_genome();
    struct Guard {               // frees the half-built handle on every early return below
        mg_genome*& g; bool armed = true;
        ~Guard() { if (armed) { mg_genome_free(g); g = nullptr; } }
    } guard{g};
    MG_CUDA(cudaGetDevice(&g->device));
    g->n_contigs = n_contigs;
    g->glen = layout(offsets, n_contigs, g);
    if (g->glen < 0) { set_error("contig %d has negative length", (int)(-1 - g->glen)); return MG_ERR_ARG; }
    g->lo = 0; g->hi = g->glen; g->own_begin = 0; g->own_end = g->glen;
    if (chunk_begin >= 0) {
        if (chunk_end > g->glen) chunk_end = g->glen;
        if (chunk_end < chunk_begin || halo < kMinHalo) { set_error("chunk [%lld, %lld) with halo %lld (>= %d needed)", (long long)chunk_begin, (long long)chunk_end, (long long)halo, kMinHalo); return MG_ERR_ARG; }
        g->is_chunk = true;
        g->halo = halo;
        g->own_begin = chunk_begin; g->own_end = chunk_end;
        g->lo = std::max<int64_t>(0, (chunk_begin - halo) / 128 * 128);
        g->hi = std::min<int64_t>(g->glen, (chunk_end + halo + 127) / 128 * 128);
        if (g->hi < g->lo) g->hi = g->lo;
    }
    g->bases_words = (size_t)((g->hi - g->lo) / 16) + 8;
    g->invalid_words = (size_t)((g->hi - g->lo) / 32) + 8;
    MG_CUDA(dev_alloc((void**)&g->d_bases, g->bases_words * 4, s));
    MG_CUDA(dev_alloc((void**)&g->d_invalid, g->invalid_words * 4, s));
    const size_t nb = sizeof(int64_t) * (size_t)std::max(1, n_contigs);
    g->last_stream = s;
    MG_CUDA(dev_alloc((void**)&g->d_cstart, nb, s));
    MG_CUDA(dev_alloc((void**)&g->d_cend, nb, s));
    if (n_contigs) {
        // on the caller's stream (a plain cudaMemcpy would join the legacy default stream and wait for a screen running there);
        // the host vectors live in the handle and every creator synchronises the stream before returning
        MG_CUDA(cudaMemcpyAsync(g->d_cstart, g->cstart.data(), sizeof(int64_t) * n_contigs, cudaMemcpyHostToDevice, s));
        MG_CUDA(cudaMemcpyAsync(g->d_cend, g->cend.data(), sizeof(int64_t) * n_contigs, cudaMemcpyHostToDevice, s));
    }
    guard.armed = false;
    *out = g;
    return MG_OK;
}

}  // namespace mg

using namespace mg;

extern "C" {

const char* mg_last_error(void) { return mg::last_error(); }
int mg_version(void) { return 100; }

int mg_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int mg_init(int device, char* name_out, int name_cap, int* sm_count_out) {
    MG_CUDA(cudaSetDevice(device));
    cudaDeviceProp p;
    MG_CUDA(cudaGetDeviceProperties(&p, device));
    if (p.major < 10) {
        set_error("device %d is sm_%d%d; this library is built for sm_100a only", device, p.major, p.minor);
        return MG_ERR_UNSUPPORTED;
    }
    if (name_out && name_cap > 0) { strncpy(name_out, p.name, name_cap - 1); name_out[name_cap - 1] = 0; }
    if (sm_count_out) *sm_count_out = p.multiProcessorCount;
    return MG_OK;
}

int mg_genome_from_device_ascii(const char* d_seq, const int64_t* offsets, int n_contigs, void* stream, mg_genome** out) {
    mg_genome* g = nullptr;
    int rc = genome_alloc(offsets, n_contigs, (cudaStream_t)stream, &g);
    if (rc) return rc;
    std::vector<int64_t> off(offsets, offsets + n_contigs + 1);
    rc = launch_pack(d_seq, off, g, (cudaStream_t)stream);
    if (rc) { mg_genome_free(g); return rc; }
    *out = g;
    return MG_OK;
}

int mg_genome_from_ascii(const char* seq, const int64_t* offsets, int n_contigs, void* stream, mg_genome** out) {
    MG_CHECK_ARG(offsets != nullptr && n_contigs >= 0);
    const int64_t total = offsets[n_contigs] - offsets[0];
    MG_CHECK_ARG(total == 0 || seq != nullptr);
    cudaStream_t s = (cudaStream_t)stream;
    char* d_seq = nullptr;
    MG_CUDA(dev_alloc((void**)&d_seq, (size_t)std::max<int64_t>(total, 16) + 32, s));
    std::vector<int64_t> rel(n_contigs + 1);
    for (int i = 0; i <= n_contigs; ++i) rel[i] = offsets[i] - offsets[0];
    if (total) {
        cudaError_t e = cudaMemcpyAsync(d_seq, seq + offsets[0], (size_t)total, cudaMemcpyHostToDevice, s);
        if (e != cudaSuccess) { dev_free(d_seq, s); set_error("H2D copy of genome failed: %s", cudaGetErrorString(e)); return MG_ERR_CUDA; }
    }
    int rc = mg_genome_from_device_ascii(d_seq, rel.data(), n_contigs, stream, out);
    dev_free(d_seq, s);      // stream-ordered: released once the pack kernels are done
    cudaStreamSynchronize(s);   // the caller may reuse `seq` as soon as we return
    return rc;
}

int64_t mg_layout_global_length(const int64_t* offsets, int n_contigs) {
    if (offsets == nullptr || n_contigs < 0) return -1;
    const int64_t r = layout(offsets, n_contigs, nullptr);
    return r < 0 ? -1 : r;
}

int mg_genome_from_device_ascii_chunk(const char* d_seq, const int64_t* offsets, int n_contigs, int64_t chunk_begin,
                                      int64_t chunk_end, int64_t halo, void* stream, mg_genome** out) {
    MG_CHECK_ARG(chunk_begin >= 0);
    mg_genome* g = nullptr;
    int rc = genome_alloc(offsets, n_contigs, (cudaStream_t)stream, &g, chunk_begin, chunk_end, halo);
    if (rc) return rc;
    std::vector<int64_t> off(n_contigs + 1, 0);
    for (int c = 0; c < n_contigs; ++c) {
        int64_t a, b;
        off[c] = offsets[c] + (stored_part(g, c, a, b) ? a : 0);
    }
    rc = launch_pack(d_seq, off, g, (cudaStream_t)stream);
    if (rc) { mg_genome_free(g); return rc; }
    *out = g;
    return MG_OK;
}

int mg_genome_from_ascii_chunk(const char* seq, const int64_t* offsets, int n_contigs, int64_t chunk_begin,
                               int64_t chunk_end, int64_t halo, void* stream, mg_genome** out) {
    MG_CHECK_ARG(offsets != nullptr && n_contigs >= 0 && chunk_begin >= 0 && (seq != nullptr || offsets[n_contigs] == offsets[0]));
    cudaStream_t s = (cudaStream_t)stream;
    mg_genome* g = nullptr;
    int rc = genome_alloc(offsets, n_contigs, s, &g, chunk_begin, chunk_end, halo);
    if (rc) return rc;
    // only the stored part of each contig crosses PCIe: pieces are laid out back to back (16-byte aligned starts)
    std::vector<int64_t> off(n_contigs + 1, 0);
    int64_t need = 0;
    for (int c = 0; c < n_contigs; ++c) {
        int64_t a, b;
        off[c] = need;
        if (stored_part(g, c, a, b)) need += (b - a + 15) / 16 * 16;
    }
    char* d_seq = nullptr;
    cudaError_t e = dev_alloc((void**)&d_seq, (size_t)std::max<int64_t>(need, 16) + 32, s);
    if (e != cudaSuccess) { mg_genome_free(g); set_error("device buffer for the chunk: %s", cudaGetErrorString(e)); return MG_ERR_NOMEM; }
    for (int c = 0; c < n_contigs && e == cudaSuccess; ++c) {
        int64_t a, b;
        if (stored_part(g, c, a, b)) e = cudaMemcpyAsync(d_seq + off[c], seq + offsets[c] + a, (size_t)(b - a), cudaMemcpyHostToDevice, s);
    }
    if (e != cudaSuccess) { dev_free(d_seq, s); mg_genome_free(g); set_error("H2D copy of the chunk failed: %s", cudaGetErrorString(e)); return MG_ERR_CUDA; }
    rc = launch_pack(d_seq, off, g, s);
    dev_free(d_seq, s);
    cudaStreamSynchronize(s);
    if (rc) { mg_genome_free(g); return rc; }
    *out = g;
    return MG_OK;
}

int mg_genome_owned_range(const mg_genome* g, int64_t* begin, int64_t* end, int64_t* stored_begin, int64_t* stored_end) {
    MG_CHECK_ARG(g != nullptr);
    if (begin) *begin = g->own_begin;
    if (end) *end = g->own_end;
    if (stored_begin) *stored_begin = g->lo;
    if (stored_end) *stored_end = g->hi;
    return MG_OK;
}

int mg_genome_n_contigs(const mg_genome* g) { return g ? g->n_contigs : 0; }
int64_t mg_genome_contig_length(const mg_genome* g, int c) { return (g && c >= 0 && c < g->n_contigs) ? g->clen[c] : -1; }
int64_t mg_genome_total_bases(const mg_genome* g) { return g ? g->total : 0; }
int64_t mg_genome_device_bytes(const mg_genome* g) { return g ? (int64_t)((g->bases_words + g->invalid_words) * 4) : 0; }
int64_t mg_genome_contig_global_start(const mg_genome* g, int c) { return (g && c >= 0 && c < g->n_contigs) ? g->cstart[c] : -1; }
int64_t mg_genome_global_length(const mg_genome* g) { return g ? g->glen : 0; }

int mg_genome_decode(const mg_genome* g, int contig, int64_t start, int64_t end, char* out) {
    MG_CHECK_ARG(g && contig >= 0 && contig < g->n_contigs && start >= 0 && end >= start && end <= g->clen[contig] && out);
    const int64_t n = end - start;
    if (n && (g->cstart[contig] + start < g->lo || g->cstart[contig] + end > g->hi)) { set_error("decode range lies outside the stored chunk"); return MG_ERR_ARG; }
    if (n == 0) return MG_OK;
    char* d = nullptr;
    MG_CUDA(cudaMalloc(&d, (size_t)n));
    decode_kernel<<<(unsigned)((n + 255) / 256), 256>>>(g->view(), g->cstart[contig] + start, n, d);
    cudaError_t e = cudaMemcpy(out, d, (size_t)n, cudaMemcpyDeviceToHost);
    cudaFree(d);
    MG_CUDA(e);
    return MG_OK;
}

int mg_genome_set_masks(mg_genome* g, const int32_t* contig, const int64_t* start, const int64_t* end, int n, void* stream) {
    MG_CHECK_ARG(g && n >= 0 && (n == 0 || (contig && start && end)));
    std::vector<std::pair<int64_t, int64_t>> iv;
    for (int i = 0; i < n; ++i) {
        MG_CHECK_ARG(contig[i] >= 0 && contig[i] < g->n_contigs);
        const int64_t a = std::min(start[i], end[i]), b = std::max(start[i], end[i]);   // functions.py:323-325 uses min/max
        iv.emplace_back(g->cstart[contig[i]] + a, g->cstart[contig[i]] + b);
    }
    std::sort(iv.begin(), iv.end());
    std::vector<int64_t> ms(iv.size()), me(iv.size());
    int64_t run = INT64_MIN;
    for (size_t i = 0; i < iv.size(); ++i) { ms[i] = iv[i].first; run = std::max(run, iv[i].second); me[i] = run; }
    // the old tables may still be read by a screen in flight on the stream that used them last: free in its order
    if (g->d_mstart) { dev_free(g->d_mstart, g->last_stream); g->d_mstart = nullptr; }
    if (g->d_mend_pm) { dev_free(g->d_mend_pm, g->last_stream); g->d_mend_pm = nullptr; }
    g->n_masks = (int)iv.size();
    if (g->n_masks) {
        MG_CUDA(dev_alloc((void**)&g->d_mstart, sizeof(int64_t) * iv.size(), (cudaStream_t)stream));
        MG_CUDA(dev_alloc((void**)&g->d_mend_pm, sizeof(int64_t) * iv.size(), (cudaStream_t)stream));
        MG_CUDA(cudaMemcpyAsync(g->d_mstart, ms.data(), sizeof(int64_t) * iv.size(), cudaMemcpyHostToDevice, (cudaStream_t)stream));
        MG_CUDA(cudaMemcpyAsync(g->d_mend_pm, me.data(), sizeof(int64_t) * iv.size(), cudaMemcpyHostToDevice, (cudaStream_t)stream));
        MG_CUDA(cudaStreamSynchronize((cudaStream_t)stream));   // ms/me are stack-owned
    }
    return MG_OK;
}

void mg_genome_free(mg_genome* g) {
    if (!g) return;
    // stream-ordered behind the last screen that read the planes (a caller's non-blocking stream is not ordered
    // against the legacy default stream)
    dev_free(g->d_bases, g->last_stream); dev_free(g->d_invalid, g->last_stream);
    dev_free(g->d_cstart, g->last_stream); dev_free(g->d_cend, g->last_stream);
    dev_free(g->d_mstart, g->last_stream); dev_free(g->d_mend_pm, g->last_stream);
    delete g;
}

}  // extern "C"
