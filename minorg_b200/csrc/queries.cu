// gRNA set resident in HBM: base codes of both orientations + the bit-sliced mismatch tables of the
// scan kernel.  Replaces the gRNA FASTA the reference hands to blastn (MINORg.py:2134-2135).
#include "common.cuh"

namespace mg {

__device__ __forceinline__ int ascii_code(unsigned char c) {
    const unsigned ch = c & 0xDFu;
    const bool ok = (ch == 'A') | (ch == 'C') | (ch == 'G') | (ch == 'T');
    unsigned code = (ch >> 1) & 3u;
    code ^= (code >> 1);
    return ok ? (int)code : kCodeInvalid;
}

// codes[q][j]: q < N forward gRNA, q >= N reverse complement of gRNA q-N.
__global__ void query_codes_kernel(const char* __restrict__ ascii, int n, int len, uint8_t* __restrict__ codes) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= 2 * n * len) return;
    const int q = idx / len, j = idx % len;
    int c;
    if (q < n) {
        c = ascii_code((unsigned char)ascii[(size_t)q * len + j]);
    } else {
        c = ascii_code((unsigned char)ascii[(size_t)(q - n) * len + (len - 1 - j)]);
        if (c != kCodeInvalid) c = 3 - c;
    }
    codes[idx] = (uint8_t)c;
}

// tab[g][4j+b] bit i = 1 iff query 32g+i does NOT have base b at position j (or is padding);
// tab[g][4L] = all ones (looked up for an invalid genome base).
__global__ void query_table_kernel(const uint8_t* __restrict__ codes, int nq, int len, int n_groups,
                                   uint32_t* __restrict__ tab) {
    const int stride = 4 * len + 1;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_groups * stride) return;
    const int g = idx / stride, slot = idx % stride;
    uint32_t w = 0xFFFFFFFFu;
    if (slot < 4 * len) {
        const int j = slot >> 2, b = slot & 3;
        w = 0;
        for (int i = 0; i < 32; ++i) {
            const int q = g * 32 + i;
            const bool mism = (q >= nq) || (codes[(size_t)q * len + j] != b);
            w |= (mism ? 1u : 0u) << i;
        }
    }
    tab[idx] = w;
}

// tab2[g][16k + b0 + 4*b1] bit i = 1 iff query 32g+i mismatches base b0 at position 2k OR base b1 at position 2k+1
// (a position beyond the gRNA never mismatches); tab2[g][16*NP] = all ones (invalid base in the pair).
__global__ void query_pair_table_kernel(const uint32_t* __restrict__ tab, int len, int n_groups, uint32_t* __restrict__ tab2) {
    const int np = (len + 1) / 2, stride = 16 * np + 1, s1 = 4 * len + 1;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_groups * stride) return;
    const int g = idx / stride, slot = idx % stride;
    uint32_t w = 0xFFFFFFFFu;
    if (slot < 16 * np) {
        const int k = slot >> 4, b0 = slot & 3, b1 = (slot >> 2) & 3;
        w = tab[(size_t)g * s1 + 4 * (2 * k) + b0];
        if (2 * k + 1 < len) w |= tab[(size_t)g * s1 + 4 * (2 * k + 1) + b1];
    }
    tab2[idx] = w;
}

// packed[q] = 2-bit codes (position j at bits 2j of the 64-bit x:y) and the mask of non-ACGT positions
__global__ void query_pack_kernel(const uint8_t* __restrict__ codes, int nq, int len, uint4* __restrict__ packed) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    unsigned long long b = 0;
    uint32_t iv = 0;
    for (int j = 0; j < len; ++j) {
        const int c = codes[(size_t)q * len + j];
        if (c >= 4) iv |= 1u << j; else b |= (unsigned long long)c << (2 * j);
    }
    packed[q] = make_uint4((uint32_t)b, (uint32_t)(b >> 32), iv, 0u);
}

int launch_build_queries(const char* d_ascii, mg_queries* q, cudaStream_t s) {
    const int total = 2 * q->n * q->len;
    if (total > 0) query_codes_kernel<<<(total + 255) / 256, 256, 0, s>>>(d_ascii, q->n, q->len, q->d_codes);
    const int words = q->n_groups * (4 * q->len + 1);
    if (words > 0) query_table_kernel<<<(words + 255) / 256, 256, 0, s>>>(q->d_codes, 2 * q->n, q->len, q->n_groups, q->d_tab);
    const int words2 = q->n_groups * (16 * ((q->len + 1) / 2) + 1);
    if (words2 > 0) query_pair_table_kernel<<<(words2 + 255) / 256, 256, 0, s>>>(q->d_tab, q->len, q->n_groups, q->d_tab2);
    if (q->n > 0) query_pack_kernel<<<(2 * q->n + 255) / 256, 256, 0, s>>>(q->d_codes, 2 * q->n, q->len, q->d_packed);
    MG_CUDA(cudaGetLastError());
    return MG_OK;
}

}  // namespace mg

using namespace mg;

extern "C" {

void mg_queries_free(mg_queries* q);

int mg_queries_create(const char* grna, int n, int len, void* stream, mg_queries** out) {
    MG_CHECK_ARG(out != nullptr && n >= 0 && len >= 1 && len <= MG_MAX_GRNA_LEN && (n == 0 || grna != nullptr));
    MG_CHECK_ARG((int64_t)n * 2 * len < (int64_t)1 << 31);
    cudaStream_t s = (cudaStream_t)stream;
    mg_queries* q = new mg_queries();
    char* d_ascii = nullptr;
    struct Guard {               // frees the half-built handle and the staging buffer on every early return below
        mg_queries*& q; char*& a; bool armed = true;
        cudaStream_t s;
        ~Guard() { dev_free(a, s); if (armed) { mg_queries_free(q); q = nullptr; } }
    } guard{q, d_ascii, true, s};
    q->n = n; q->len = len; q->n_groups = (2 * n + 31) / 32;
    q->stream = s;
    const size_t nbytes = (size_t)n * len;
    // stream-ordered pool allocations: a cudaMalloc/cudaFree pair costs 0.1-0.3 ms each and synchronises the device
    MG_CUDA(dev_alloc((void**)&d_ascii, nbytes + 16, s));
    MG_CUDA(dev_alloc((void**)&q->d_codes, 2 * nbytes + 16, s));
    MG_CUDA(dev_alloc((void**)&q->d_tab, ((size_t)q->n_groups * (4 * len + 1) + 4) * 4, s));
    MG_CUDA(dev_alloc((void**)&q->d_tab2, ((size_t)q->n_groups * (16 * ((len + 1) / 2) + 1) + 4) * 4, s));
    MG_CUDA(dev_alloc((void**)&q->d_packed, sizeof(uint4) * ((size_t)2 * n + 1), s));
    if (nbytes) MG_CUDA(cudaMemcpyAsync(d_ascii, grna, nbytes, cudaMemcpyHostToDevice, s));
    int rc = launch_build_queries(d_ascii, q, s);
    MG_CUDA(cudaStreamSynchronize(s));
    if (rc) return rc;
    guard.armed = false;
    *out = q;
    return MG_OK;
}

int mg_queries_count(const mg_queries* q) { return q ? q->n : 0; }
int mg_queries_length(const mg_queries* q) { return q ? q->len : 0; }

void mg_queries_free(mg_queries* q) {
    if (!q) return;
    mg::seed_index_free(q->sj);
    dev_free(q->d_codes, q->stream); dev_free(q->d_tab, q->stream); dev_free(q->d_tab2, q->stream); dev_free(q->d_packed, q->stream);
    delete q;
}

}  // extern "C"
