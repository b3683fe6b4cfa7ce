// K2: PAM-site enumeration over a target record, both strands.
// Replaces the regex.finditer(pattern, seq, overlapped=True) loops of find_multi_gRNA
// (generate_grna.py:38-41) with pattern = (?<=five).{L}(?=three), IGNORECASE (pam.py:103-125).
#include <algorithm>

#include "common.cuh"

namespace mg {

struct PamDev {            // device copy of one side, allow sets as 256-bit ASCII classes
    int n_alts;
    int len[MG_MAX_PAM_ALTS];
    uint32_t allow[MG_MAX_PAM_ALTS][MG_MAX_PAM_LEN][8];
};

__device__ __forceinline__ unsigned char comp_byte(unsigned char c) {
    // Bio.Seq.reverse_complement: IUPAC-aware, case preserving, other bytes unchanged
    const char* from = "ACGTMRWSYKVHDBNXacgtmrwsykvhdbnx";
    const char* to   = "TGCAKYWSRMBDHVNXtgcakywsrmbdhvnx";
#pragma unroll
    for (int i = 0; i < 32; ++i) if (c == (unsigned char)from[i]) return (unsigned char)to[i];
    return c;
}

__device__ __forceinline__ bool side_matches(const PamDev& side, const unsigned char* seq, int64_t n, bool minus,
                                             int64_t at, bool behind) {
    // behind: some alternative must END at `at` (exclusive); else START at `at`.  Index into the strand's text.
    if (side.n_alts == 0) return true;
    for (int a = 0; a < side.n_alts; ++a) {
        const int len = side.len[a];
        const int64_t s = behind ? at - len : at;
        if (s < 0 || s + len > n) continue;
        bool ok = true;
        for (int k = 0; k < len && ok; ++k) {
            const int64_t i = s + k;
            const unsigned char c = minus ? comp_byte(seq[n - 1 - i]) : seq[i];
            ok = (side.allow[a][k][c >> 5] >> (c & 31)) & 1u;
        }
        if (ok) return true;
    }
    return false;
}

__global__ void enumerate_kernel(const unsigned char* __restrict__ seq, int64_t n, const PamDev* __restrict__ five,
                                 const PamDev* __restrict__ three, int L, int64_t cap, int64_t* out_plus,
                                 int64_t* out_minus, unsigned long long* counters) {
    const int64_t nstart = n - L + 1;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= 2 * nstart) return;
    const bool minus = idx >= nstart;
    const int64_t s = minus ? idx - nstart : idx;
    if (!side_matches(*five, seq, n, minus, s, true)) return;
    if (!side_matches(*three, seq, n, minus, s + L, false)) return;
    const unsigned long long slot = atomicAdd(counters + (minus ? 1 : 0), 1ull);
    if ((int64_t)slot < cap) (minus ? out_minus : out_plus)[slot] = s;
}

static void fill_side(const mg_pam_side& in, PamDev& out) {
    out.n_alts = in.n_alts;
    memcpy(out.len, in.len, sizeof out.len);
    memcpy(out.allow, in.allow, sizeof out.allow);
}

}  // namespace mg

using namespace mg;

extern "C" int mg_enumerate_sites(const char* seq, int64_t len, const mg_pam* pam, int grna_len, int64_t cap,
                                  int64_t* out_plus, int64_t* n_plus, int64_t* out_minus, int64_t* n_minus, void* stream) {
    MG_CHECK_ARG(pam && n_plus && n_minus && grna_len >= 1 && len >= 0 && cap >= 0);
    MG_CHECK_ARG(pam->five.n_alts >= 0 && pam->five.n_alts <= MG_MAX_PAM_ALTS && pam->three.n_alts >= 0 && pam->three.n_alts <= MG_MAX_PAM_ALTS);
    *n_plus = *n_minus = 0;
    if (len < grna_len) return MG_OK;
    cudaStream_t s = (cudaStream_t)stream;
    unsigned char* d_seq = nullptr;
    PamDev* d_pam = nullptr;
    int64_t *d_plus = nullptr, *d_minus = nullptr;
    unsigned long long* d_cnt = nullptr;
    std::vector<PamDev> hp(2);
    fill_side(pam->five, hp[0]);
    fill_side(pam->three, hp[1]);
    int rc = MG_OK;
    auto fail = [&](cudaError_t e, const char* what) { set_error("%s: %s", what, cudaGetErrorString(e)); rc = MG_ERR_CUDA; };
    cudaError_t e;
    do {
        if ((e = cudaMalloc(&d_seq, (size_t)len + 16)) != cudaSuccess) { fail(e, "cudaMalloc seq"); break; }
        if ((e = cudaMalloc(&d_pam, sizeof(PamDev) * 2)) != cudaSuccess) { fail(e, "cudaMalloc pam"); break; }
        if ((e = cudaMalloc(&d_plus, sizeof(int64_t) * (size_t)(cap + 1))) != cudaSuccess) { fail(e, "cudaMalloc out"); break; }
        if ((e = cudaMalloc(&d_minus, sizeof(int64_t) * (size_t)(cap + 1))) != cudaSuccess) { fail(e, "cudaMalloc out"); break; }
        if ((e = cudaMalloc(&d_cnt, 16)) != cudaSuccess) { fail(e, "cudaMalloc cnt"); break; }
        cudaMemcpyAsync(d_seq, seq, (size_t)len, cudaMemcpyHostToDevice, s);
        cudaMemcpyAsync(d_pam, hp.data(), sizeof(PamDev) * 2, cudaMemcpyHostToDevice, s);
        cudaMemsetAsync(d_cnt, 0, 16, s);
        const int64_t total = 2 * (len - grna_len + 1);
        enumerate_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(d_seq, len, d_pam, d_pam + 1, grna_len, cap, d_plus, d_minus, d_cnt);
        unsigned long long cnt[2] = {0, 0};
        if ((e = cudaMemcpyAsync(cnt, d_cnt, 16, cudaMemcpyDeviceToHost, s)) != cudaSuccess) { fail(e, "D2H"); break; }
        if ((e = cudaStreamSynchronize(s)) != cudaSuccess) { fail(e, "enumerate kernel"); break; }
        *n_plus = (int64_t)cnt[0];
        *n_minus = (int64_t)cnt[1];
        if (*n_plus > cap || *n_minus > cap) { set_error("enumerate: %lld/%lld sites exceed capacity %lld", (long long)*n_plus, (long long)*n_minus, (long long)cap); rc = MG_ERR_OVERFLOW; break; }
        if (*n_plus) cudaMemcpyAsync(out_plus, d_plus, sizeof(int64_t) * (size_t)*n_plus, cudaMemcpyDeviceToHost, s);
        if (*n_minus) cudaMemcpyAsync(out_minus, d_minus, sizeof(int64_t) * (size_t)*n_minus, cudaMemcpyDeviceToHost, s);
        if ((e = cudaStreamSynchronize(s)) != cudaSuccess) { fail(e, "D2H sites"); break; }
        // the reference yields ascending starts; the device appends in arbitrary order
        std::sort(out_plus, out_plus + *n_plus);
        std::sort(out_minus, out_minus + *n_minus);
    } while (0);
    cudaFree(d_seq); cudaFree(d_pam); cudaFree(d_plus); cudaFree(d_minus); cudaFree(d_cnt);
    return rc;
}
