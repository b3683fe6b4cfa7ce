// K3: exact-occurrence mask finder.  Replaces mask_identical -> blast_mask (megablast, keep only
// end-to-end perfect HSPs; filter_grna.py:67-119): every genome interval equal to a mask sequence over
// its full length, on either strand.  A 2-bit seed of up to 32 bases is compared at every genome position
// (XOR + AND on 64 bits); seed matches are extended base by base (rare).
#include <algorithm>

#include "common.cuh"

namespace mg {

struct Seed {               // one (sequence, strand)
    uint32_t key_lo, key_hi, mask_lo, mask_hi;
    int32_t seq;            // sequence index
    int32_t strand;         // +1 / -1
    int32_t len;
    int32_t off;            // offset of this orientation's codes in d_codes
};

constexpr int kSeedChunk = 1024;

__global__ void __launch_bounds__(256)
maskfind_kernel(GenomeView gv, const Seed* __restrict__ seeds, int n_seeds, const uint8_t* __restrict__ codes,
                int64_t p_begin, int64_t p_end, int64_t cap, int32_t* out_seed, int64_t* out_pos, unsigned long long* counter) {
    __shared__ Seed sh[kSeedChunk];
    const int64_t p = p_begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t lo = 0, hi = 0, inv = 0xFFFFFFFFu;
    const bool active = p < p_end;
    if (active) load_window(gv, p, lo, hi, inv);
    for (int base = 0; base < n_seeds; base += kSeedChunk) {
        const int m = min(kSeedChunk, n_seeds - base);
        __syncthreads();
        for (int i = threadIdx.x; i < m; i += blockDim.x) sh[i] = seeds[base + i];
        __syncthreads();
        if (!active) continue;
        for (int i = 0; i < m; ++i) {
            const Seed sd = sh[i];
            if (((lo ^ sd.key_lo) & sd.mask_lo) | ((hi ^ sd.key_hi) & sd.mask_hi)) continue;
            const int slen = min(sd.len, 32);
            const uint32_t need = slen >= 32 ? 0xFFFFFFFFu : ((1u << slen) - 1u);
            if (inv & need) continue;
            bool ok = true;
            for (int k = 32; k < sd.len && ok; ++k) ok = base_at(gv, p + k) == codes[sd.off + k];
            if (!ok) continue;
            const unsigned long long slot = atomicAdd(counter, 1ull);
            if ((int64_t)slot < cap) { out_seed[slot] = base + i; out_pos[slot] = p; }
        }
    }
}

}  // namespace mg

using namespace mg;

extern "C" int mg_mask_find(const mg_genome* g, const char* seqs, const int64_t* offsets, int n_seqs, int64_t cap,
                            int32_t* out_seq, int32_t* out_contig, int64_t* out_start, int64_t* out_end,
                            int8_t* out_strand, int64_t* n_found, void* stream) {
    MG_CHECK_ARG(g && offsets && n_seqs >= 0 && n_found && cap >= 0);
    *n_found = 0;
    cudaStream_t s = (cudaStream_t)stream;
    std::vector<Seed> seeds;
    std::vector<uint8_t> codes;
    auto code_of = [](char c) -> int {
        switch (c) { case 'A': case 'a': return 0; case 'C': case 'c': return 1; case 'G': case 'g': return 2;
                     case 'T': case 't': return 3; default: return -1; }
    };
    for (int i = 0; i < n_seqs; ++i) {
        const int64_t a = offsets[i], b = offsets[i + 1];
        const int64_t len = b - a;
        if (len <= 0) continue;
        MG_CHECK_ARG(len < ((int64_t)1 << 30));
        if (g->is_chunk && len + 64 > g->halo) {      // an occurrence starting in the owned range must end inside the stored one
            set_error("mask sequence %d (%lld bases) is longer than the chunk's halo (%lld): build the chunk with halo >= longest mask + 64", i, (long long)len, (long long)g->halo);
            return MG_ERR_ARG;
        }
        std::vector<uint8_t> fwd((size_t)len);
        bool plain = true;
        for (int64_t k = 0; k < len; ++k) { const int c = code_of(seqs[a + k]); if (c < 0) { plain = false; break; } fwd[(size_t)k] = (uint8_t)c; }
        if (!plain) { set_error("mask sequence %d has a non-ACGT character; the caller must route it to the exact string search", i); return MG_ERR_UNSUPPORTED; }
        std::vector<uint8_t> rev((size_t)len);
        for (int64_t k = 0; k < len; ++k) rev[(size_t)k] = (uint8_t)(3 - fwd[(size_t)(len - 1 - k)]);
        for (int strand = 1; strand >= -1; strand -= 2) {
            const std::vector<uint8_t>& v = strand > 0 ? fwd : rev;
            if (strand < 0 && rev == fwd) continue;      // palindrome: one interval, like set(Masked) (MINORg.py:1953-1955)
            Seed sd{};
            sd.seq = i; sd.strand = strand; sd.len = (int32_t)len; sd.off = (int32_t)codes.size();
            const int slen = (int)std::min<int64_t>(len, 32);
            for (int k = 0; k < slen; ++k) {
                if (k < 16) { sd.key_lo |= (uint32_t)v[k] << (2 * k); sd.mask_lo |= 3u << (2 * k); }
                else { sd.key_hi |= (uint32_t)v[k] << (2 * (k - 16)); sd.mask_hi |= 3u << (2 * (k - 16)); }
            }
            codes.insert(codes.end(), v.begin(), v.end());
            seeds.push_back(sd);
        }
    }
    if (seeds.empty() || g->glen == 0) return MG_OK;
    MG_CHECK_ARG(codes.size() < ((size_t)1 << 31));
    Seed* d_seeds = nullptr; uint8_t* d_codes = nullptr; int32_t* d_oseed = nullptr; int64_t* d_opos = nullptr;
    unsigned long long* d_cnt = nullptr;
    int rc = MG_OK;
    cudaError_t e = cudaSuccess;
    auto fail = [&](const char* what) { set_error("%s: %s", what, cudaGetErrorString(e)); rc = MG_ERR_CUDA; };
    std::vector<int32_t> h_seed;
    std::vector<int64_t> h_pos;
    do {
        if ((e = cudaMalloc(&d_seeds, sizeof(Seed) * seeds.size())) != cudaSuccess) { fail("cudaMalloc"); break; }
        if ((e = cudaMalloc(&d_codes, codes.size() + 16)) != cudaSuccess) { fail("cudaMalloc"); break; }
        if ((e = cudaMalloc(&d_oseed, sizeof(int32_t) * (size_t)(cap + 1))) != cudaSuccess) { fail("cudaMalloc"); break; }
        if ((e = cudaMalloc(&d_opos, sizeof(int64_t) * (size_t)(cap + 1))) != cudaSuccess) { fail("cudaMalloc"); break; }
        if ((e = cudaMalloc(&d_cnt, 8)) != cudaSuccess) { fail("cudaMalloc"); break; }
        cudaMemcpyAsync(d_seeds, seeds.data(), sizeof(Seed) * seeds.size(), cudaMemcpyHostToDevice, s);
        cudaMemcpyAsync(d_codes, codes.data(), codes.size(), cudaMemcpyHostToDevice, s);
        cudaMemsetAsync(d_cnt, 0, 8, s);
        // a chunk reports the occurrences that START in its owned range (each occurrence has exactly one owner)
        const int64_t pb = g->own_begin, pe = std::min(g->own_end, g->glen);
        const unsigned grid = (unsigned)((std::max<int64_t>(pe - pb, 0) + 255) / 256);
        if (grid) maskfind_kernel<<<grid, 256, 0, s>>>(g->view(), d_seeds, (int)seeds.size(), d_codes, pb, pe, cap, d_oseed, d_opos, d_cnt);
        unsigned long long cnt = 0;
        if ((e = cudaMemcpyAsync(&cnt, d_cnt, 8, cudaMemcpyDeviceToHost, s)) != cudaSuccess) { fail("D2H"); break; }
        if ((e = cudaStreamSynchronize(s)) != cudaSuccess) { fail("maskfind kernel"); break; }
        *n_found = (int64_t)cnt;
        if ((int64_t)cnt > cap) { set_error("mask_find: %llu intervals exceed capacity %lld", cnt, (long long)cap); rc = MG_ERR_OVERFLOW; break; }
        h_seed.resize((size_t)cnt); h_pos.resize((size_t)cnt);
        if (cnt) {
            cudaMemcpyAsync(h_seed.data(), d_oseed, sizeof(int32_t) * (size_t)cnt, cudaMemcpyDeviceToHost, s);
            cudaMemcpyAsync(h_pos.data(), d_opos, sizeof(int64_t) * (size_t)cnt, cudaMemcpyDeviceToHost, s);
            if ((e = cudaStreamSynchronize(s)) != cudaSuccess) { fail("D2H intervals"); break; }
        }
    } while (0);
    cudaFree(d_seeds); cudaFree(d_codes); cudaFree(d_oseed); cudaFree(d_opos); cudaFree(d_cnt);
    if (rc) return rc;
    // global -> per-contig coordinates, deterministic order
    std::vector<size_t> order(h_seed.size());
    for (size_t i = 0; i < order.size(); ++i) order[i] = i;
    std::sort(order.begin(), order.end(), [&](size_t x, size_t y) {
        if (seeds[h_seed[x]].seq != seeds[h_seed[y]].seq) return seeds[h_seed[x]].seq < seeds[h_seed[y]].seq;
        if (h_pos[x] != h_pos[y]) return h_pos[x] < h_pos[y];
        return seeds[h_seed[x]].strand > seeds[h_seed[y]].strand;
    });
    for (size_t r = 0; r < order.size(); ++r) {
        const size_t i = order[r];
        const Seed& sd = seeds[h_seed[i]];
        const int64_t gp = h_pos[i];
        int c = (int)(std::upper_bound(g->cstart.begin(), g->cstart.end(), gp) - g->cstart.begin()) - 1;
        out_seq[r] = sd.seq; out_contig[r] = c; out_start[r] = gp - g->cstart[c]; out_end[r] = gp - g->cstart[c] + sd.len;
        out_strand[r] = (int8_t)sd.strand;
    }
    return MG_OK;
}
