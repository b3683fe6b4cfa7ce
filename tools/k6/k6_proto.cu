// K6 prototype: the ungapped mismatch screen as a one-hot GEMM on tcgen05 tensor cores,
// measured against the integer pair-word scan (screen_pairs_kernel: 97 cells/clk/SM = 28.2 Tcell/s).
//
// Formulation: base b -> vertex v(b) of a tetrahedron with +-1 coordinates, (1,1,1) (1,-1,-1) (-1,1,-1) (-1,-1,1):
// v(a).v(b) = 3 if a == b else -1, so score(gRNA g, window w) = sum_j v(g_j).v(t_{w+j}) = 3 L - 4 mismatches, K = 3 L = 60 -> 64
// (zero padded), operands e4m3 (+-1 exact), accumulators f16 (exact: |score| <= 60) in TMEM.  mismatches <= 4  <=>  score >= 3L-16.
// One tile = M 128 gRNA x N 256 windows x K 64 = two tcgen05.mma.kind::f8f6f4 (K = 32 each) = 32 768 cells.
// Epilogue per tile: tcgen05.ld .pack::16b (two f16 accumulators per register) + packed 16x2 running max per gRNA row
// (__vimax3_s16x2: four cells per ALU instruction; f16 bit patterns of the scores order like s16 where it matters) - a row is
// only looked at again when its tile maximum reaches the threshold (hits are 4e-7 of the cells).
//
// This prototype measures the CEILING of that design: the B operand (one-hot text tile, Toeplitz: row n+1 = row n shifted by
// 3 bytes) is built once and reused by every tile, i.e. the cost of expanding the 2-bit genome into operand tiles (16 KB of
// shared-memory writes per tile on top of the 12 KB the MMA reads) is NOT included.  If the ceiling does not beat the integer
// scan, the full kernel cannot.   Modes: 0 = MMA + epilogue pipelined (two TMEM stages), 1 = MMA only, 2 = epilogue only.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o k6_proto k6_proto.cu ; run: ./k6_proto [tiles] [mode]
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s -> %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); exit(2); } } while (0)

constexpr int kM = 128, kN = 256, kK = 64;          // tile: gRNA x windows x one-hot depth (bytes)
constexpr int kL = 20;
// warps 0..EW-1: epilogue (warp w reads TMEM lanes 32 (w % 4) .. +31; with EW = 8 the two warps of a lane quarter split the 256
// columns), warp EW: MMA issuer.  OP: 0 = VIMNMX3.S16x2 running max (4 cells per instruction), 1 = HMNMX2 (max.f16x2, 2 cells per
// instruction), 2 = load only (one LOP3 per 3 registers keeps the loads alive): the TMEM read rate by itself.
constexpr uint32_t kTmemCols = 512;                  // two accumulator stages of 256 columns (one f16 per 32-bit column)

// K-major, no swizzle ("interleave") canonical layout in bytes: 8 rows x 16 B core matrices; rows / 8 -> SBO, 16-B K chunk -> LBO
__host__ __device__ inline uint32_t tile_off(int row, int kb, int rows) {
    return (uint32_t)((kb >> 4) * (rows * 16) + (row >> 3) * 128 + (row & 7) * 16 + (kb & 15));
}

__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;                          // descriptor version (Blackwell)
    return d;                                        // layout_type 0 = no swizzle, base_offset 0
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" :: "r"(a), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}

__device__ __forceinline__ void mma_f8_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
        :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                 :: "r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}

// 64 packed registers = 128 accumulator columns of this thread's TMEM lane
#define LD128(r, addr) \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x64.pack::16b.b32 " \
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, " \
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, " \
        "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, " \
        "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];" \
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), \
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), \
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), \
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), \
          "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), \
          "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), \
          "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), \
          "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63]) \
        : "r"(addr) : "memory")

template <int EW, int OP>
__global__ void __launch_bounds__(32 * (EW + 1), 1)
k6_kernel(const uint8_t* __restrict__ a_tile, const uint8_t* __restrict__ b_tile, int tiles, int mode,
          uint32_t* __restrict__ row_max_out, uint32_t* __restrict__ dump, unsigned long long* __restrict__ hits, int threshold_bits) {
    constexpr int kThreads = 32 * (EW + 1);
    constexpr int kHalves = EW == 4 ? 2 : 1;             // 128-column loads per thread and tile
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* sA = smem;                               // 8 KB
    uint8_t* sB = smem + kM * kK;                     // 16 KB
    __shared__ __align__(8) uint64_t full_bar[2], empty_bar[2];
    __shared__ uint32_t tmem_base_sh;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    for (int i = threadIdx.x; i < kM * kK / 16; i += kThreads) reinterpret_cast<uint4*>(sA)[i] = reinterpret_cast<const uint4*>(a_tile)[i];
    for (int i = threadIdx.x; i < kN * kK / 16; i += kThreads) reinterpret_cast<uint4*>(sB)[i] = reinterpret_cast<const uint4*>(b_tile)[i];
    if (threadIdx.x == 0) {
        mbar_init(&full_bar[0], 1); mbar_init(&full_bar[1], 1);
        mbar_init(&empty_bar[0], 32 * EW); mbar_init(&empty_bar[1], 32 * EW);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == EW) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"((uint32_t)__cvta_generic_to_shared(&tmem_base_sh)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // make the generic-proxy writes of the operand tiles visible to the tensor core (async proxy)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_sh;

    // instruction descriptor: D f16, A/B e4m3, both K-major, N = 256, M = 128
    const uint32_t idesc = (0u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(kN >> 3) << 17) | ((uint32_t)(kM >> 4) << 24);

    if (warp == EW) {
        if (mode != 2 && lane == 0) {
            const uint32_t a0 = (uint32_t)__cvta_generic_to_shared(sA), b0 = (uint32_t)__cvta_generic_to_shared(sB);
            for (int t = 0; t < tiles; ++t) {
                const int st = t & 1;
                if (mode == 0 && t >= 2) mbar_wait(&empty_bar[st], ((t >> 1) - 1) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d = tmem_base + (uint32_t)st * 256u;
#pragma unroll
                for (int k = 0; k < kK / 32; ++k) {
                    // K chunk pair k: two 16-byte chunks, LBO apart; SBO between 8-row groups
                    const uint64_t ad = make_desc(a0 + k * 2 * (kM * 16), kM * 16, 128);
                    const uint64_t bd = make_desc(b0 + k * 2 * (kN * 16), kN * 16, 128);
                    mma_f8_ss(d, ad, bd, idesc, k > 0 ? 1u : 0u);
                }
                mma_commit(&full_bar[st]);
            }
            if (mode == 1) {     // wait for the last commits so that the kernel time covers the MMAs
                mbar_wait(&full_bar[(tiles - 1) & 1], ((tiles - 1) >> 1) & 1);
                if (tiles > 1) mbar_wait(&full_bar[(tiles - 2) & 1], ((tiles - 2) >> 1) & 1);
            }
        }
    } else if (mode != 1) {
        const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(warp >> 2) * 128u;
        const int row = (warp & 3) * 32 + lane;
        uint32_t run0 = 0x80008000u, run1 = 0x80008000u;     // packed s16x2 minima (OP 1: f16x2 -inf)
        if (OP == 1) run0 = run1 = 0xFC00FC00u;
        unsigned long long my_hits = 0;
        const uint32_t thr2 = (uint32_t)threshold_bits | ((uint32_t)threshold_bits << 16);
        for (int t = 0; t < tiles; ++t) {
            const int st = t & 1;
            if (mode == 0) mbar_wait(&full_bar[st], (t >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t tmax0 = OP == 1 ? 0xFC00FC00u : 0x80008000u, tmax1 = tmax0;
#pragma unroll
            for (int h = 0; h < kHalves; ++h) {
                uint32_t r[64];
                LD128(r, lane_addr + (uint32_t)st * 256u + (uint32_t)h * 128u);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (dump != nullptr && t == 0 && blockIdx.x == 0) {
#pragma unroll
                    for (int i = 0; i < 64; ++i) dump[(size_t)row * 128 + ((warp >> 2) + h) * 64 + i] = r[i];
                }
                if (OP == 0) {
#pragma unroll
                    for (int i = 0; i < 64; i += 4) {
                        tmax0 = __vimax3_s16x2(tmax0, r[i], r[i + 1]);
                        tmax1 = __vimax3_s16x2(tmax1, r[i + 2], r[i + 3]);
                    }
                } else if (OP == 1) {
#pragma unroll
                    for (int i = 0; i < 64; i += 2) {
                        asm("max.f16x2 %0, %0, %1;" : "+r"(tmax0) : "r"(r[i]));
                        asm("max.f16x2 %0, %0, %1;" : "+r"(tmax1) : "r"(r[i + 1]));
                    }
                } else {
#pragma unroll
                    for (int i = 0; i + 2 < 64; i += 2) {
                        asm("lop3.b32 %0, %0, %1, %2, 0xFE;" : "+r"(tmax0) : "r"(r[i]), "r"(r[i + 1]));
                    }
                    tmax1 = tmax0 | r[63] | r[62];
                }
            }
            if (mode == 0) {
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                mbar_arrive(&empty_bar[st]);
            }
            // rare path stand-in: a tile whose maximum reaches the threshold would be re-examined cell by cell
            if (OP == 1) {
                uint32_t m;
                asm("max.f16x2 %0, %1, %2;" : "=r"(m) : "r"(tmax0), "r"(tmax1));
                const __half2 mh = *reinterpret_cast<const __half2*>(&m), th = *reinterpret_cast<const __half2*>(&thr2);
                if (__hbge2(mh, th) || __hge(__low2half(mh), __low2half(th)) || __hge(__high2half(mh), __high2half(th))) ++my_hits;
                asm("max.f16x2 %0, %0, %1;" : "+r"(run0) : "r"(tmax0));
                asm("max.f16x2 %0, %0, %1;" : "+r"(run1) : "r"(tmax1));
            } else {
                const uint32_t m = __vmaxs2(tmax0, tmax1);
                if (__vcmpges2(m, thr2)) ++my_hits;
                run0 = __vmaxs2(run0, tmax0);
                run1 = __vmaxs2(run1, tmax1);
            }
        }
        uint32_t m;
        if (OP == 1) asm("max.f16x2 %0, %1, %2;" : "=r"(m) : "r"(run0), "r"(run1));
        else m = __vmaxs2(run0, run1);
        int hi, lo;
        if (OP == 1) {
            const __half2 mh = *reinterpret_cast<const __half2*>(&m);
            hi = (int)__half2float(__high2half(mh)); lo = (int)__half2float(__low2half(mh));
            const __half best = __float2half((float)(hi > lo ? hi : lo));
            row_max_out[((size_t)blockIdx.x * 2 + (warp >> 2)) * kM + row] = (uint32_t)*reinterpret_cast<const unsigned short*>(&best);
        } else {
            hi = (int)(int16_t)(m >> 16); lo = (int)(int16_t)(m & 0xFFFF);
            row_max_out[((size_t)blockIdx.x * 2 + (warp >> 2)) * kM + row] = (uint32_t)(hi > lo ? hi : lo) & 0xFFFFu;
        }
        if (my_hits) atomicAdd(hits, my_hits);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == EW) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

static uint8_t e4m3(int v) { return v > 0 ? 0x38 : v < 0 ? 0xB8 : 0x00; }     // +1, -1, 0
static const int kVertex[4][3] = {{1, 1, 1}, {1, -1, -1}, {-1, 1, -1}, {-1, -1, 1}};

int main(int argc, char** argv) {
    const int tiles = argc > 1 ? atoi(argv[1]) : 4096;
    const int only_mode = argc > 2 ? atoi(argv[2]) : -1;
    int dev = 0, sms = 0, clock_khz = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CK(cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, dev));
    // a random text of kN + kL bases and kM gRNA, some planted near-copies
    srand(7);
    std::vector<int> text(kN + kL + 4), grna(kM * kL);
    for (auto& b : text) b = rand() & 3;
    for (int g = 0; g < kM; ++g) {
        const int w = rand() % kN, mm = g % 7;
        for (int j = 0; j < kL; ++j) grna[g * kL + j] = (g % 3 == 0) ? text[w + j] : (rand() & 3);
        for (int x = 0; x < mm && g % 3 == 0; ++x) grna[g * kL + rand() % kL] = rand() & 3;
    }
    std::vector<uint8_t> hA(kM * kK, 0), hB(kN * kK, 0);
    for (int g = 0; g < kM; ++g)
        for (int j = 0; j < kL; ++j)
            for (int c = 0; c < 3; ++c) hA[tile_off(g, 3 * j + c, kM)] = e4m3(kVertex[grna[g * kL + j]][c]);
    for (int n = 0; n < kN; ++n)
        for (int j = 0; j < kL; ++j)
            for (int c = 0; c < 3; ++c) hB[tile_off(n, 3 * j + c, kN)] = e4m3(kVertex[text[n + j]][c]);
    std::vector<int> want(kM * kN), want_max(kM, -1000);
    for (int g = 0; g < kM; ++g)
        for (int n = 0; n < kN; ++n) {
            int mm = 0;
            for (int j = 0; j < kL; ++j) mm += grna[g * kL + j] != text[n + j];
            want[g * kN + n] = 3 * kL - 4 * mm;
            if (want[g * kN + n] > want_max[g]) want_max[g] = want[g * kN + n];
        }
    uint8_t *dA, *dB; uint32_t *dMax, *dDump; unsigned long long* dHits;
    CK(cudaMalloc(&dA, hA.size())); CK(cudaMalloc(&dB, hB.size()));
    CK(cudaMalloc(&dMax, sizeof(uint32_t) * sms * 2 * kM)); CK(cudaMalloc(&dDump, sizeof(uint32_t) * 128 * 128)); CK(cudaMalloc(&dHits, 8));
    CK(cudaMemcpy(dA, hA.data(), hA.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), hB.size(), cudaMemcpyHostToDevice));
    const size_t smem = kM * kK + kN * kK;
    const __half thr_h = __float2half((float)(3 * kL - 16));
    unsigned short thr_bits; memcpy(&thr_bits, &thr_h, 2);
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const char* mode_names[3] = {"mma_plus_epilogue", "mma_only", "epilogue_only"};
    const char* op_names[3] = {"vimnmx3_s16x2", "hmnmx2_f16x2", "load_only"};
    int rc = 0;
    printf("{\"tile\": \"M128 x N256 x K64 e4m3 -> f16 (32768 cells)\", \"sms\": %d, \"clock_mhz\": %.0f, \"tiles_per_sm\": %d, \"runs\": [", sms, clock_khz / 1e3, tiles);
    bool first = true;
    auto run_cfg = [&](auto kern, int ew, int op) {
        const int threads = 32 * (ew + 1);
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        // ---- correctness: one tile, every accumulator against the CPU scores
        CK(cudaMemset(dHits, 0, 8));
        CK(cudaMemset(dDump, 0, sizeof(uint32_t) * 128 * 128));
        kern<<<1, threads, smem>>>(dA, dB, 1, 0, dMax, dDump, dHits, thr_bits);
        CK(cudaDeviceSynchronize());
        std::vector<uint32_t> hd(128 * 128), hmax(2 * kM);
        CK(cudaMemcpy(hd.data(), dDump, hd.size() * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(hmax.data(), dMax, 2 * kM * 4, cudaMemcpyDeviceToHost));
        long bad = 0, bad_max = 0;
        for (int g = 0; g < kM; ++g)
            for (int i = 0; i < 128; ++i) {
                const uint32_t pr = hd[(size_t)g * 128 + i];
                for (int half = 0; half < 2; ++half) {
                    unsigned short bits = (unsigned short)(half ? pr >> 16 : pr & 0xFFFF);
                    __half hv; memcpy(&hv, &bits, 2);
                    const int got = (int)__half2float(hv), n = 2 * i + half;
                    if (got != want[g * kN + n]) { if (bad < 3) fprintf(stderr, "EW%d OP%d mismatch g=%d n=%d got %d want %d\n", ew, op, g, n, got, want[g * kN + n]); ++bad; }
                }
            }
        if (op != 2)
            for (int g = 0; g < kM; ++g) {
                int best = -1000;
                for (int part = 0; part < (ew == 8 ? 2 : 1); ++part) {
                    unsigned short bits = (unsigned short)hmax[part * kM + g];
                    __half hv; memcpy(&hv, &bits, 2);
                    best = std::max(best, (int)__half2float(hv));
                }
                if (best != want_max[g]) ++bad_max;
            }
        if (bad || bad_max) rc = 1;
        printf("%s{\"epilogue_warps\": %d, \"epilogue_op\": \"%s\", \"check\": {\"cells\": %d, \"wrong_scores\": %ld, \"wrong_row_maxima\": %ld}", first ? "" : ", ",
               ew, op_names[op], kM * kN, bad, bad_max);
        first = false;
        for (int mode = 0; mode < 3; ++mode) {
            if (only_mode >= 0 && mode != only_mode) continue;
            if (mode == 1 && !(ew == 4 && op == 0)) continue;       // the MMA-only rate does not depend on the epilogue
            float best = 1e30f;
            for (int rep = 0; rep < 4; ++rep) {
                CK(cudaMemset(dHits, 0, 8));
                CK(cudaEventRecord(e0));
                kern<<<sms, threads, smem>>>(dA, dB, tiles, mode, dMax, nullptr, dHits, thr_bits);
                CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1));
                CK(cudaGetLastError());
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                if (rep > 0 && ms < best) best = ms;
            }
            const double cells = (double)sms * tiles * kM * kN;
            const double per_clk_sm = cells / (best * 1e-3) / sms / (clock_khz * 1e3);
            printf(", \"%s\": {\"ms\": %.3f, \"tcell_per_s\": %.2f, \"cells_per_clk_per_sm\": %.1f, \"clk_per_tile\": %.1f}", mode_names[mode], best,
                   cells / (best * 1e-3) / 1e12, per_clk_sm, (double)kM * kN / per_clk_sm);
        }
        printf("}");
        fflush(stdout);
    };
    run_cfg(k6_kernel<4, 0>, 4, 0);
    run_cfg(k6_kernel<4, 1>, 4, 1);
    run_cfg(k6_kernel<4, 2>, 4, 2);
    run_cfg(k6_kernel<8, 0>, 8, 0);
    run_cfg(k6_kernel<8, 1>, 8, 1);
    run_cfg(k6_kernel<8, 2>, 8, 2);
    printf("], \"integer_scan_for_comparison\": {\"kernel\": \"screen_pairs_kernel<20,4,1024,0>\", \"cells_per_clk_per_sm\": 97.0, \"tcell_per_s\": 28.2}}\n");
    return rc;
}
