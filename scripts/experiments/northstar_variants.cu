// scripts/experiments/northstar_variants.cu — MEASUREMENT ONLY, not part of libswec.so.
//
// BASELINE.json's north_star sketches the kernel as "TMA bulk copies of the log/antilog multiply tables
// into shared memory, and warp-shuffle XOR reductions across the 10 data lanes".  DESIGN.md §4 ships a
// different formulation (in-thread bit-plane Horner, no tables, no shuffles) and claims the sketched one
// cannot reach the roofline.  This file is the evidence: it implements the sketch literally,
//
//   lane_per_shard_logexp   lane (g, i) of a warp owns 16 bytes of data shard i at column c+g (three
//                           column groups × ten shards = 30 lanes); log[256] + antilog[512] tables in
//                           shared memory fetched by ONE cp.async.bulk (TMA) with mbarrier expect-tx;
//                           each lane multiplies its bytes by its four coefficients through
//                           antilog[log[b] + log[M[p][i]]]; the ten partial products of every output word
//                           are XOR-reduced with __shfl_down_sync (4 stages × 16 words); lane i = 0 stores.
//
// checks it bit for bit against the shipped kernel (swec_encode_device) on the same seeded shards, and
// times both with CUDA events.  Build + run: scripts/experiments/run_northstar_variants.sh on the GPU box.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "swec.h"

#define CK(x)                                                                              \
    do {                                                                                   \
        cudaError_t e_ = (x);                                                              \
        if (e_ != cudaSuccess) {                                                           \
            fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
            exit(1);                                                                       \
        }                                                                                  \
    } while (0)

struct Params {
    const uint8_t* in[10];
    uint8_t* out[4];
    uint64_t nvec;      // 16-byte columns per shard
    uint8_t logm[4][10];  // log of the parity coefficients (all non-zero for RS(10,4))
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(256) lane_per_shard_logexp(const __grid_constant__ Params p,
                                                             const uint8_t* __restrict__ tables /* 768 B */) {
    __shared__ __align__(128) uint8_t tab[768];  // [0,256) log, [256,768) antilog (doubled: no mod 255)
    __shared__ __align__(8) uint64_t mbar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(768u) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_u32(tab)),
                     "l"(tables), "r"(768u), "r"(smem_u32(&mbar))
                     : "memory");
    }
    {
        uint32_t done = 0;
        while (!done)
            asm volatile("{ .reg .pred q; mbarrier.try_wait.parity.shared::cta.b64 q, [%1], 0; selp.u32 %0, 1, 0, q; }"
                         : "=r"(done)
                         : "r"(smem_u32(&mbar))
                         : "memory");
    }
    const uint8_t* lg = tab;
    const uint8_t* ex = tab + 256;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane % 3, i = lane / 3;  // lanes 30, 31: i = 10 → idle, but they take part in the shuffles
    const bool active = i < 10;
    uint32_t lm[4] = {0, 0, 0, 0};
    if (active)
        for (int q = 0; q < 4; q++) lm[q] = p.logm[q][i];
    const uint64_t warps = (uint64_t)gridDim.x * 8;
    for (uint64_t c0 = ((uint64_t)blockIdx.x * 8 + warp) * 3; c0 < p.nvec; c0 += warps * 3) {
        const uint64_t c = c0 + g;
        uint32_t acc[4][4];
#pragma unroll
        for (int q = 0; q < 4; q++)
#pragma unroll
            for (int w = 0; w < 4; w++) acc[q][w] = 0;
        if (active && c < p.nvec) {
            const uint4 d = *reinterpret_cast<const uint4*>(p.in[i] + (c << 4));
            const uint32_t w4[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
            for (int w = 0; w < 4; w++)
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    const uint32_t byte = (w4[w] >> (8 * b)) & 0xffu;
                    if (byte) {
                        const uint32_t l = lg[byte];
#pragma unroll
                        for (int q = 0; q < 4; q++) acc[q][w] |= (uint32_t)ex[l + lm[q]] << (8 * b);
                    }
                }
        }
        // XOR-reduce over the ten data lanes of this column group (lanes g, g+3, …, g+27)
#pragma unroll
        for (int d = 8; d >= 1; d >>= 1) {
#pragma unroll
            for (int q = 0; q < 4; q++)
#pragma unroll
                for (int w = 0; w < 4; w++) {
                    const uint32_t t = __shfl_down_sync(0xffffffffu, acc[q][w], 3 * d);
                    if (i + d < 10) acc[q][w] ^= t;
                }
        }
        if (i == 0 && c < p.nvec) {
#pragma unroll
            for (int q = 0; q < 4; q++)
                *reinterpret_cast<uint4*>(p.out[q] + (c << 4)) = make_uint4(acc[q][0], acc[q][1], acc[q][2], acc[q][3]);
        }
    }
}

int main(int argc, char** argv) {
    const size_t gib = argc > 1 ? size_t(atoll(argv[1])) : 10;  // volume size in GiB (10 shards of gib/10)
    const size_t n = (gib << 30) / 10 & ~size_t(15);
    const double peak = argc > 2 ? atof(argv[2]) : 6501.2;
    CK(cudaSetDevice(0));

    // field tables, polynomial 0x11D, generator 2
    uint8_t tables[768];
    {
        uint8_t ex[512];
        int x = 1;
        for (int k = 0; k < 255; k++) {
            ex[k] = uint8_t(x);
            tables[x] = uint8_t(k);
            x <<= 1;
            if (x & 0x100) x ^= 0x11D;
        }
        for (int k = 255; k < 512; k++) ex[k] = ex[k - 255];
        tables[0] = 0;
        memcpy(tables + 256, ex, 512);
    }
    swec_encoder* enc = nullptr;
    if (swec_encoder_new(10, 4, 0, &enc) != SWEC_OK) return 1;
    uint8_t gen[14 * 10];
    swec_encoder_matrix(enc, gen);

    Params p;
    memset(&p, 0, sizeof p);
    uint8_t* ref[4];
    for (int i = 0; i < 10; i++) {
        CK(cudaMalloc((void**)&p.in[i], n));
        if (swec_synth_fill_device(0, (void*)p.in[i], uint64_t(i) * n, n & ~size_t(7), 0x5EA3EED5F00DCAFEull, nullptr)) return 1;
    }
    for (int q = 0; q < 4; q++) {
        CK(cudaMalloc((void**)&p.out[q], n));
        CK(cudaMalloc((void**)&ref[q], n));
        CK(cudaMemset(p.out[q], 0, n));
        for (int i = 0; i < 10; i++) p.logm[q][i] = tables[gen[(10 + q) * 10 + i]];
    }
    p.nvec = n / 16;
    uint8_t* dtab;
    CK(cudaMalloc((void**)&dtab, 768));
    CK(cudaMemcpy(dtab, tables, 768, cudaMemcpyHostToDevice));

    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    auto time_it = [&](auto&& launch, int reps) {
        for (int r = 0; r < 3; r++) launch();
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(a));
        for (int r = 0; r < reps; r++) launch();
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, a, b));
        return double(ms) / reps;
    };
    const void* din[10];
    for (int i = 0; i < 10; i++) din[i] = p.in[i];
    void* dref[4] = {ref[0], ref[1], ref[2], ref[3]};
    const double ms_ship = time_it([&] { if (swec_encode_device(enc, din, dref, n, nullptr)) exit(1); }, 10);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const double ms_lane = time_it([&] { lane_per_shard_logexp<<<sms * 8, 256>>>(p, dtab); }, 3);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());

    bool same = true;
    for (int q = 0; q < 4; q++) {
        uint64_t d0 = 0, d1 = 1;
        swec_digest_device(0, p.out[q], n, &d0, nullptr);
        swec_digest_device(0, ref[q], n, &d1, nullptr);
        same = same && d0 == d1;
    }
    const double bytes = 14.0 * double(n);
    auto line = [&](const char* name, double ms, const char* what) {
        printf("{\"kernel\": \"%s\", \"what\": \"%s\", \"shard_bytes\": %zu, \"ms\": %.3f, \"input_GBps\": %.1f, "
               "\"algorithmic_GBps\": %.1f, \"frac_of_hbm_peak\": %.4f, \"bit_exact_vs_shipped\": %s}\n",
               name, what, n, ms, 10.0 * n / ms / 1e6, bytes / ms / 1e6, bytes / ms / 1e6 / peak, same ? "true" : "false");
    };
    line("rs10x4_encode<512,2> (shipped)", ms_ship, "in-thread bit-plane Horner, SWAR, no tables, no shuffles");
    line("lane_per_shard_logexp (north_star sketch)", ms_lane,
         "lane per data shard, TMA-loaded log/antilog tables in smem, warp-shuffle XOR reduction over 10 lanes");
    return same ? 0 : 3;
}
