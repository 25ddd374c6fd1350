// seaweedfs_b200/csrc/device_common.cuh — device-side building blocks shared by the ahead-of-time
// kernels (kernels.cu) and the NVRTC-specialised kernels (jit.cc embeds this text verbatim, so it
// must not include any header).  sm_100a only.
//
// Everything here computes, per byte column x,  out_p[x] = XOR_i M[p][i] ⊗ in_i[x]  over
// GF(2^8)/0x11D — the arithmetic of reedsolomon.Encoder.Encode / Reconstruct
// (weed/storage/erasure_coding/ec_encoder.go:265,360).
#pragma once

#include "apply_params.h"

// ---- SWAR GF(2^8) primitives on four packed bytes ------------------------------------------
// 3-input XOR = one LOP3 (immLut 0x96)
__device__ __forceinline__ u32 swec_x3(u32 a, u32 b, u32 c) {
    u32 r;
    asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
#define SWEC_X2(a, b) ((a) ^ (b))
#define SWEC_X3(a, b, c) swec_x3((a), (b), (c))

// Multiply each of the four bytes by 2 (the field generator) and add s:
//   hi  = msb of every byte                       (LOP, alu pipe)
//   m   = (hi >> 7) * 0x1D  = hi * (0x1D<<25) >> 32  (IMAD.HI, fma pipe) — the reduction term
//   b2  = 2*(a - hi) = 2a - 2hi                   (two IMADs, fma pipe)   — bytes shifted left
//   out = b2 ^ m ^ s                              (one LOP3, alu pipe)
// Two alu-pipe and three fma-pipe instructions per step: the integer-ALU pipe is the scarce one.
// Variants of the step (SWEC_XT_VARIANT; the arithmetic is identical, the instruction mix is not):
//   0  hi = a & 0x80..; m = mulhi(hi, 0x1D<<25); b2 = 2a - 2hi; out = b2 ^ m ^ s      2 alu + 3 fma   (5 instr)
//   1  shift-based                                                                    4 alu + 1 fma   (5 instr)
//   2  mask = prmt(a) (msb of every byte replicated over the byte); a2 = 2a;
//      out = ((a2 & 0xFE..) ^ s) ^ (mask & 0x1D..)                                    3 alu + 1 fma   (4 instr)
//   3  call sites alternate between 0 and 2: 5 alu + 4 fma per two steps (9 instr) — neither pipe ahead
template <int V>
__device__ __forceinline__ u32 swec_xt1_v(u32 a, u32 s) {
    if (V == 2) {
        u32 mask, a2, u, r;
        asm("prmt.b32 %0, %1, 0, 0xba98;" : "=r"(mask) : "r"(a));
        asm("mul.lo.u32 %0, %1, 2;" : "=r"(a2) : "r"(a));
        asm("lop3.b32 %0, %1, 0xfefefefe, %2, 0x6a;" : "=r"(u) : "r"(a2), "r"(s));     // (a2 & c) ^ s
        asm("lop3.b32 %0, %1, 0x1d1d1d1d, %2, 0x6a;" : "=r"(r) : "r"(mask), "r"(u));   // (mask & c) ^ u
        return r;
    }
    const u32 hi = a & 0x80808080u;
    if (V == 1) {  // shift-based variant (all alu pipe except the multiply)
        const u32 m = (hi >> 7) * 0x1du;
        const u32 b2 = (a ^ hi) << 1;
        return swec_x3(b2, m, s);
    }
    u32 m, a2, b2;
    asm("mul.hi.u32 %0, %1, 0x3a000000;" : "=r"(m) : "r"(hi));
    asm("mul.lo.u32 %0, %1, 2;" : "=r"(a2) : "r"(a));
    asm("mad.lo.u32 %0, %1, 0xfffffffe, %2;" : "=r"(b2) : "r"(hi), "r"(a2));
    return swec_x3(b2, m, s);
}
#if SWEC_XT_VARIANT == 3
#define SWEC_XT1A(a, s) swec_xt1_v<0>((a), (s))
#define SWEC_XT1B(a, s) swec_xt1_v<2>((a), (s))
#else
#define SWEC_XT1A(a, s) swec_xt1_v<SWEC_XT_VARIANT>((a), (s))
#define SWEC_XT1B(a, s) swec_xt1_v<SWEC_XT_VARIANT>((a), (s))
#endif
#define SWEC_XT0A(a) SWEC_XT1A((a), 0u)
#define SWEC_XT0B(a) SWEC_XT1B((a), 0u)

// ---- streaming 16-byte global accesses (read-once / write-once data: keep it out of L1) -----
#ifndef SWEC_LD_POLICY
#define SWEC_LD_POLICY 0
#endif
__device__ __forceinline__ uint4 swec_ldg_stream(const u8* p) {
    uint4 r;
#if SWEC_LD_POLICY == 1
    asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v4.u32 {%0, %1, %2, %3}, [%4];"
#elif SWEC_LD_POLICY == 2
    asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];"
#elif SWEC_LD_POLICY == 3
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
#else
    asm volatile("ld.global.nc.L1::no_allocate.L2::128B.v4.u32 {%0, %1, %2, %3}, [%4];"
#endif
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void swec_stg_stream(u8* p, const uint4& v) {
#if defined(SWEC_ST_POLICY) && SWEC_ST_POLICY == 1
    asm volatile("st.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
#else
    asm volatile("st.global.cs.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
#endif
}

// One thread = UNROLL 16-byte column slices of every stream per iteration: UNROLL×K coalesced
// LDG.128 issued back to back (memory-level parallelism), 4×UNROLL independent 32-bit Horner
// evaluations (ILP), UNROLL×R coalesced STG.128.
template <class Combiner>
__device__ __forceinline__ void swec_combine_vec(const uint4 (&d)[Combiner::K], uint4 (&o)[Combiner::R]) {
    constexpr int K = Combiner::K, R = Combiner::R;
    u32 x[K], y[R];
#pragma unroll
    for (int i = 0; i < K; i++) x[i] = d[i].x;
    Combiner::combine(x, y);
#pragma unroll
    for (int r = 0; r < R; r++) o[r].x = y[r];
#pragma unroll
    for (int i = 0; i < K; i++) x[i] = d[i].y;
    Combiner::combine(x, y);
#pragma unroll
    for (int r = 0; r < R; r++) o[r].y = y[r];
#pragma unroll
    for (int i = 0; i < K; i++) x[i] = d[i].z;
    Combiner::combine(x, y);
#pragma unroll
    for (int r = 0; r < R; r++) o[r].z = y[r];
#pragma unroll
    for (int i = 0; i < K; i++) x[i] = d[i].w;
    Combiner::combine(x, y);
#pragma unroll
    for (int r = 0; r < R; r++) o[r].w = y[r];
}

template <class Combiner, bool BLOCKED, int UNROLL = 1>
__device__ __forceinline__ void swec_horner_body(const SwecApplyParams& p) {
    constexpr int K = Combiner::K, R = Combiner::R;
    auto in_offset = [&](u64 vv) -> u64 {
        u64 ioff = vv << 4;
        if (BLOCKED) {
            const u64 row = p.block_shift >= 0 ? (vv >> p.block_shift) : (vv / p.block_vecs);
            ioff += row * p.row_extra;
        }
        return ioff;
    };
    // A CTA iteration covers UNROLL*blockDim.x consecutive vectors of every stream (a contiguous
    // UNROLL*4 KiB run per stream for 256 threads): slice u of thread t is vector base+u*blockDim+t,
    // so every LDG.128/STG.128 warp instruction is one fully coalesced 512-byte segment.
    const u64 tile = (u64)blockDim.x * UNROLL;
    const u64 stride = (u64)gridDim.x * tile;
    for (u64 base = (u64)blockIdx.x * tile; base < p.nvec; base += stride) {
        uint4 d[UNROLL][K];
        bool live[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            const u64 v = base + (u64)u * blockDim.x + threadIdx.x;
            live[u] = v < p.nvec;
            if (live[u]) {
                const u64 ioff = in_offset(v);
#pragma unroll
                for (int i = 0; i < K; i++) d[u][i] = swec_ldg_stream(p.in[i] + ioff);
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            if (live[u]) {
                uint4 o[R];
                swec_combine_vec<Combiner>(d[u], o);
                const u64 off = (base + (u64)u * blockDim.x + threadIdx.x) << 4;
#pragma unroll
                for (int r = 0; r < R; r++) swec_stg_stream(p.out[r] + off, o[r]);
            }
        }
    }
}
