// tcgen05 GEMM engine: fp32-grade products on the 5th-gen tensor cores by error-compensated TF32 splitting.
//
//   Y = epilogue( LN?(gather(A)) [M,K] * W^T ),  W [N,K] (nn.Linear layout = K-major B operand)
//
// Why 3xTF32: the RVQ code indices must match the fp32 reference bit for bit, and plain TF32 / BF16 inputs flip
// 3..742 of 3600 codes per stream (hard part 1).  Every operand x is split into
// hi = tf32(x), lo = tf32(x - hi); D += A_lo*B_hi + A_hi*B_lo + A_hi*B_hi in the fp32 TMEM accumulator
// (the dropped lo*lo term is 2^-22 relative).  Weights are split once at escb_finalize(); activations are split
// by the A-producer warps after the fused gather / LayerNorm.  The tensor core TRUNCATES its accumulator (~0.6 ulp of
// bias per MMA), so the hi*hi products and the corrections of long reductions go to separate TMEM accumulators that the
// epilogue adds in fp32 ("accumulator split": acc_policy below, TcWeight::nmain / corr).
//
// One persistent CTA per SM (832 threads), warp-specialised, looping over 128 x NT output tiles (NT <= 512).
// 24 of the 26 warps are shared between the epilogue and the A producers in one of three splits (struct Roles:
// 8 + 16, 12 + 12 or 16 + 8, chosen per layer class from measurements); shown here for 8 + 16, warp ids as in the plain
// GEMMs (the fused attention kernel puts its epilogue warps above the producers):
//   warps 0-7   epilogue : tcgen05.ld a finished sub-tile (one TMEM lane quadrant x one column share per warp),
//                          transpose 32x16 chunks through a swizzled smem staging tile so that global stores and
//                          residual loads are 64-byte row segments, apply bias / GELU / residual / scatter functor -
//                          or, for the fused qkv kernel, run the whole window-attention core on the accumulator;
//   warps 8-23  producer : gather the logical A rows (window partition + cyclic shift, frequency-row pairing,
//                          im2col ... loaders.cuh), LayerNorm, cvt.rna.tf32 split, write the hi / lo images of a
//                          32-wide K block in the 128-byte-swizzled K-major layout into a 3-slot ring.  The
//                          (tile, K block) jobs form one flat stream with the loads of four jobs in flight per
//                          thread and the following tile prefetched into L2;
//   warp 24     MMA      : one elected lane (elect.sync) issues tcgen05.mma.cta_group::1.kind::tf32 (M=128, N=BN,
//                          K=8) x 3 per k-step and sub-tile; tcgen05.commit releases the A slot / weight slot and
//                          publishes each sub-tile's accumulator;
//   warp 25     weights  : pre-swizzled [hi | lo] weight images, ONE cp.async.bulk (SASS UBLKCP) per (K block,
//                          sub-tile).  Layers whose whole n-tile fits keep it RESIDENT in smem for the life of
//                          the CTA; the others stream it through a 2..8 slot ring.
// An output tile is nsub sub-tiles of BN columns, each with its own TMEM region and full/empty barriers, so the
// A operand is produced ONCE for up to 512 output columns while the epilogue of a sub-tile overlaps the MMAs of
// the next ones (and of the next tile when two tiles fit the 512 columns).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <atomic>

#include "gemm.cuh"

namespace escb {
namespace tc {

constexpr int BM = 128;               // UMMA M
constexpr int KB = 32;                // tf32 per 128-byte swizzle row = one K block
constexpr int WARPS = 26;              // 24 epilogue + producer warps, 1 MMA warp, 1 weight-loader warp
constexpr int THREADS = WARPS * 32;
constexpr int MAX_NB = 8;             // weight ring slots (upper bound)
constexpr int MAX_BN = 208;
constexpr int A_SLOT = 2 * BM * 128;  // hi + lo images of one A block
constexpr int TMEM_COLS = 512;
constexpr int MAX_REG = 8;            // accumulator regions (sub-tiles in flight)
constexpr int SMEM_MAX = 232448;      // 227 KB opt-in limit per CTA
constexpr int MAX_NA = 3;
constexpr int NBARS = 2 * MAX_NA + 2 * MAX_NB + 2 * MAX_REG;

// Role split: E epilogue warps per TMEM lane quadrant.  E = 2: 8 epilogue + 16 producer warps, 3 A slots (the
// default, producer-heavy layers such as mlp2); E = 4: 16 + 8, 2 A slots (epilogue-heavy layers: proj, PatchSplit,
// the PVQ up-projection); E = 3: 12 + 12, where a producer thread owns rows r0, r0 + 48 and - for r0 < 32 only -
// r0 + 96 (the GELU GEMM and the fused attention kernel, which are lopsided either way with 8 + 16 or 16 + 8).
// PW: producer warps (0: 24 - epilogue warps).  The fused attention kernel of the top level (head width 15, C = 45)
// runs 12 + 8: its A operand is two K blocks per tile, eight producer warps keep up, and the smaller CTA (704 threads)
// lifts the register cap from 72 to 88, which the attention core - the critical role, a long dependent chain per
// thread - uses (275.7 -> 253.4 us per launch; the deeper levels, which need the producers, measured 3-7 % slower).
#ifndef ESCB_ATTN_PROD15
#define ESCB_ATTN_PROD15 8
#endif
template <int E, bool ATTN = false, int PW = 0>
struct Roles {
    static constexpr int EPI_WARPS = 4 * E;
    static constexpr int PROD_WARPS = PW > 0 ? PW : 24 - EPI_WARPS;
    static constexpr int THREADS = (EPI_WARPS + PROD_WARPS + 2) * 32;
    static constexpr int PROD_THREADS = PROD_WARPS * 32;
    static constexpr int RPT = (32 + PROD_WARPS - 1) / PROD_WARPS;   // rows per producer thread: 2, 3 (12 warps: the last one only for r0 < 32) or 4
    static constexpr int ROW_STEP = 4 * PROD_WARPS;        // distance between a thread's rows
    static constexpr int DEPTH = 8 / RPT;                  // producer jobs with loads in flight
    static constexpr int NA = E == 4 ? 2 : 3;              // A ring slots
    // per-warp staging: a 32 x 16 transpose tile; the fused attention epilogue keeps a K / V image [HD][32] and an
    // output tile [32][HD + 4] there, up to 896 floats (HD = 24)
    static constexpr int STG_WARP_BYTES = ATTN ? 3584 : 2048;
    static constexpr int STG_BYTES = EPI_WARPS * STG_WARP_BYTES;
    static constexpr int CTX_BYTES = EPI_WARPS * 32 * 16;
    static constexpr int TAIL_BYTES = STG_BYTES + CTX_BYTES + EPI_WARPS * 32 * 4 + NBARS * 8 + 16;
    static constexpr int B_BUDGET = SMEM_MAX - 1024 - NA * A_SLOT - TAIL_BYTES;
};
#ifndef ESCB_ATTN_E
#define ESCB_ATTN_E 3
#endif
constexpr int kAttnE = ESCB_ATTN_E;   // role split of the fused attention kernel: 3 = 12 epilogue + 12 producer warps (5.89 ms per step; 4 = 16 + 8: 5.98 ms)
// role-split code stored with a packed weight (TcWeight::wide): 0 = 8 epilogue + 16 producer warps (E = 2),
// 1 = 16 + 8 (E = 4), 2 = 12 + 12 (E = 3)
constexpr int role_e(int wide) { return wide == 1 ? 4 : (wide == 2 ? 3 : 2); }
constexpr int role_code(int e) { return e == 4 ? 1 : (e == 3 ? 2 : 0); }
inline int b_budget(int wide) { return wide == 1 ? Roles<4>::B_BUDGET : (wide == 2 ? Roles<3>::B_BUDGET : Roles<2>::B_BUDGET); }
inline int b_budget_attn() { return Roles<kAttnE, true>::B_BUDGET; }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Parity wait.  try_wait suspends the warp in hardware (up to the hint, woken as soon as the phase completes), so
// the loop rarely spins; a wait that is still pending after ~2 s of suspensions traps (a launch failure the
// caller sees as ESCB_ECUDA) instead of hanging the device.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done, spins = 0;
    for (;;) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}" : "=r"(done) : "r"(bar), "r"(parity), "r"(2000u) : "memory");
        if (done) break;
        if (++spins > 4000000u) __trap();
    }
}
// TMA bulk copy global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// One lane of a converged warp.  Issuing tcgen05.mma / cp.async.bulk under `if (elect_one())` lets ptxas emit them as
// plain uniform-datapath instructions; under `if (lane == 0)` each one is wrapped in an ELECT / BRA.U.ANY loop
// (5 extra dependent instructions per MMA, which made the single issuing thread the bound of every N <= 192 GEMM).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t"
        "}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}

// D[tmem] (+)= A[smem] * B[smem]^T, M = 128, N from idesc, K = 8 (tf32)
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// N consecutive accumulator columns of this thread's TMEM lane (N = 16a + 8b + 4c + 2d), no wait
template <int N>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, float* v) {
    if constexpr (N >= 16) {
        uint32_t r[16];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(taddr));
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
        tmem_ld_cols<N - 16>(taddr + 16, v + 16);
    } else if constexpr (N >= 8) {
        uint32_t r[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                     : "r"(taddr));
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
        tmem_ld_cols<N - 8>(taddr + 8, v + 8);
    } else if constexpr (N >= 4) {
        uint32_t r[4];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr));
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
        tmem_ld_cols<N - 4>(taddr + 4, v + 4);
    } else if constexpr (N >= 2) {
        uint32_t r[2];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(taddr));
        v[0] = __uint_as_float(r[0]);
        v[1] = __uint_as_float(r[1]);
        tmem_ld_cols<N - 2>(taddr + 2, v + 2);
    } else if constexpr (N == 1) {
        uint32_t r;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr));
        v[0] = __uint_as_float(r);
    }
}
// wait for the loads above; the empty volatile asms pin every consumer of v[] behind the wait
template <int N>
__device__ __forceinline__ void tmem_ld_wait(float* v) {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < N; ++i) asm volatile("" : "+f"(v[i]) :: "memory");
}

// Sum of the `accs` accumulators of a split region (N columns of each, BN columns apart): mains in order, the
// correction accumulator last.  The first load is left in flight when accs == 1 (the caller waits either way).
template <int N>
__device__ __forceinline__ void tmem_ld_acc(uint32_t taddr, float* v, int accs, int BN) {
    tmem_ld_cols<N>(taddr, v);
    for (int a = 1; a < accs; ++a) {
        float u[N];
        tmem_ld_cols<N>(taddr + (uint32_t)(a * BN), u);
        tmem_ld_wait<N>(u);                                // completes the loads of v too
#pragma unroll
        for (int i = 0; i < N; ++i) { asm volatile("" : "+f"(v[i])); v[i] += u[i]; }
    }
}

template <int N>
__device__ __forceinline__ void add_bias(float* v, const float* __restrict__ b) {     // b 8-byte aligned, N even
    if constexpr (N % 4 == 0) {
#pragma unroll
        for (int c = 0; c < N / 4; ++c) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(b) + c);
            v[4 * c] += t.x; v[4 * c + 1] += t.y; v[4 * c + 2] += t.z; v[4 * c + 3] += t.w;
        }
    } else {
#pragma unroll
        for (int c = 0; c < N / 2; ++c) {
            const float2 t = __ldg(reinterpret_cast<const float2*>(b) + c);
            v[2 * c] += t.x; v[2 * c + 1] += t.y;
        }
    }
}

// post-GEMM LayerNorm of N accumulator columns of one row: v = rstd * (v - mean * cs) + f * bw
template <int N>
__device__ __forceinline__ void ln_post(float* v, const float* __restrict__ cs, const float* __restrict__ bw, float mean,
                                        float rstd, float f) {
    if constexpr (N % 4 == 0) {
#pragma unroll
        for (int c = 0; c < N / 4; ++c) {
            const float4 s4 = __ldg(reinterpret_cast<const float4*>(cs) + c), b4 = __ldg(reinterpret_cast<const float4*>(bw) + c);
            v[4 * c] = fmaf(rstd, fmaf(-mean, s4.x, v[4 * c]), f * b4.x);
            v[4 * c + 1] = fmaf(rstd, fmaf(-mean, s4.y, v[4 * c + 1]), f * b4.y);
            v[4 * c + 2] = fmaf(rstd, fmaf(-mean, s4.z, v[4 * c + 2]), f * b4.z);
            v[4 * c + 3] = fmaf(rstd, fmaf(-mean, s4.w, v[4 * c + 3]), f * b4.w);
        }
    } else {
#pragma unroll
        for (int c = 0; c < N / 2; ++c) {
            const float2 s2 = __ldg(reinterpret_cast<const float2*>(cs) + c), b2 = __ldg(reinterpret_cast<const float2*>(bw) + c);
            v[2 * c] = fmaf(rstd, fmaf(-mean, s2.x, v[2 * c]), f * b2.x);
            v[2 * c + 1] = fmaf(rstd, fmaf(-mean, s2.y, v[2 * c + 1]), f * b2.y);
        }
    }
}

// epilogues that run the window-attention core on the accumulator instead of storing it define kAttn
template <class T, class = void> struct IsAttn { static constexpr bool value = false; };
template <class T> struct IsAttn<T, decltype((void)T::kAttn)> { static constexpr bool value = true; };

// producer-warp override per epilogue type (see Roles)
template <class T, class = void> struct AttnProd { static constexpr int value = 0; };
#ifndef ESCB_ATTN_PROD
#define ESCB_ATTN_PROD 0
#endif
#ifndef ESCB_ATTN_PROD72
#define ESCB_ATTN_PROD72 8
#endif
// head widths 12 and 24 are the C = 72 levels (12 also the decoder's C = 144 level, which measured neutral): 239 -> 228 us
template <class T> struct AttnProd<T, decltype((void)T::kAttn)> {
    static constexpr int value = T::HD == 15 ? ESCB_ATTN_PROD15 : ((T::HD == 12 || T::HD == 24) ? ESCB_ATTN_PROD72 : ESCB_ATTN_PROD);
};
template <int E, class EP> using RolesFor = Roles<E, IsAttn<EP>::value, AttnProd<EP>::value>;

// K-major, 128-byte swizzle, 8-row groups 1024 bytes apart; version 1 (sm_100)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// c_format F32 (1 << 4), a/b format TF32 (2 << 7, 2 << 10), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
__device__ __forceinline__ uint32_t make_idesc(int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

// Round to tf32 (nearest, ties away): cvt.rna.tf32.f32 compiles to five SASS instructions (its NaN / infinity handling:
// FSETP + IADD + LOP3 + SEL ...); for the finite values of this path the same bits come out of "add half an ulp of the
// 10-bit mantissa, clear the 13 low bits" in two.  ESCB_TF32_CVT restores the instruction (A/B builds).
__device__ __forceinline__ float tf32_rna(float x) {
#ifdef ESCB_TF32_CVT
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
#else
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
#endif
}
__device__ __forceinline__ void split4(const float4 v, float4& hi, float4& lo) {
    hi.x = tf32_rna(v.x); lo.x = tf32_rna(v.x - hi.x);
    hi.y = tf32_rna(v.y); lo.y = tf32_rna(v.y - hi.y);
    hi.z = tf32_rna(v.z); lo.z = tf32_rna(v.z - hi.z);
    hi.w = tf32_rna(v.w); lo.w = tf32_rna(v.w - hi.w);
}

#ifdef ESCB_TC_TRACE
#define TC_T0(v) const long long v = clock64()
#define TC_ACC(slot, v) tr[slot] += clock64() - (v)
#else
#define TC_T0(v)
#define TC_ACC(slot, v)
#endif

template <bool LN, class AL, class EP, int E, bool LNP = false>
__global__ void __launch_bounds__((RolesFor<E, EP>::THREADS), 1)
tc_gemm_kernel(const AL al, const LnParams ln, const TcWeight w, const long long M, const EP ep, const int ntiles,
               const int NB, const int resident, const int pf_dist) {
    static_assert(sizeof(typename EP::Row) <= 16, "epilogue row context must fit 16 bytes");
    using R = RolesFor<E, EP>;
    constexpr int EPI_WARPS = R::EPI_WARPS, PROD_WARPS = R::PROD_WARPS, PROD_THREADS = R::PROD_THREADS;
    constexpr int RPT = R::RPT, ROW_STEP = R::ROW_STEP, DEPTH = R::DEPTH, NA = R::NA;
    constexpr int STG_BYTES = R::STG_BYTES, CTX_BYTES = R::CTX_BYTES;
    // The warp schedulers favour high warp ids.  The plain GEMMs are bound by their A producers, which therefore sit
    // above the epilogue warps; the attention epilogue is the critical role of the fused qkv kernel and sits on top.
    constexpr int EPI_BASE = IsAttn<EP>::value ? PROD_WARPS : 0, PROD_BASE = IsAttn<EP>::value ? 0 : EPI_WARPS;
    extern __shared__ uint8_t smem_raw[];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int K = w.K, BN = w.BN, nkb = w.nkb, ntn = w.ntn, nsub = w.nsub;
    const int NT = BN * nsub;                             // columns of one output tile
    // accumulator regions of BN columns form a ring over TMEM: sub-tile number g (counted over this CTA's tiles)
    // lives in region g % nreg, so whenever nreg > nsub the next tile starts while this one is being drained
    // Accumulator split: a region holds `accs` accumulators of BN columns, [main 0 | ... | main nmain-1 | corr]
    const int nmain = w.nmain, accs = w.nmain + w.corr, RW = BN * accs;
    const int nreg = TMEM_COLS / RW < MAX_REG ? TMEM_COLS / RW : MAX_REG;

    // 1024-byte alignment for the 128-byte swizzle atoms, as an offset so the pointers stay in the shared window
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t sA = smem_u32(smem);                    // NA slots x (hi image, lo image), 16 KB each
    const uint32_t b_img = (uint32_t)BN * 128u, b_stage = 2u * b_img;
    const uint32_t sB = sA + NA * A_SLOT;                  // NB slots x (hi image, lo image), BN x 128 B each
    uint8_t* tail = smem + NA * A_SLOT + (size_t)NB * b_stage;
    float* stg_all = reinterpret_cast<float*>(tail);
    typename EP::Row* ectx_all = reinterpret_cast<typename EP::Row*>(tail + STG_BYTES);
    int* eok_all = reinterpret_cast<int*>(tail + STG_BYTES + CTX_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(tail + STG_BYTES + CTX_BYTES + EPI_WARPS * 32 * 4);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NBARS);
    const uint32_t a_full = smem_u32(bars), a_empty = a_full + 8 * MAX_NA, b_full = a_empty + 8 * MAX_NA,
                   b_empty = b_full + 8 * MAX_NB, acc_full = b_empty + 8 * MAX_NB, acc_empty = acc_full + 8 * MAX_REG;

    if (warp == EPI_WARPS + PROD_WARPS) tmem_alloc(smem_u32(tmem_slot), (uint32_t)TMEM_COLS);
    if (tid == R::THREADS - 32) {
        for (int i = 0; i < NA; ++i) { mbar_init(a_full + 8 * i, PROD_THREADS); mbar_init(a_empty + 8 * i, 1); }
        for (int i = 0; i < MAX_NB; ++i) { mbar_init(b_full + 8 * i, 1); mbar_init(b_empty + 8 * i, 1); }
        for (int i = 0; i < MAX_REG; ++i) { mbar_init(acc_full + 8 * i, 1); mbar_init(acc_empty + 8 * i, EPI_WARPS); }
        fence_barrier_init();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
#ifdef ESCB_TC_TRACE
    long long tr[4] = {0, 0, 0, 0};
    const long long tr_start = clock64();
#endif

    if (warp >= EPI_BASE && warp < EPI_BASE + EPI_WARPS) {
      const int ew = warp - EPI_BASE;                // epilogue warp index; EPI_BASE % 4 == 0, so ew & 3 == warp & 3 = TMEM lane quadrant
      if constexpr (IsAttn<EP>::value) {
        // ============================================================== epilogue: fused window attention
        // The tile's columns are [head][q | k | v][HDP] (api.cu put_qkv_heads) and its 128 rows are 8 whole windows
        // in window order, so a warp's TMEM lane quadrant holds q, k and v of two complete windows: thread =
        // (window, query token).  Per head: q and k come out of TMEM, K is staged d-major in the warp's smem tile
        // (conflict-free scalar stores; every score step reads four keys with one broadcast LDS.128), 16 scores +
        // relative-position bias + shift mask + softmax in registers, V through the same tile, and the head's
        // output leaves through it as row segments.  Arithmetic order is that of window_attn_kernel
        // (attention.py:222-241), which the unfused path and the SIMT build still use.
        constexpr int HD = EP::HD, HDP = EP::HDP, HPB = EP::HPB, OPITCH = HD | 1;
        static_assert(HD * 32 * 4 <= R::STG_WARP_BYTES && 32 * (HD + 4) * 4 <= R::STG_WARP_BYTES, "attention staging tile too small");
        const int q = ew & 3, part = ew >> 2;
        float* stg = stg_all + ew * (R::STG_WARP_BYTES / 4);
        const int wl = lane >> 4, i = lane & 15;           // which of the warp's two windows, query token
        uint32_t reg = 0, rphase = 0;
        int tile_it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tile_it) {
            const long long mrow0 = (long long)(tile / ntn) * BM + q * 32;
            const long long m = mrow0 + lane;
            float lmean = 0.f, lrstd = 0.f, lf = 0.f;      // post-GEMM LayerNorm of my row (zero-padded rows: all zero)
            if constexpr (LNP) {
                if (m < M) { const float2 t2 = __ldg(ln.stats + m); lmean = t2.x; lrstd = t2.y; lf = t2.y != 0.f ? 1.f : 0.f; }
            }
            const int nrows = M - mrow0 >= 32 ? 32 : (M - mrow0 > 0 ? (int)(M - mrow0) : 0);   // rows of this quadrant that exist
            unsigned diff = 0;                             // keys of my window that lie in another shift region than me
            if (ep.masked) {
                // region ids of the shifted map (attention.py:56-75): 0 | 1 | 2 along each axis, id = 3*rh + rw
                const unsigned wi = (unsigned)(m >> 4);
                const int win = (int)(wi - ep.dW.div(wi) * (unsigned)ep.nW);
                const int wh = (int)ep.dWw.div((unsigned)win), ww = win - wh * ep.nWw;
                int rh[4], rw[4];
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    const int hs = wh * 4 + a, ws = ww * 4 + a;
                    rh[a] = hs < ep.Hp - 4 ? 0 : (hs < ep.Hp - 2 ? 1 : 2);
                    rw[a] = ws < ep.Wp - 4 ? 0 : (ws < ep.Wp - 2 ? 1 : 2);
                }
                const int mine = 3 * rh[i >> 2] + rw[i & 3];
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (3 * rh[j >> 2] + rw[j & 3] != mine) diff |= 1u << j;
            }
            for (int sub = 0; sub < nsub; ++sub) {
                { TC_T0(t_); mbar_wait(acc_full + 8 * reg, rphase); TC_ACC(0, t_); }
                tc_fence_after();
                const uint32_t tbase = tmem + reg * RW + ((uint32_t)(q * 32) << 16);
                const int h0 = ((tile % ntn) * nsub + sub) * HPB;
                // heads of the sub-tile alternate between the quadrant's E warps; with an odd head count the warp
                // that takes the extra head alternates from tile to tile
                for (int hl = (HPB % E) ? (part + tile_it) % E : part; hl < HPB; hl += E) {
                    const int h = h0 + hl;
                    if (h >= ep.heads) break;               // zero-padded head slots of the last sub-tile
                    const float* bh = ep.bias + h * (3 * HDP);
                    const uint32_t th = tbase + (uint32_t)(hl * 3 * HDP);
                    unsigned long long s2[8];              // scores (2a, 2a + 1) packed for fma.rn.f32x2
                    {
                        float kv[HDP], qv[HDP];
                        tmem_ld_acc<HDP>(th + HDP, kv, accs, BN);
                        tmem_ld_acc<HDP>(th, qv, accs, BN);
                        tmem_ld_wait<HDP>(kv);
                        tmem_ld_wait<HDP>(qv);
                        if constexpr (LNP) {
                            ln_post<HDP>(kv, ln.cs + (bh - ep.bias) + HDP, ln.bw + (bh - ep.bias) + HDP, lmean, lrstd, lf);
                            ln_post<HDP>(qv, ln.cs + (bh - ep.bias), ln.bw + (bh - ep.bias), lmean, lrstd, lf);
                        }
                        add_bias<HDP>(kv, bh + HDP);
                        add_bias<HDP>(qv, bh);
                        __syncwarp();                      // the previous head's output rows have left the tile
#pragma unroll
                        for (int d = 0; d < HD; ++d) stg[d * 32 + lane] = kv[d];
                        __syncwarp();
#pragma unroll
                        for (int a = 0; a < 8; ++a) s2[a] = 0ull;
#pragma unroll
                        for (int d = 0; d < HD; ++d) {
                            const float qd = qv[d] * ep.scale;
                            const unsigned long long qq = pack2(qd, qd);
                            const float4* kp = reinterpret_cast<const float4*>(stg + d * 32 + wl * 16);
#pragma unroll
                            for (int j4 = 0; j4 < 4; ++j4) {
                                const float4 k4 = kp[j4];
                                s2[2 * j4] = ffma2(qq, pack2(k4.x, k4.y), s2[2 * j4]);
                                s2[2 * j4 + 1] = ffma2(qq, pack2(k4.z, k4.w), s2[2 * j4 + 1]);
                            }
                        }
                    }
                    float sc[16];
#pragma unroll
                    for (int a = 0; a < 8; ++a) unpack2(s2[a], sc[2 * a], sc[2 * a + 1]);
                    {
                        const float4* rb = reinterpret_cast<const float4*>(ep.relbias + (h * 16 + i) * 16);
#pragma unroll
                        for (int j4 = 0; j4 < 4; ++j4) {
                            const float4 b4 = __ldg(rb + j4);
                            sc[4 * j4 + 0] += b4.x; sc[4 * j4 + 1] += b4.y; sc[4 * j4 + 2] += b4.z; sc[4 * j4 + 3] += b4.w;
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if ((diff >> j) & 1u) sc[j] += -100.0f;
                    float mx = sc[0];
#pragma unroll
                    for (int j = 1; j < 16; ++j) mx = fmaxf(mx, sc[j]);
                    float sum = 0.f;
#pragma unroll
                    for (int j = 0; j < 16; ++j) { sc[j] = expf(sc[j] - mx); sum += sc[j]; }
                    const float inv = 1.0f / sum;
#pragma unroll
                    for (int a = 0; a < 8; ++a) s2[a] = pack2(sc[2 * a] * inv, sc[2 * a + 1] * inv);
                    float o[HDP];
                    {
                        float vv[HDP];
                        tmem_ld_acc<HDP>(th + 2 * HDP, vv, accs, BN);
                        tmem_ld_wait<HDP>(vv);
                        if constexpr (LNP) ln_post<HDP>(vv, ln.cs + (bh - ep.bias) + 2 * HDP, ln.bw + (bh - ep.bias) + 2 * HDP, lmean, lrstd, lf);
                        add_bias<HDP>(vv, bh + 2 * HDP);
                        __syncwarp();                      // every lane is done with the K image
#pragma unroll
                        for (int d = 0; d < HD; ++d) stg[d * 32 + lane] = vv[d];
                        __syncwarp();
#pragma unroll
                        for (int d = 0; d < HD; ++d) {
                            // even and odd keys accumulate in the two halves of one f32x2 chain, joined at the end
                            const float4* vp = reinterpret_cast<const float4*>(stg + d * 32 + wl * 16);
                            unsigned long long acc = 0ull;
#pragma unroll
                            for (int j4 = 0; j4 < 4; ++j4) {
                                const float4 v4 = vp[j4];
                                acc = ffma2(s2[2 * j4], pack2(v4.x, v4.y), acc);
                                acc = ffma2(s2[2 * j4 + 1], pack2(v4.z, v4.w), acc);
                            }
                            float e0, e1;
                            unpack2(acc, e0, e1);
                            o[d] = e0 + e1;
                        }
#pragma unroll
                        for (int d = HD; d < HDP; ++d) o[d] = 0.f;
                    }
                    __syncwarp();                          // every lane is done with the V image
                    // the head's 32 x HD outputs leave as row segments: LPR lanes per row, 32 / LPR rows per pass
                    float* op = ep.out + mrow0 * (long long)ep.ldo + h * HD;
                    if constexpr (HD % 4 == 0) {
                        constexpr int VP = HD + 4, LPR = HD / 4 <= 2 ? 2 : (HD / 4 <= 4 ? 4 : 8), RPP = 32 / LPR;
#pragma unroll
                        for (int c = 0; c < HD / 4; ++c)
                            *reinterpret_cast<float4*>(stg + lane * VP + 4 * c) = make_float4(o[4 * c], o[4 * c + 1], o[4 * c + 2], o[4 * c + 3]);
                        __syncwarp();
                        const int c = lane % LPR, r0 = lane / LPR;
                        if (c < HD / 4) {
                            float* pr = op + (long long)r0 * ep.ldo + 4 * c;
                            for (int r = r0; r < nrows; r += RPP, pr += (long long)RPP * ep.ldo)
                                *reinterpret_cast<float4*>(pr) = *reinterpret_cast<const float4*>(stg + r * VP + 4 * c);
                        }
                    } else {
                        constexpr int LPR = HD <= 8 ? 8 : (HD <= 16 ? 16 : 32), RPP = 32 / LPR;
#pragma unroll
                        for (int d = 0; d < HD; ++d) stg[lane * OPITCH + d] = o[d];
                        __syncwarp();
                        const int d = lane % LPR, r0 = lane / LPR;
                        if (d < HD) {
                            float* pr = op + (long long)r0 * ep.ldo + d;
                            for (int r = r0; r < nrows; r += RPP, pr += (long long)RPP * ep.ldo) *pr = stg[r * OPITCH + d];
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(acc_empty + 8 * reg);
                if (++reg == (uint32_t)nreg) { reg = 0; rphase ^= 1; }
            }
        }
      } else {
        // ======================================================================================== epilogue
        const int q = ew & 3, part = ew >> 2;              // TMEM lane quadrant, column share (of E)
        float* stg = stg_all + ew * (R::STG_WARP_BYTES / 4);
        typename EP::Row* ectx = ectx_all + ew * 32;
        int* eok = eok_all + ew * 32;
        const int nch = BN >> 4, N = w.N;
        const int rr = lane >> 2, c4 = lane & 3;           // transposed side: rows rr + 8 i, 16-byte chunk c4
        uint32_t reg = 0, rphase = 0;                      // accumulator region of the next sub-tile and its phase
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const long long m = (long long)(tile / ntn) * BM + q * 32 + lane;
            const int n0 = (tile % ntn) * NT;
            {
                typename EP::Row er;
                int ok = m < M;
                if (ok) ok = ep.row(m, er) ? 1 : 0;
                __syncwarp();
                ectx[lane] = er;
                eok[lane] = ok;
                __syncwarp();
            }
            if (part == 0 && pf_dist > 0 && tile + pf_dist * (int)gridDim.x < ntiles) {   // residual rows of a later tile -> L2
                const long long mn = (long long)((tile + pf_dist * gridDim.x) / ntn) * BM + q * 32 + lane;
                typename EP::Row pr;
                if (mn < M && ep.row(mn, pr)) ep.prefetch(pr, N);
            }
            typename EP::Row cr[4];                        // contexts of this lane's four transposed-side rows
            unsigned okm = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                cr[i] = ectx[rr + 8 * i];
                if (eok[rr + 8 * i]) okm |= 1u << i;
            }
            float2 lst[LNP ? 4 : 1];                       // post-GEMM LayerNorm: (mean, rstd) of those rows
            if constexpr (LNP) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const long long mr = (long long)(tile / ntn) * BM + q * 32 + rr + 8 * i;
                    lst[i] = mr < M ? __ldg(ln.stats + mr) : make_float2(0.f, 0.f);
                }
            }
            for (int sub = 0; sub < nsub; ++sub) {
                // bias and residual of this lane's four row segments are requested before the accumulator is read, so
                // their latency (L2 hits: the rows were prefetched a tile ahead) overlaps the staging - and for the
                // warp's FIRST chunk of a sub-tile before the wait for the accumulator itself (ESCB_EPI_NOPRE: after it)
                float4 b4 = zero4(), res[4], cs4 = zero4(), bw4 = zero4();
                auto request = [&](int ch) {
                    const int n = n0 + sub * BN + ch * 16 + c4 * 4;
                    if (n + 3 < N) {
                        if constexpr (LNP) { cs4 = ldg4(ln.cs + n); bw4 = ldg4(ln.bw + n); }
                        b4 = ep.bias4(n);
#pragma unroll
                        for (int i = 0; i < 4; ++i) res[i] = ((okm >> i) & 1u) ? ep.resid4(cr[i], n) : zero4();
                    }
                };
#ifndef ESCB_EPI_NOPRE
                if (part < nch) request(part);
#endif
                { TC_T0(t_); mbar_wait(acc_full + 8 * reg, rphase); TC_ACC(0, t_); }
                tc_fence_after();
                const uint32_t tbase = tmem + reg * RW + ((uint32_t)(q * 32) << 16);
                for (int ch = part; ch < nch; ch += E) {
                    const int n = n0 + sub * BN + ch * 16 + c4 * 4;
                    const bool full4 = n + 3 < N;
#ifndef ESCB_EPI_NOPRE
                    if (ch != part) request(ch);
#else
                    request(ch);
#endif
                    float v[16];
                    tmem_ld_acc<16>(tbase + (uint32_t)(ch * 16), v, accs, BN);
                    tmem_ld_wait<16>(v);
                    __syncwarp();                          // previous chunk's reads of the staging tile are done
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        *reinterpret_cast<float4*>(stg + lane * 16 + ((j ^ ((lane >> 1) & 3)) << 2)) =
                            make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    __syncwarp();
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        if (!((okm >> i) & 1u) || n >= N) continue;
                        const int r = rr + 8 * i;
                        float4 o = *reinterpret_cast<const float4*>(stg + r * 16 + ((c4 ^ ((r >> 1) & 3)) << 2));
                        if (full4) {
                            if constexpr (LNP) {
                                const float mean = lst[i].x, rstd = lst[i].y, f = rstd != 0.f ? 1.f : 0.f;
                                o.x = fmaf(rstd, fmaf(-mean, cs4.x, o.x), f * bw4.x);
                                o.y = fmaf(rstd, fmaf(-mean, cs4.y, o.y), f * bw4.y);
                                o.z = fmaf(rstd, fmaf(-mean, cs4.z, o.z), f * bw4.z);
                                o.w = fmaf(rstd, fmaf(-mean, cs4.w, o.w), f * bw4.w);
                            }
                            o.x += b4.x; o.y += b4.y; o.z += b4.z; o.w += b4.w;
                            ep.fin4(cr[i], n, o, res[i]);
                        } else {
                            if constexpr (LNP) {
                                const float mean = lst[i].x, rstd = lst[i].y, f = rstd != 0.f ? 1.f : 0.f;
                                o.x = fmaf(rstd, fmaf(-mean, __ldg(ln.cs + n), o.x), f * __ldg(ln.bw + n));
                                if (n + 1 < N) o.y = fmaf(rstd, fmaf(-mean, __ldg(ln.cs + n + 1), o.y), f * __ldg(ln.bw + n + 1));
                                if (n + 2 < N) o.z = fmaf(rstd, fmaf(-mean, __ldg(ln.cs + n + 2), o.z), f * __ldg(ln.bw + n + 2));
                            }
                            ep.store(cr[i], n, o.x);
                            if (n + 1 < N) ep.store(cr[i], n + 1, o.y);
                            if (n + 2 < N) ep.store(cr[i], n + 2, o.z);
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(acc_empty + 8 * reg);
                if (++reg == (uint32_t)nreg) { reg = 0; rphase ^= 1; }
            }
        }
      }
    } else if (warp >= PROD_BASE && warp < PROD_BASE + PROD_WARPS) {
        // ======================================================================================== A producer
        // 16-byte chunk c of the 128-byte K-block row, rows r0 and r0 + 64 (8 consecutive lanes share a row).  The
        // (tile, K block) jobs of this CTA form one flat stream; the loads of job j + DEPTH are issued right after
        // job j is converted, so DEPTH K blocks of global loads (plus the L2 prefetch of the following tile) are
        // always in flight and tile boundaries cost nothing.  LayerNorm statistics come from ln_stats_kernel.
        const int pt = tid - PROD_BASE * 32;
        const int c = pt & 7, r0 = pt >> 3;
        constexpr bool RAGGED = RPT * ROW_STEP != BM;      // 12 producer warps: row r0 + 2 * 48 exists only for r0 < 32
        struct Job { float4 a[RPT]; int k; unsigned m0, vm; };
        typename AL::Row myrow[RPT];
        unsigned cur_vm = 0, cur_m0 = 0;
        int ld_tile = blockIdx.x, ld_kb = 0;

        auto issue = [&](Job& j) -> bool {
            if (ld_tile >= ntiles) return false;
            if (ld_kb == 0) {
                TC_T0(ti_);
                cur_m0 = (unsigned)(ld_tile / ntn) * BM;
                cur_vm = 0;
#pragma unroll
                for (int i = 0; i < RPT; ++i) {
                    const bool in_tile = !RAGGED || r0 + ROW_STEP * i < BM;
                    al.init(in_tile ? (long long)cur_m0 + r0 + ROW_STEP * i : M, M, myrow[i]);      // m = M: an invalid row
                    if (al.valid(myrow[i])) cur_vm |= 1u << i;
                }
                const int next = ld_tile + pf_dist * gridDim.x;
                if (pf_dist > 0 && next < ntiles) {                       // pull the following tile's rows into L2 (four threads per row)
                    typename AL::Row pr;
#pragma unroll
                    for (int rr = pt >> 2; rr < BM; rr += PROD_THREADS / 4) {
                        al.init((long long)(next / ntn) * BM + rr, M, pr);
                        if (al.valid(pr)) al.prefetch(pr, K, pt & 3);
                    }
                }
                TC_ACC(3, ti_);
            }
            const int k = ld_kb * KB + c * 4;
            j.k = k;
            j.vm = cur_vm;
            j.m0 = cur_m0;
#pragma unroll
            for (int i = 0; i < RPT; ++i)
                // plain GEMMs issue the raw load and mask at conversion time; the LayerNorm GEMMs measured slower that
                // way (their rows were just read by ln_stats_kernel and arrive from L2 almost at once), so they keep load4
                j.a[i] = (((cur_vm >> i) & 1u) && k < K) ? (LN ? al.load4(myrow[i], k, K) : al.load4_raw(myrow[i], k, K)) : zero4();
            if (++ld_kb == nkb) { ld_kb = 0; ld_tile += gridDim.x; }
            return true;
        };

        uint32_t slot = 0, phase = 0;                     // ring position of the next K block
        auto convert = [&](const Job& j) {
            const int k = j.k;
            float4 g = zero4(), be = zero4();
            float2 st[RPT];
#pragma unroll
            for (int i = 0; i < RPT; ++i) st[i] = make_float2(0.f, 0.f);
            if (LN && k < K) {
                g = ldg4(ln.gamma + k);
                be = ldg4(ln.beta + k);
#pragma unroll
                for (int i = 0; i < RPT; ++i)
                    if ((j.vm >> i) & 1u) st[i] = __ldg(ln.stats + (j.m0 + r0 + ROW_STEP * i));
            }
            { TC_T0(t_); mbar_wait(a_empty + 8 * slot, phase ^ 1); TC_ACC(0, t_); }
            uint8_t* dst = smem + slot * A_SLOT;
#pragma unroll
            for (int i = 0; i < RPT; ++i) {
                const int r = r0 + ROW_STEP * i;
                if (RAGGED && r >= BM) continue;
                float4 v = j.a[i];
                if (AL::kRawMask && !LN) v = mask4(v, k, K);      // (the LN branch masks after normalising)
                if (LN && k < K && ((j.vm >> i) & 1u)) {
                    const float mean = st[i].x, rstd = st[i].y;
                    v.x = (v.x - mean) * rstd * g.x + be.x;
                    v.y = (v.y - mean) * rstd * g.y + be.y;
                    v.z = (v.z - mean) * rstd * g.z + be.z;
                    v.w = (v.w - mean) * rstd * g.w + be.w;
                    v = mask4(v, k, K);
                }
                float4 hi, lo;
                split4(v, hi, lo);
                const uint32_t off = (uint32_t)r * 128u + ((uint32_t)(c ^ (r & 7)) << 4);
                *reinterpret_cast<float4*>(dst + off) = hi;
                *reinterpret_cast<float4*>(dst + BM * 128 + off) = lo;
            }
            fence_proxy_async();
            mbar_arrive(a_full + 8 * slot);                // per thread: a warp-level arrival would hold every lane's next loads behind the slowest one
            if (++slot == NA) { slot = 0; phase ^= 1; }
        };

        Job jobs[DEPTH];
        bool have[DEPTH];
#pragma unroll
        for (int d = 0; d < DEPTH; ++d) have[d] = issue(jobs[d]);
        while (have[0]) {
#pragma unroll
            for (int d = 0; d < DEPTH; ++d) {
                if (have[d]) {
                    { TC_T0(t_); convert(jobs[d]); TC_ACC(1, t_); }
                    { TC_T0(t_); have[d] = issue(jobs[d]); TC_ACC(2, t_); }
                }
            }
        }
    } else if (warp == EPI_WARPS + PROD_WARPS) {
        // ======================================================================================== MMA issuer
        const uint32_t idesc = make_idesc(BN);
        uint32_t aslot = 0, aphase = 0, bslot = 0, bphase = 0, reg0 = 0, rphase0 = 0;   // reg0: region of sub-tile 0
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            for (int kb = 0; kb < nkb; ++kb) {
                { TC_T0(t_); mbar_wait(a_full + 8 * aslot, aphase); TC_ACC(0, t_); }
                int rem = K - kb * KB;
                if (rem > KB) rem = KB;
                const int ksteps = (rem + 7) >> 3;
                const uint64_t a_hi = make_desc(sA + aslot * A_SLOT), a_lo = make_desc(sA + aslot * A_SLOT + BM * 128);
                const int mj = kb % nmain;
                uint32_t reg = reg0, rphase = rphase0;
                for (int sub = 0; sub < nsub; ++sub) {
                    if (kb == 0) { TC_T0(t_); mbar_wait(acc_empty + 8 * reg, rphase ^ 1); TC_ACC(1, t_); }   // the epilogue drained this region
                    const uint32_t bs = resident ? (uint32_t)(kb * nsub + sub) : bslot;
                    { TC_T0(t_); mbar_wait(b_full + 8 * bs, resident ? 0u : bphase); TC_ACC(2, t_); }
                    tc_fence_after();
                    TC_T0(ti_);
                    if (elect_one()) {
                        const uint64_t b_hi = make_desc(sB + bs * b_stage), b_lo = make_desc(sB + bs * b_stage + b_img);
                        // accumulator split: hi*hi of this K block -> main kb % nmain; corrections -> corr (or the same main)
                        const uint32_t d_main = tmem + reg * RW + (uint32_t)(mj * BN);
                        const uint32_t d_corr = w.corr ? tmem + reg * RW + (uint32_t)(nmain * BN) : d_main;
                        const bool fresh_main = kb < nmain, fresh_corr = w.corr ? kb == 0 : fresh_main;
                        for (int ks = 0; ks < ksteps; ++ks) {
                            const uint64_t adv = (uint64_t)(ks * 2);      // 32 bytes >> 4 inside the swizzle row
                            umma_tf32(d_corr, a_lo + adv, b_hi + adv, idesc, (fresh_corr && ks == 0) ? 0u : 1u);
                            umma_tf32(d_corr, a_hi + adv, b_lo + adv, idesc, 1u);
                            umma_tf32(d_main, a_hi + adv, b_hi + adv, idesc, (w.corr && fresh_main && ks == 0) ? 0u : 1u);
                        }
                        if (!resident) umma_commit(b_empty + 8 * bs);
                        if (kb + 1 == nkb) umma_commit(acc_full + 8 * reg);
                        if (sub + 1 == nsub) umma_commit(a_empty + 8 * aslot);
                    }
                    __syncwarp();
                    TC_ACC(3, ti_);
                    if (!resident && ++bslot == (uint32_t)NB) { bslot = 0; bphase ^= 1; }
                    if (++reg == (uint32_t)nreg) { reg = 0; rphase ^= 1; }
                }
                if (++aslot == NA) { aslot = 0; aphase ^= 1; }
                if (kb + 1 == nkb) { reg0 = reg; rphase0 = rphase; }
            }
        }
    } else {
        // ======================================================================================== weight loader
        if (elect_one()) {
            const uint8_t* wimg = (const uint8_t*)w.img;
            if (resident) {
                if ((int)blockIdx.x < ntiles) {
                    const int nt = blockIdx.x % ntn;          // grid is a multiple of ntn: fixed n-tile per CTA
                    for (int j = 0; j < nkb * nsub; ++j) {         // stage j = kb * nsub + sub
                        mbar_expect_tx(b_full + 8 * j, b_stage);
                        bulk_g2s(sB + j * b_stage, wimg + ((size_t)nt * nkb * nsub + j) * b_stage, b_stage, b_full + 8 * j);
                    }
                }
            } else {
                uint32_t bslot = 0, bphase = 0;
                for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                    const int nt = tile % ntn;
                    for (int j = 0; j < nkb * nsub; ++j) {
                        { TC_T0(t_); mbar_wait(b_empty + 8 * bslot, bphase ^ 1); TC_ACC(0, t_); }
                        mbar_expect_tx(b_full + 8 * bslot, b_stage);
                        bulk_g2s(sB + bslot * b_stage, wimg + ((size_t)nt * nkb * nsub + j) * b_stage, b_stage, b_full + 8 * bslot);
                        if (++bslot == (uint32_t)NB) { bslot = 0; bphase ^= 1; }
                    }
                }
            }
        }
        __syncwarp();
    }

#ifdef ESCB_TC_TRACE
    if (ln.trace && lane == 0) {
        // slots: 0 role total | epilogue: 1 wait acc_full | producer: 2 wait a_empty, 3 convert (incl. wait), 4 issue
        //        | mma: 5 wait a_full, 6 wait acc_empty, 7 wait b_full | loader: 8 wait b_empty | 9.. role totals
        const long long total = clock64() - tr_start;
        unsigned long long* t = ln.trace;
        if (warp == EPI_BASE) { atomicAdd(t + 1, (unsigned long long)tr[0]); atomicAdd(t + 9, (unsigned long long)total); atomicAdd(t + 13, 1ull); }
        if (warp == PROD_BASE) { atomicAdd(t + 2, (unsigned long long)tr[0]); atomicAdd(t + 3, (unsigned long long)tr[1]); atomicAdd(t + 4, (unsigned long long)tr[2]); atomicAdd(t + 10, (unsigned long long)total); atomicAdd(t + 12, (unsigned long long)tr[3]); }
        if (warp == EPI_WARPS + PROD_WARPS) { atomicAdd(t + 5, (unsigned long long)tr[0]); atomicAdd(t + 6, (unsigned long long)tr[1]); atomicAdd(t + 7, (unsigned long long)tr[2]); atomicAdd(t + 0, (unsigned long long)tr[3]); atomicAdd(t + 11, (unsigned long long)total); }
        if (warp == EPI_WARPS + PROD_WARPS + 1) { atomicAdd(t + 8, (unsigned long long)tr[0]); }
        if (tid == 0 && blockIdx.x == 0) { t[14] = (unsigned long long)ntiles; t[15] = ((unsigned long long)w.N << 40) | ((unsigned long long)w.K << 20) | ((unsigned long long)w.BN << 8) | ((unsigned long long)w.nsub << 4) | (unsigned long long)w.resident; }
    }
#endif
    tc_fence_before();
    __syncthreads();
    if (warp == EPI_WARPS + PROD_WARPS) {
        tc_fence_after();
        tmem_dealloc(tmem, (uint32_t)TMEM_COLS);
    }
}

// LayerNorm statistics of the logical A rows (mean, 1/sqrt(var + eps)): 8 lanes per row, the row held in
// registers (K <= 384: 12 float4 per lane), two-pass (mean, then centred squares) like the reference's
// layer_norm with a single read of the data and every load independent.
constexpr int LN_MAX_K = 384;
template <int NCH, class AL>
__global__ void __launch_bounds__(256)
ln_stats_kernel(const AL al, const long long M, const int K, const float eps, float2* __restrict__ out) {
    const int c = threadIdx.x & 7;
    const long long m = (long long)blockIdx.x * 32 + (threadIdx.x >> 3);
    typename AL::Row r;
    al.init(m, M, r);                                      // m >= M yields an invalid row
    const bool ok = al.valid(r);
    float4 v[NCH];
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
        const int k = (c + 8 * j) * 4;
        v[j] = (ok && k < K) ? al.load4(r, k, K) : zero4();
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < NCH; ++j) s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    const float mean = s / (float)K;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
        const int k = (c + 8 * j) * 4;
        if (k < K) { const float d = v[j].x - mean; q = fmaf(d, d, q); }
        if (k + 1 < K) { const float d = v[j].y - mean; q = fmaf(d, d, q); }
        if (k + 2 < K) { const float d = v[j].z - mean; q = fmaf(d, d, q); }
        if (k + 3 < K) { const float d = v[j].w - mean; q = fmaf(d, d, q); }
    }
    q += __shfl_xor_sync(0xffffffffu, q, 1);
    q += __shfl_xor_sync(0xffffffffu, q, 2);
    q += __shfl_xor_sync(0xffffffffu, q, 4);
    if (c == 0 && m < M) out[m] = ok ? make_float2(mean, 1.0f / sqrtf(q / (float)K + eps)) : make_float2(0.f, 0.f);
}

template <class AL>
inline cudaError_t launch_ln_stats(cudaStream_t st, const AL& al, long long M, int K, float eps, float2* out) {
    const unsigned grid = (unsigned)((M + 31) / 32);
    if (K <= 64) ln_stats_kernel<2, AL><<<grid, 256, 0, st>>>(al, M, K, eps, out);
    else if (K <= 96) ln_stats_kernel<3, AL><<<grid, 256, 0, st>>>(al, M, K, eps, out);
    else if (K <= 192) ln_stats_kernel<6, AL><<<grid, 256, 0, st>>>(al, M, K, eps, out);
    else if (K <= LN_MAX_K) ln_stats_kernel<LN_MAX_K / 32, AL><<<grid, 256, 0, st>>>(al, M, K, eps, out);
    else return cudaErrorInvalidValue;
    return cudaGetLastError();
}

// Tiling of one weight, decided at pack time (api.cu put_tc): N is cut into ntn output tiles of nsub sub-tiles of
// BN columns each (BN = the UMMA N and the granularity of the weight ring; nsub * BN <= 512 TMEM columns).  The
// A operand is produced once per output tile (~100 columns worth of tensor time per K block), so the chooser
// takes the fewest n-tiles, then the least column padding, preferring resident weights and a >= 3 slot ring.
struct Tiling { int BN, nsub, ntn, nkb, resident, nmain, corr; };

// Accumulator split of a GEMM with reduction length K.  tcgen05.mma truncates its fp32 accumulator toward zero (about
// 0.6 ulp of the accumulator per MMA, profiles/r2_microbench_mma_acc.txt), so the error of a dot product grows with the
// number of MMAs chained into one accumulator: with everything in one accumulator (3 K / 8 MMAs) a 3xTF32 product is
// 3x (K = 48) to 15x (K = 1536) less accurate than an fp32 FMA chain, and 12 of 288 bench clips had a code decision
// that differs from the reference's (the fp32 SIMT engine: 3, all split policies: 3-4; profiles/r2_acc_sweep*.txt).
// The lo*hi + hi*lo corrections are ~2^-11 of the result: in an accumulator of their own their truncation is invisible
// and the main chain shrinks to K / 8; `nmain` main accumulators taking alternate K blocks cut it to K / (8 nmain) on
// values ~1/sqrt(nmain) as large (profiles/r2_microbench_mma_split.txt).  Policy (`nmain` = what must fit, `want` = what
// is taken when it costs neither an extra n-tile nor the last double-buffered region): reductions up to 96 keep one
// accumulator (chains <= 36 MMAs, the three top levels, where TMEM pays for tile overlap), longer ones split off the
// corrections, K >= 1024 also gets a second main.  ESCB_ACC="nmain,corr" / ESCB_ACC_KMIN override it at pack time.
struct AccSplit { int nmain, corr, want; };
inline AccSplit acc_policy(int K, int max_accs, bool attn = false) {
    AccSplit a{1, 0, 1};
    const int nkb = (K + KB - 1) / KB;
    // The fused attention kernel's sub-tiles are 144 columns: a second accumulator leaves one region and one sub-tile
    // per output tile (A produced 2-4x as often, MMAs serialised with the attention epilogue: +22 % at C = 144, +42 % at
    // C = 192), so it keeps one accumulator up to K = 192 (chains <= 72 MMAs, ~1.3e-6 rms) and splits at C = 384.
    int kmin = attn ? 193 : 97;
    if (const char* e = getenv(attn ? "ESCB_ACC_KMIN_ATTN" : "ESCB_ACC_KMIN")) kmin = atoi(e);
    if (const char* e = getenv("ESCB_ACC")) {
        int m = 1, c = 0;
        if (sscanf(e, "%d,%d", &m, &c) >= 1) { a.nmain = m < 1 ? 1 : (m > 4 ? 4 : m); a.corr = c ? 1 : 0; a.want = a.nmain; }
    } else {
        a.corr = 1;
        a.nmain = K >= 1024 ? 2 : 1;
        a.want = K >= 1024 ? 3 : (K >= 256 ? 2 : 1);
    }
    if (K < kmin) a = AccSplit{1, 0, 1};
    if (a.nmain > nkb) a.nmain = nkb;
    while (a.nmain + a.corr > max_accs && a.nmain > 1) --a.nmain;
    if (a.nmain + a.corr > max_accs) a.corr = 0;
    if (a.want > nkb) a.want = nkb;
    if (a.want + a.corr > max_accs) a.want = max_accs - a.corr;
    if (a.want < a.nmain) a.want = a.nmain;
    return a;
}

inline Tiling choose_tiling(int N, int K, int wide, int max_accs = 4) {
    const long long B_BUDGET = b_budget(wide);
    const AccSplit as = acc_policy(K, max_accs);
    const int accs = as.nmain + as.corr;
    Tiling best{0, 0, 0, 0, 0, 1, 0};
    double best_cost = 1e30;
    const int nkb = (K + KB - 1) / KB;
    for (int ntn = 1; ntn <= 64; ++ntn)
        for (int nsub = 1; nsub <= MAX_REG; ++nsub) {
            const int bn = (((N + ntn * nsub - 1) / (ntn * nsub)) + 15) / 16 * 16;
            if (bn > MAX_BN || bn * nsub * accs > TMEM_COLS) continue;
            if (bn < 48 && ntn * nsub > 1) continue;
            const long long stage = (long long)bn * 256;
            const bool res = stage * nkb * nsub <= B_BUDGET && nkb * nsub <= MAX_NB;
            const int nb = res ? nkb * nsub : (int)(B_BUDGET / stage);
            if (!res && nb < 2) continue;
            const double padn = (double)bn * nsub * ntn;
            const int nreg = TMEM_COLS / (bn * accs) < MAX_REG ? TMEM_COLS / (bn * accs) : MAX_REG;
            // an SS-mode MMA reads (128 + bn) * 32 bytes of shared memory per bn / 2 clocks: wider is cheaper per MAC
            const double narrow = ntn * nsub > 1 ? (bn < 96 ? 0.3 : (bn < 128 ? 0.15 : (bn < 176 ? 0.05 : 0.0))) : 0.0;
            // accumulator regions: with nreg == nsub the MMAs of the next tile wait for the epilogue of this one (all of
            // it when the tile is ONE sub-tile, the first 1 / nsub of it otherwise); full overlap needs nreg >= 2 nsub
            static const double serial_w = getenv("ESCB_TILE_SERIAL") ? atof(getenv("ESCB_TILE_SERIAL")) : 0.15;
            const double serial = nreg >= 2 * nsub ? 0.0 : (nreg > nsub ? 0.5 * serial_w : (accs > 1 ? 2.0 * serial_w / nsub : serial_w));
            const double cost = padn * (1.0 + (res ? 0.0 : 0.15) + serial + narrow) +
                                96.0 * ntn + 4.0 * nsub;
            if (cost < best_cost) { best_cost = cost; best = Tiling{bn, nsub, ntn, nkb, res ? 1 : 0, as.nmain, as.corr}; }
        }
    // free upgrades: more main accumulators while the tile still fits TMEM and keeps its region ring (or had none)
    while (best.BN > 0 && best.nmain < as.want && best.BN * best.nsub * (best.nmain + 1 + best.corr) <= TMEM_COLS) {
        const int nreg0 = TMEM_COLS / (best.BN * (best.nmain + best.corr)), nreg1 = TMEM_COLS / (best.BN * (best.nmain + 1 + best.corr));
        if (nreg1 < nreg0 && (nreg1 < 2 || nreg1 <= best.nsub)) break;
        ++best.nmain;
    }
    return best;
}

// Tiling of the head-major qkv weight of the fused attention GEMM: `nslots` head slots come in sub-tiles of
// kAttnBN = 144 columns (HPB whole heads); as many sub-tiles per output tile as TMEM holds (3), evenly split.
inline Tiling attn_tiling(int nsubs_total, int K, int* ntn_out_subs = nullptr) {
    const int nkb = (K + KB - 1) / KB;
    const AccSplit as = acc_policy(K, TMEM_COLS / 144, true);
    const int max_sub = TMEM_COLS / (144 * (as.nmain + as.corr));
    // fewest padded (all-zero) sub-tiles first, then the most sub-tiles per output tile (A is produced once per tile)
    int ntn = 1, nsub = 1, best_waste = 1 << 30;
    for (int ns = max_sub; ns >= 1; --ns) {
        const int nt = (nsubs_total + ns - 1) / ns, waste = nt * ns - nsubs_total;
        if (waste < best_waste) { best_waste = waste; ntn = nt; nsub = ns; }
    }
    const long long stage = 144LL * 256;
    const bool res = stage * nkb * nsub <= b_budget_attn() && nkb * nsub <= MAX_NB;
    if (ntn_out_subs) *ntn_out_subs = ntn * nsub;
    int nmain = as.nmain;                                  // free upgrade, as in choose_tiling
    while (nmain < as.want && 144 * nsub * (nmain + 1 + as.corr) <= TMEM_COLS) {
        const int nreg0 = TMEM_COLS / (144 * (nmain + as.corr)), nreg1 = TMEM_COLS / (144 * (nmain + 1 + as.corr));
        if (nreg1 < nreg0 && (nreg1 < 2 || nreg1 <= nsub)) break;
        ++nmain;
    }
    return Tiling{144, nsub, ntn, nkb, res ? 1 : 0, nmain, as.corr};
}

// Per-device caches: a process may drive several GPUs (codec.py keeps one handle per device), and both the SM count
// and cudaFuncAttributeMaxDynamicSharedMemorySize are per-device properties.
constexpr int kMaxDevices = 64;
inline int current_device() {
    int dev = 0;
    cudaGetDevice(&dev);
    return dev >= 0 && dev < kMaxDevices ? dev : 0;
}
inline int sm_count() {
    static std::atomic<int> n[kMaxDevices];
    const int dev = current_device();
    int v = n[dev].load(std::memory_order_relaxed);
    if (v == 0) {
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        n[dev].store(v, std::memory_order_relaxed);
    }
    return v;
}

template <bool LN, class AL, class EP, int E, bool LNP = false>
inline cudaError_t launch_e(cudaStream_t st, const AL& al, const LnParams& ln, const TcWeight& w, long long M, const EP& ep) {
    using R = RolesFor<E, EP>;
    static std::atomic<bool> configured[kMaxDevices];     // per instantiation and device
    const int dev = current_device();
    if (!configured[dev].load(std::memory_order_acquire)) {
        const cudaError_t e = cudaFuncSetAttribute(tc_gemm_kernel<LN, AL, EP, E, LNP>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX);
        if (e != cudaSuccess) return e;
        configured[dev].store(true, std::memory_order_release);
    }
    const long long stage = (long long)w.BN * 256;
    const int NB = w.resident ? w.nkb * w.nsub : (int)(R::B_BUDGET / stage < MAX_NB ? R::B_BUDGET / stage : MAX_NB);
    const size_t smem = 1024 + (size_t)R::NA * A_SLOT + (size_t)NB * stage + R::TAIL_BYTES;
    const long long ntm = (M + BM - 1) / BM;
    const long long ntiles = ntm * w.ntn;
    if (ntiles > 0x7fffffffLL) return cudaErrorInvalidValue;
    long long grid = sm_count();
    if (w.resident) grid = grid / w.ntn * w.ntn;       // a resident CTA serves one n-tile
    if (grid > ntiles) grid = ntiles;
    // tiles of L2 prefetch distance (ESCB_TC_PREFETCH overrides).  Measured per class at 36 clips: the fused attention
    // kernel and the long-row GEMMs (K >= 2N: mlp2) are fastest without the prefetch, the others with one tile.
    static int pf_env = -2;
    if (pf_env == -2) { const char* e = getenv("ESCB_TC_PREFETCH"); pf_env = e ? atoi(e) : -1; }
    const int pf_dist = pf_env >= 0 ? pf_env : ((IsAttn<EP>::value || w.K >= 2 * w.N) ? 0 : 1);
    tc_gemm_kernel<LN, AL, EP, E, LNP><<<(unsigned)grid, R::THREADS, smem, st>>>(al, ln, w, M, ep, (int)ntiles, NB, w.resident, pf_dist);
    return cudaGetLastError();
}

// WIDE is the role-split code (role_e) the weight was tiled with (choose_tiling(.., wide)).
// Tiling for this row count: the persistent grid runs ceil(tiles / SMs) rounds of the slowest CTA; with few row tiles
// the narrower alternative (r times the tiles, each ~1/r of the work plus its own A production) needs fewer
// round-equivalents, e.g. 169 row tiles of N = 384: 2 rounds vs 4 rounds of thirds.
inline const TcWeight& pick(const GemmWeight& gw, long long M) {
    if (!gw.tc_alt.img || !gw.tc.img) return gw.tc;
    // ESCB_TC_ALT=1 / 0 forces the narrow / default tiling (A-B debugging and the parity tests, which must be able to
    // switch it inside one process: read per call)
    if (const char* e = getenv("ESCB_TC_ALT")) { if (*e) return atoi(e) ? gw.tc_alt : gw.tc; }
    const long long ntm = (M + BM - 1) / BM, sms = sm_count();
    const double r = (double)(gw.tc_alt.ntn * gw.tc_alt.nsub) / (double)(gw.tc.ntn);      // tiles ratio
    const double cost_def = (double)((ntm * gw.tc.ntn + sms - 1) / sms);
    const double cost_alt = (double)((ntm * gw.tc_alt.ntn + sms - 1) / sms) / r * 1.10;   // +10%: A produced r times
    return cost_alt < cost_def ? gw.tc_alt : gw.tc;
}

// LNP: the LayerNorm is applied after the GEMM (gw must hold the gamma-scaled image, ln.cs / ln.bw its vectors); the
// producers then run the plain path (raw loads, no statistics / gamma / beta traffic, no normalisation arithmetic).
// stats_ready: ln.stats already holds (mean, rstd) of every valid logical row (written by the producer of the input map,
// e.g. the fused MLP kernel's epilogue): the statistics pre-kernel is skipped.
template <bool LN, class AL, class EP, int WIDE = 0, bool LNP = false>
inline cudaError_t launch(cudaStream_t st, const AL& al, const LnParams& ln, const GemmWeight& gw, long long M, const EP& ep,
                          bool stats_ready = false) {
    static_assert(!(LN && LNP), "LayerNorm is applied either in the producers or after the GEMM");
    const TcWeight& w = pick(gw, M);
    if (!w.img || M <= 0) return M <= 0 ? cudaSuccess : cudaErrorInvalidValue;
    if (M >= (1LL << 31)) return cudaErrorInvalidValue;      // loaders / epilogues use 32-bit row arithmetic
    constexpr int E = IsAttn<EP>::value ? kAttnE : role_e(WIDE);
    if (w.wide != role_code(E)) return cudaErrorInvalidValue;
    if (LNP && (!ln.cs || !ln.bw)) return cudaErrorInvalidValue;
    if (LN || LNP) {
        if (!ln.stats) return cudaErrorInvalidValue;
        if (!stats_ready) {
            const cudaError_t e = launch_ln_stats(st, al, M, w.K, ln.eps, ln.stats);
            if (e != cudaSuccess) return e;
        }
    }
    return launch_e<LN, AL, EP, E, LNP>(st, al, ln, w, M, ep);
}

}  // namespace tc
}  // namespace escb
