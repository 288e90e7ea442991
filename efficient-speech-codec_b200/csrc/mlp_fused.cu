// Kernel and launcher of the fused Swin MLP (design notes: mlp_fused.h).
#include <cuda.h>

#include <atomic>

#include "mlp_fused.h"

namespace escb {
namespace mf {

struct Params {
    int C, ld;                      // channels, row pitch of the token map (floats)
    int Kp16;                       // C rounded up to 16: columns of one A1 image in TMEM
    int ksteps1, nkb1;              // fc1: k-steps of 8 (ceil(C / 8)) and 32-wide K blocks
    int nch;                        // hidden chunks: ceil(4C / 64)
    int N2;                         // C rounded up to 16: UMMA N of fc2 / ACC2 columns
    int nx, na1, nl, nacc;          // ring depths: x slots, A1 buffers, L buffers, ACC2 buffers
    int resident, ns1, ns2;         // weights resident in smem | ring slots of the fc1 / fc2 rings when streamed
    int nboxf, rem;                 // x tile = nboxf full 32-column boxes + a remainder of `rem` columns (dense rows)
    unsigned st2_bytes, slot_bytes, chunk_bytes, xslot_bytes;
    int col_a1, col_r, col_l, col_acc;
    int corr2;                      // 1: fc2 corrections accumulate at col_acc + (nacc + ab) * N2
    const float* w_img;             // per chunk: fc1 stages kb = 0..nkb1-1, then fc2 stages kb = 0, 1
    const float* b1;                // [nch * 64], zero padded
    const float* b2;                // [N2]
    const float* gamma;             // [Kp16]
    const float* beta;
    float eps;
    long long M;
    int ntiles;
    // optional: LayerNorm statistics of the OUTPUT rows for the next consumer (null: none).  stat_geom != 0: the row's
    // index in the window order of `ng` (the next SwinBlock's partition); 0: the token index.
    int dbg;                        // ESCB_MF_DBG experiments (results wrong): 1 skip the GELU polynomial, 2 skip the tf32 splits, 4 one MMA per k-step
    unsigned long long* trace;      // ESCB_TC_TRACE builds: 16 counters of this launch (null otherwise)
    float2* stats_out;
    int stat_geom;
    int H, W;
    FastDiv dHW, dW_;
    WindowGeom ng;
};

// Parity wait of this kernel: a non-blocking probe (the common case is "already done"), then probes spaced by
// nanosleep with exponential back-off up to `cap_ns`.  Why not tc::mbar_wait / a bare try_wait loop: up to 16 of the 20
// warps are parked on a barrier at any time, and both forms re-poll at a high rate (try_wait returns after a few tens of
// clocks; the 2 us suspend-hint form is woken by every barrier event of the CTA) - in the first ncu capture the wait
// loops were 14.3k of the 29.7k warp instructions per tile and starved the MMA-issuing warp and the GELU warps of issue
// slots.  A sleeping warp issues nothing.  Traps after ~5 s.
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    return done != 0;
}
template <unsigned CAP_NS = 64>
__device__ __forceinline__ void mbar_wait_fast(uint32_t bar, uint32_t parity) {
    if (mbar_test(bar, parity)) return;
    unsigned ns = 16;
    long long t0 = 0;
    for (uint32_t spins = 0;; ++spins) {
        __nanosleep(ns);
        if (mbar_test(bar, parity)) return;
        if (ns < CAP_NS) ns *= 2;
        if ((spins & 0xfffu) == 0xfffu) {
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 10000000000LL) __trap();
        }
    }
}
// cvt.rna.tf32.f32 without its NaN / infinity handling (5 SASS instructions): round to nearest, ties away = add half an
// ulp of the 10-bit mantissa and clear the 13 low bits.  Bit-identical for finite inputs (the values here are).
__device__ __forceinline__ float tf32_rn_fast(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u); }

#ifdef ESCB_TC_TRACE
// Debug timeline of CTA 0 (trace builds): event records (id << 56 | arg << 40 | clock) of the launches whose channel count is armed
__device__ int g_tl_arm = 0;
__device__ unsigned int g_tl_n = 0;
__device__ unsigned long long g_tl[8192];
// events go to a per-thread local array (a clock read + one local store) and are copied out when the kernel ends
#define MF_TL(ev, arg) do { if (tl_on && lane == 0 && tl_n < 640) tl_buf[tl_n++] = ((unsigned long long)(ev) << 56) | ((unsigned long long)((arg) & 0xFFFF) << 40) | ((unsigned long long)clock64() & 0xFFFFFFFFFFull); } while (0)
#else
#define MF_TL(ev, arg) do { } while (0)
#endif

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, int c0, int c1, uint32_t src) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(src) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// D[tmem] (+)= A[tmem] * B[smem]^T : A is 128 lanes x 8 columns of tf32 at a_tmem
__device__ __forceinline__ void umma_ts_tf32(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float* v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
          "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])),
          "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])),
          "r"(__float_as_uint(v[11])), "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])),
          "r"(__float_as_uint(v[15])) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// Compile-time geometry of one channel width (the same arithmetic as make_plan): every per-row loop below unrolls
// completely, so the shared-memory reads of a pass are issued back to back instead of one per loop trip.
template <int C_>
struct Geo {
    static constexpr int C = C_, LD = (C_ + 3) & ~3, KP16 = (C_ + 15) & ~15, KSTEPS1 = (C_ + 7) / 8, NKB1 = (KSTEPS1 + 3) / 4;
    static constexpr int NCH = (4 * C_ + HC - 1) / HC, N2 = (C_ + 15) & ~15, NBOXF = LD / 32, REM = LD % 32, NV = LD / 4;
    static constexpr int TAIL = KSTEPS1 % 4;                     // k-steps of the last fc1 K block (0: full)
    static constexpr bool PACKED = TAIL == 1 || TAIL == 2;       // its hi and lo share one image (mlp_fused.h ST1T_BYTES)
    static constexpr unsigned ST1_LAST = PACKED ? (unsigned)ST1T_BYTES : (unsigned)ST1_BYTES;
    static constexpr unsigned FC1 = (unsigned)(NKB1 - 1) * ST1_BYTES + ST1_LAST;      // fc1 bytes of one chunk
    static constexpr unsigned ST2 = (unsigned)N2 * 256u, CHUNK = FC1 + 2u * ST2;
    static constexpr __host__ __device__ unsigned st1_bytes(int kb) { return kb + 1 == NKB1 ? ST1_LAST : (unsigned)ST1_BYTES; }
    // byte offset of the 16-byte group holding columns k..k+3 (k % 4 == 0, k < LD) of row r inside an x slot
    static __device__ __forceinline__ uint32_t x_off(int r, int k) {
        if ((k >> 5) < NBOXF) return (uint32_t)((k >> 5) * BOX_BYTES + r * 128 + ((((k & 31) >> 2) ^ (r & 7)) << 4));
        return (uint32_t)(NBOXF * BOX_BYTES + r * (REM * 4) + (k - 32 * NBOXF) * 4);
    }
};

template <int C_>
__global__ void __launch_bounds__(THREADS, 1)
mlp_fused_kernel(const __grid_constant__ CUtensorMap mapLf, const __grid_constant__ CUtensorMap mapLr,
                 const __grid_constant__ CUtensorMap mapSf, const __grid_constant__ CUtensorMap mapSr, const Params p) {
    using namespace tc;
    using G = Geo<C_>;
    constexpr int C = G::C, LD = G::LD, KP16 = G::KP16, KSTEPS1 = G::KSTEPS1, NKB1 = G::NKB1, NCH = G::NCH, N2 = G::N2;
    constexpr int NBOXF = G::NBOXF, REM = G::REM, NV = G::NV;
    constexpr unsigned ST2 = G::ST2, CHUNK = G::CHUNK;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t sX = smem_u32(smem);
    // weights: fc1 region, then fc2 region (resident: every stage of a tile at a fixed place; streamed: two rings)
    const uint32_t sW1 = sX + (uint32_t)p.nx * p.xslot_bytes;
    const uint32_t w1_bytes = p.resident ? (uint32_t)NCH * G::FC1 : (uint32_t)p.ns1 * ST1_BYTES;
    const uint32_t sW2 = sW1 + w1_bytes;
    const uint32_t w_bytes = w1_bytes + (p.resident ? (uint32_t)NCH * 2u * ST2 : (uint32_t)p.ns2 * ST2);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.nx * p.xslot_bytes + w_bytes);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NBARS);
    const uint32_t bar0 = smem_u32(bars);
    auto bar = [&](int i) -> uint32_t { return bar0 + 8u * (uint32_t)i; };

    const int my_tiles = (int)blockIdx.x < p.ntiles ? (p.ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int total = my_tiles * NCH;

    if (warp == W_ALLOC) tmem_alloc(smem_u32(tmem_slot), 512u);
    if (tid == 0) {
        for (int i = 0; i < MAX_NX; ++i) { mbar_init(bar(B_XFULL + i), 1); mbar_init(bar(B_XFREE + i), 4); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar(B_A1FULL + i), 4); mbar_init(bar(B_A1FREE + i), 1);
            mbar_init(bar(B_RFULL + i), 1); mbar_init(bar(B_HFULL + i), GELU_WARPS);
            mbar_init(bar(B_ACCFULL + i), 1); mbar_init(bar(B_ACCFREE + i), 4);
        }
        mbar_init(bar(B_LFREE), 1);
        mbar_init(bar(B_RFREE), 1); mbar_init(bar(B_RFREE + 1), 1);
        for (int i = 0; i < MAX_WST; ++i) {
            mbar_init(bar(B_W1FULL + i), 1); mbar_init(bar(B_W1FREE + i), 1);
            mbar_init(bar(B_W2FULL + i), 1); mbar_init(bar(B_W2FREE + i), 1);
        }
        fence_barrier_init();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
#ifdef ESCB_TC_TRACE
    long long tr[4] = {0, 0, 0, 0};
    const long long tr_start = clock64();
    const bool tl_on = blockIdx.x == 0 && p.C == g_tl_arm;
    unsigned long long tl_buf[640];
    int tl_n = 0;
#define MF_WAIT(slot, barid, par) do { const long long t0_ = clock64(); mbar_wait_fast(barid, par); tr[slot] += clock64() - t0_; } while (0)
#else
#define MF_WAIT(slot, barid, par) mbar_wait_fast(barid, par)
#endif

    if (warp == W_XLOAD) {
        // ================================================================================ x tile loader (TMA)
        if (elect_one()) {
            constexpr uint32_t tile_bytes = (uint32_t)(NBOXF * BOX_BYTES + BM * REM * 4);
            for (int t = 0; t < my_tiles; ++t) {
                const int use = p.nx == 2 ? t >> 1 : t / 3, s = t - use * p.nx;
                mbar_wait_fast<256>(bar(B_XFREE + s), (uint32_t)((use & 1) ^ 1));     // the store of the tile that used the slot has read it
                const int row0 = ((int)blockIdx.x + t * (int)gridDim.x) * BM;
                const uint32_t dst = sX + (uint32_t)s * p.xslot_bytes;
                mbar_expect_tx(bar(B_XFULL + s), tile_bytes);
#pragma unroll
                for (int f = 0; f < NBOXF; ++f) tma_load_2d(dst + f * BOX_BYTES, &mapLf, 32 * f, row0, bar(B_XFULL + s));
                if (REM) tma_load_2d(dst + NBOXF * BOX_BYTES, &mapLr, 32 * NBOXF, row0, bar(B_XFULL + s));
            }
        }
        __syncwarp();
    } else if (warp == W_WLOAD1 || warp == W_WLOAD2) {
        // ================================================================================ weight loaders (fc1 | fc2)
        if (elect_one() && my_tiles > 0) {
            const uint8_t* img = reinterpret_cast<const uint8_t*>(p.w_img);
            const bool fc1 = warp == W_WLOAD1;
            const int nst = fc1 ? NKB1 : 2;                                      // stages per chunk
            const uint32_t sbase = fc1 ? sW1 : sW2;
            const int bfull = fc1 ? B_W1FULL : B_W2FULL, bfree = fc1 ? B_W1FREE : B_W2FREE;
            if (p.resident) {
                for (int j = 0; j < NCH; ++j) {
                    uint32_t goff = (uint32_t)j * CHUNK + (fc1 ? 0u : G::FC1);
                    uint32_t soff = (uint32_t)j * (fc1 ? G::FC1 : 2u * ST2);
                    for (int kb = 0; kb < nst; ++kb) {
                        const uint32_t bytes = fc1 ? G::st1_bytes(kb) : ST2;
                        mbar_expect_tx(bar(bfull + j * nst + kb), bytes);
                        bulk_g2s(sbase + soff, img + goff, bytes, bar(bfull + j * nst + kb));
                        goff += bytes;
                        soff += bytes;
                    }
                }
            } else {
                const uint32_t slot_bytes = fc1 ? (uint32_t)ST1_BYTES : ST2;
                const uint32_t nslots = (uint32_t)(fc1 ? p.ns1 : p.ns2);
                uint32_t slot = 0, phase = 0;
                for (int c = 0; c < total; ++c) {
                    uint32_t goff = (uint32_t)(c % NCH) * CHUNK + (fc1 ? 0u : G::FC1);
                    for (int kb = 0; kb < nst; ++kb) {
                        const uint32_t bytes = fc1 ? G::st1_bytes(kb) : ST2;
                        mbar_wait_fast<256>(bar(bfree + slot), phase ^ 1);
                        mbar_expect_tx(bar(bfull + slot), bytes);
                        bulk_g2s(sbase + slot * slot_bytes, img + goff, bytes, bar(bfull + slot));
                        goff += bytes;
                        if (++slot == nslots) { slot = 0; phase ^= 1; }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == W_MMA1) {
        // ================================================================================ MMA issuer 1: G1(c), R[c % 2] = A1 * W1[chunk]^T
        const uint32_t idesc1 = make_idesc(HC);
        uint32_t wslot = 0, wphase = 0;
        // one K block of fc1: 3 MMAs per k-step into R (A = the LayerNorm images in TMEM)
        auto issue1 = [&](uint32_t saddr, int kb, uint32_t a_hi, uint32_t a_lo, uint32_t d) {
            // packed tail block: lo(k-step ks) sits 64 bytes (4 descriptor units) behind hi(k-step ks) in the same image
            const bool packed = G::PACKED && kb + 1 == NKB1;
            const uint64_t b_hi = make_desc(saddr), b_lo = packed ? b_hi + 4 : make_desc(saddr + HC * 128);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                if (4 * kb + ks >= KSTEPS1) break;
                const uint32_t ac = (uint32_t)((4 * kb + ks) * 8);
                const uint64_t adv = (uint64_t)(ks * 2);
                umma_ts_tf32(d, a_lo + ac, b_hi + adv, idesc1, (kb | ks) ? 1u : 0u);
                if (p.dbg & 4) continue;
                umma_ts_tf32(d, a_hi + ac, b_lo + adv, idesc1, 1u);
                umma_ts_tf32(d, a_hi + ac, b_hi + adv, idesc1, 1u);
            }
        };
        for (int c = 0; c < total; ++c) {
            const int t = c / NCH, j = c - t * NCH, a = p.na1 == 2 ? t & 1 : 0;
            MF_TL(1, c);
            if (j == 0) MF_WAIT(0, bar(B_A1FULL + a), (uint32_t)((p.na1 == 2 ? t >> 1 : t) & 1));
            MF_WAIT(1, bar(B_RFREE + (c & 1)), (uint32_t)(((c >> 1) & 1) ^ 1));      // G2(c - 2) has consumed R[c % 2] (and its lo twin)
            const uint32_t a_hi = tmem + (uint32_t)(p.col_a1 + a * 2 * KP16), a_lo = a_hi + (uint32_t)KP16;
            const uint32_t d = tmem + (uint32_t)(p.col_r + (c & 1) * HC);
            if (p.resident) {
                // the images were loaded once: only the first tile has to wait for them
                if (c < NCH)
                    for (int kb = 0; kb < NKB1; ++kb) MF_WAIT(3, bar(B_W1FULL + j * NKB1 + kb), 0u);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t base = sW1 + (uint32_t)j * G::FC1;
#pragma unroll
                    for (int kb = 0; kb < NKB1; ++kb) issue1(base + (uint32_t)kb * ST1_BYTES, kb, a_hi, a_lo, d);
                    umma_commit(bar(B_RFULL + (c & 1)));
                    if (j + 1 == NCH) umma_commit(bar(B_A1FREE + a));
                }
                __syncwarp();
            } else {
#pragma unroll
                for (int kb = 0; kb < NKB1; ++kb) {
                    MF_WAIT(3, bar(B_W1FULL + wslot), wphase);
                    tc_fence_after();
                    if (elect_one()) {
                        issue1(sW1 + wslot * ST1_BYTES, kb, a_hi, a_lo, d);
                        umma_commit(bar(B_W1FREE + wslot));
                        if (kb + 1 == NKB1) {
                            umma_commit(bar(B_RFULL + (c & 1)));
                            if (j + 1 == NCH) umma_commit(bar(B_A1FREE + a));
                        }
                    }
                    __syncwarp();
                    if (++wslot == (uint32_t)p.ns1) { wslot = 0; wphase ^= 1; }
                }
            }
            MF_TL(2, c);
        }
    } else if (warp == W_MMA2) {
        // ================================================================================ MMA issuer 2: G2(c), ACC2 += GELU chunk * W2[:, chunk]^T
        const uint32_t idesc2 = make_idesc(N2);
        uint32_t wslot = 0, wphase = 0;
        // one K block of fc2: A = the GELU chunk (hi in R, lo in L), accumulating into ACC2
        auto issue2 = [&](uint32_t saddr, int kb, uint32_t a_hi, uint32_t a_lo, uint32_t d, uint32_t dc, bool first) {
            const uint64_t b_hi = make_desc(saddr), b_lo = make_desc(saddr + (uint32_t)N2 * 128);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                const uint32_t ac = (uint32_t)((4 * kb + ks) * 8);
                const uint64_t adv = (uint64_t)(ks * 2);
                const uint32_t fresh = (first && kb == 0 && ks == 0) ? 0u : 1u;
                umma_ts_tf32(dc, a_lo + ac, b_hi + adv, idesc2, fresh);
                if (p.dbg & 4) continue;
                umma_ts_tf32(dc, a_hi + ac, b_lo + adv, idesc2, 1u);
                umma_ts_tf32(d, a_hi + ac, b_hi + adv, idesc2, p.corr2 ? fresh : 1u);
            }
        };
        for (int c = 0; c < total; ++c) {
            const int t = c / NCH, j = c - t * NCH, ab = p.nacc == 2 ? t & 1 : 0;
            MF_WAIT(1, bar(B_HFULL + (c & 1)), (uint32_t)((c >> 1) & 1));
            if (j == 0) MF_WAIT(2, bar(B_ACCFREE + ab), (uint32_t)(((p.nacc == 2 ? t >> 1 : t) & 1) ^ 1));
            const uint32_t a_hi = tmem + (uint32_t)(p.col_r + (c & 1) * HC);
            const uint32_t a_lo = tmem + (uint32_t)(p.col_l + (p.nl == 2 ? (c & 1) : 0) * HC);
            const uint32_t d = tmem + (uint32_t)(p.col_acc + ab * N2);
            const uint32_t dc = p.corr2 ? tmem + (uint32_t)(p.col_acc + (p.nacc + ab) * N2) : d;
            MF_TL(3, c);
            if (p.resident) {
                if (c < NCH)
                    for (int kb = 0; kb < 2; ++kb) MF_WAIT(3, bar(B_W2FULL + j * 2 + kb), 0u);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t base = sW2 + (uint32_t)j * 2u * ST2;
                    issue2(base, 0, a_hi, a_lo, d, dc, j == 0);
                    issue2(base + ST2, 1, a_hi, a_lo, d, dc, j == 0);
                    umma_commit(bar(B_RFREE + (c & 1)));
                    if (p.nl == 1) umma_commit(bar(B_LFREE));
                    if (j + 1 == NCH) umma_commit(bar(B_ACCFULL + ab));
                }
                __syncwarp();
            } else {
#pragma unroll
                for (int kb = 0; kb < 2; ++kb) {
                    MF_WAIT(3, bar(B_W2FULL + wslot), wphase);
                    tc_fence_after();
                    if (elect_one()) {
                        issue2(sW2 + wslot * ST2, kb, a_hi, a_lo, d, dc, j == 0);
                        umma_commit(bar(B_W2FREE + wslot));
                        if (kb == 1) {
                            umma_commit(bar(B_RFREE + (c & 1)));
                            if (p.nl == 1) umma_commit(bar(B_LFREE));
                            if (j + 1 == NCH) umma_commit(bar(B_ACCFULL + ab));
                        }
                    }
                    __syncwarp();
                    if (++wslot == (uint32_t)p.ns2) { wslot = 0; wphase ^= 1; }
                }
            }
            MF_TL(4, c);
        }
    } else if (warp >= LN_BASE && warp < LN_BASE + 4) {
        // ================================================================================ LayerNorm -> A1 (TMEM)
        const int q = warp & 3, r = q * 32 + lane;
        const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
        constexpr float invC = 1.0f / (float)C;
        for (int t = 0; t < my_tiles; ++t) {
            const int xuse = p.nx == 2 ? t >> 1 : t / 3, s = t - xuse * p.nx, a = p.na1 == 2 ? t & 1 : 0;
            MF_WAIT(0, bar(B_XFULL + s), (uint32_t)(xuse & 1));
            const uint8_t* xs = smem + (size_t)s * p.xslot_bytes;
            if (warp == LN_BASE) MF_TL(7, t);
            float sum = 0.f;
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const float4 v = mask4(*reinterpret_cast<const float4*>(xs + G::x_off(r, 4 * i)), 4 * i, C);
                sum += (v.x + v.y) + (v.z + v.w);
            }
            const float mean = sum * invC;
            float sq = 0.f;
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const float4 v = *reinterpret_cast<const float4*>(xs + G::x_off(r, 4 * i));
                { const float d = v.x - mean; sq = fmaf(d, d, sq); }
                if (4 * i + 1 < C) { const float d = v.y - mean; sq = fmaf(d, d, sq); }
                if (4 * i + 2 < C) { const float d = v.z - mean; sq = fmaf(d, d, sq); }
                if (4 * i + 3 < C) { const float d = v.w - mean; sq = fmaf(d, d, sq); }
            }
            const float rstd = 1.0f / sqrtf(sq * invC + p.eps);
            MF_WAIT(1, bar(B_A1FREE + a), (uint32_t)(((p.na1 == 2 ? t >> 1 : t) & 1) ^ 1));     // the G1s of the tile that used this buffer are done
            tc_fence_after();
            const uint32_t t_hi = lane_base + (uint32_t)(p.col_a1 + a * 2 * KP16), t_lo = t_hi + (uint32_t)KP16;
#pragma unroll
            for (int g = 0; g < KP16; g += 16) {
                float hi[16], lo[16];
#pragma unroll
                for (int c4 = 0; c4 < 4; ++c4) {
                    const int k = g + 4 * c4;
                    float4 v = zero4();
                    if (k < C) {
                        v = *reinterpret_cast<const float4*>(xs + G::x_off(r, k));
                        const float4 gm = __ldg(reinterpret_cast<const float4*>(p.gamma + k)), be = __ldg(reinterpret_cast<const float4*>(p.beta + k));
                        v.x = (v.x - mean) * rstd * gm.x + be.x;
                        v.y = (v.y - mean) * rstd * gm.y + be.y;
                        v.z = (v.z - mean) * rstd * gm.z + be.z;
                        v.w = (v.w - mean) * rstd * gm.w + be.w;
                        v = mask4(v, k, C);
                    }
                    float4 h4, l4;
                    h4.x = tf32_rn_fast(v.x); l4.x = tf32_rn_fast(v.x - h4.x);
                    h4.y = tf32_rn_fast(v.y); l4.y = tf32_rn_fast(v.y - h4.y);
                    h4.z = tf32_rn_fast(v.z); l4.z = tf32_rn_fast(v.z - h4.z);
                    h4.w = tf32_rn_fast(v.w); l4.w = tf32_rn_fast(v.w - h4.w);
                    hi[4 * c4] = h4.x; hi[4 * c4 + 1] = h4.y; hi[4 * c4 + 2] = h4.z; hi[4 * c4 + 3] = h4.w;
                    lo[4 * c4] = l4.x; lo[4 * c4 + 1] = l4.y; lo[4 * c4 + 2] = l4.z; lo[4 * c4 + 3] = l4.w;
                }
                tmem_st16(t_hi + (uint32_t)g, hi);
                tmem_st16(t_lo + (uint32_t)g, lo);
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(B_A1FULL + a));
            if (warp == LN_BASE) MF_TL(8, t);
        }
    } else if (warp >= GELU_BASE && warp < GELU_BASE + GELU_WARPS) {
        // ================================================================================ bias + GELU + split, in TMEM
        constexpr int GC = GELU_COLS;                      // hidden columns of a chunk per thread: 32 (8 warps) or 16 (16 warps)
        const int q = warp & 3, part = (warp - GELU_BASE) >> 2;
        const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
        for (int c = 0; c < total; ++c) {
            const int j = c % NCH, b = c & 1;
            // the chunk's bias slice does not depend on the accumulator: fetch it before sleeping on the barrier
            float4 bs[GC / 4];
            {
                const float4* bp = reinterpret_cast<const float4*>(p.b1 + j * HC + part * GC);
#pragma unroll
                for (int i = 0; i < GC / 4; ++i) bs[i] = __ldg(bp + i);
            }
            MF_WAIT(0, bar(B_RFULL + b), (uint32_t)((c >> 1) & 1));
            if (p.nl == 1 && c > 0) MF_WAIT(1, bar(B_LFREE), (uint32_t)((c - 1) & 1));   // G2(c - 1) has read the single lo buffer
            tc_fence_after();
            if (warp == GELU_BASE) MF_TL(5, c);
            const uint32_t t_r = lane_base + (uint32_t)(p.col_r + b * HC + part * GC);
            const uint32_t t_l = lane_base + (uint32_t)(p.col_l + (p.nl == 2 ? b : 0) * HC + part * GC);
            float v[GC];
            tmem_ld_cols<GC>(t_r, v);
            tmem_ld_wait<GC>(v);
#pragma unroll
            for (int i = 0; i < GC / 4; ++i) { v[4 * i] += bs[i].x; v[4 * i + 1] += bs[i].y; v[4 * i + 2] += bs[i].z; v[4 * i + 3] += bs[i].w; }
            if (!(p.dbg & 1)) {
#pragma unroll
                for (int i = 0; i < GC; i += 2) gelu_erf2(v[i], v[i + 1]);
            }
#pragma unroll
            for (int g = 0; g < GC / 16; ++g) {
                float lo[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float h = tf32_rn_fast(v[16 * g + i]);
                    lo[i] = tf32_rn_fast(v[16 * g + i] - h);
                    v[16 * g + i] = h;
                }
                tmem_st16(t_r + (uint32_t)(16 * g), v + 16 * g);
                tmem_st16(t_l + (uint32_t)(16 * g), lo);
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(B_HFULL + b));
            if (warp == GELU_BASE) MF_TL(6, c);
        }
    } else if (warp >= OUT_BASE && warp < OUT_BASE + 4) {
        // ================================================================================ bias + residual + store
        const int q = warp & 3, r = q * 32 + lane;
        const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
        constexpr float invC = 1.0f / (float)C;
        const bool leader = elect_one();                   // bulk async-groups are per thread: one lane issues every store and waits
        for (int t = 0; t < my_tiles; ++t) {
            const int xuse = p.nx == 2 ? t >> 1 : t / 3, s = t - xuse * p.nx, ab = p.nacc == 2 ? t & 1 : 0;
            const int row0 = ((int)blockIdx.x + t * (int)gridDim.x) * BM;
            MF_WAIT(0, bar(B_ACCFULL + ab), (uint32_t)((p.nacc == 2 ? t >> 1 : t) & 1));
            tc_fence_after();
            if (warp == OUT_BASE) MF_TL(9, t);
            uint8_t* xs = smem + (size_t)s * p.xslot_bytes;
            const uint32_t t_acc = lane_base + (uint32_t)(p.col_acc + ab * N2);
            float sum = 0.f;
            // The accumulator is copied to registers (in at most two rounds of <= 48 columns) and released to the MMA warp
            // before the residual pass: with one ACC2 buffer the first G2 of the next tile waits for this release.
            constexpr int R0 = N2 <= 48 ? N2 : N2 / 32 * 16;       // 48 | 32 (N2 = 80) | 48 (N2 = 96)
            constexpr int R1 = N2 - R0;                            //  0 | 48           | 48
            auto residual = [&](const float* v, int g0, int n) {
#pragma unroll
                for (int c4 = 0; c4 < 12; ++c4) {
                    const int k = g0 + 4 * c4;
                    if (4 * c4 < n && k < LD) {
                        const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.b2 + k));
                        float4* px = reinterpret_cast<float4*>(xs + G::x_off(r, k));
                        const float4 x4 = *px;
                        float4 o;
                        o.x = x4.x + (v[4 * c4] + b4.x); o.y = x4.y + (v[4 * c4 + 1] + b4.y);
                        o.z = x4.z + (v[4 * c4 + 2] + b4.z); o.w = x4.w + (v[4 * c4 + 3] + b4.w);
                        *px = o;
                        const float4 m4 = mask4(o, k, C);
                        sum += (m4.x + m4.y) + (m4.z + m4.w);
                    }
                }
            };
            const int accs2 = 1 + p.corr2, cstride = p.nacc * N2;      // main + corrections, cstride columns apart
            {
                float v0[R0];
                tmem_ld_acc<R0>(t_acc, v0, accs2, cstride);
                tmem_ld_wait<R0>(v0);
                if (R1 == 0) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar(B_ACCFREE + ab));
                }
                residual(v0, 0, R0);
            }
            if (R1 > 0) {
                float v1[R1 > 0 ? R1 : 16];
                tmem_ld_acc<(R1 > 0 ? R1 : 16)>(t_acc + (uint32_t)R0, v1, accs2, cstride);
                tmem_ld_wait<(R1 > 0 ? R1 : 16)>(v1);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar(B_ACCFREE + ab));
                residual(v1, R0, R1);
            }
            if (p.stats_out) {
                const float mean = sum * invC;
                float sq = 0.f;
#pragma unroll
                for (int i = 0; i < NV; ++i) {
                    const float4 v = *reinterpret_cast<const float4*>(xs + G::x_off(r, 4 * i));
                    { const float d = v.x - mean; sq = fmaf(d, d, sq); }
                    if (4 * i + 1 < C) { const float d = v.y - mean; sq = fmaf(d, d, sq); }
                    if (4 * i + 2 < C) { const float d = v.z - mean; sq = fmaf(d, d, sq); }
                    if (4 * i + 3 < C) { const float d = v.w - mean; sq = fmaf(d, d, sq); }
                }
                const long long m = (long long)row0 + r;
                if (m < p.M) {
                    long long idx = m;
                    if (p.stat_geom) {
                        // token (b, h, w) -> its row in the window order of the next block's (shifted) partition
                        const unsigned mm = (unsigned)m, bb = p.dHW.div(mm), hw = mm - bb * (unsigned)(p.H * p.W);
                        const unsigned h = p.dW_.div(hw), w = hw - h * (unsigned)p.W;
                        int hs = (int)h - p.ng.shift, ws = (int)w - p.ng.shift;
                        if (hs < 0) hs += p.ng.Hp;
                        if (ws < 0) ws += p.ng.Wp;
                        const int win = (hs >> 2) * p.ng.nWw + (ws >> 2);
                        idx = ((long long)bb * p.ng.nW + win) * 16 + (hs & 3) * 4 + (ws & 3);
                    }
                    p.stats_out[idx] = make_float2(mean, 1.0f / sqrtf(sq * invC + p.eps));
                }
            }
            fence_proxy_async();                          // the tile's rows (generic-proxy writes) -> visible to the TMA store
            __syncwarp();
            if (leader) {
                const uint32_t src = sX + (uint32_t)s * p.xslot_bytes;
#pragma unroll
                for (int f = 0; f < NBOXF; ++f) tma_store_2d(&mapSf, 32 * f, row0 + q * 32, src + f * BOX_BYTES + q * 32 * 128);
                if (REM) tma_store_2d(&mapSr, 32 * NBOXF, row0 + q * 32, src + NBOXF * BOX_BYTES + q * 32 * REM * 4);
                tma_store_commit();
                tma_store_wait_read();                    // shared memory has been read: the slot may be refilled
                mbar_arrive(bar(B_XFREE + s));
            }
            __syncwarp();
            if (warp == OUT_BASE) MF_TL(10, t);
        }
        if (leader) tma_store_wait_all();                 // global writes complete before the CTA retires
        __syncwarp();
    }

#ifdef ESCB_TC_TRACE
    if (tl_on && lane == 0 && tl_n > 0) {
        const unsigned base = atomicAdd(&g_tl_n, (unsigned)tl_n);
        for (int i = 0; i < tl_n; ++i) if (base + i < 8192u) g_tl[base + i] = tl_buf[i];
    }
    if (p.trace) {
        // slots: mma 0 total, 1 wait a1_full, 2 wait h_full, 3 wait acc_free, 4 wait w_full | ln 5 total, 6 wait x_full,
        // 7 wait a1_free | gelu 8 total, 9 wait r_full, 10 wait l_free | out 11 total, 12 wait acc_full | 13 = 0 (GEMM-engine
        // marker) | 14 CTAs | 15 signature
        const unsigned long long total_clk = (unsigned long long)(clock64() - tr_start);
        unsigned long long* t = p.trace;
        // mma columns: 1 = issuer 1's wait for A1, 2 / 3 = issuer 2's waits for the GELU chunk / the ACC2 release, 4 = both issuers' weight waits
        if (warp == W_MMA2 && lane == 0) { atomicAdd(t + 0, total_clk); atomicAdd(t + 2, (unsigned long long)tr[1]); atomicAdd(t + 3, (unsigned long long)tr[2]); atomicAdd(t + 4, (unsigned long long)tr[3]); atomicAdd(t + 14, 1ull); }
        if (warp == W_MMA1 && lane == 0) { atomicAdd(t + 1, (unsigned long long)tr[0]); atomicAdd(t + 4, (unsigned long long)tr[3]); }
        if (warp == LN_BASE && lane == 0) { atomicAdd(t + 5, total_clk); atomicAdd(t + 6, (unsigned long long)tr[0]); atomicAdd(t + 7, (unsigned long long)tr[1]); }
        if (warp == GELU_BASE && lane == 0) { atomicAdd(t + 8, total_clk); atomicAdd(t + 9, (unsigned long long)tr[0]); atomicAdd(t + 10, (unsigned long long)tr[1]); }
        if (warp == OUT_BASE && lane == 0) { atomicAdd(t + 11, total_clk); atomicAdd(t + 12, (unsigned long long)tr[0]); }
        if (tid == 0 && blockIdx.x == 0) t[15] = (0xFull << 60) | ((unsigned long long)p.ntiles << 20) | (unsigned long long)p.C;
    }
#endif
    tc_fence_before();
    __syncthreads();
    if (warp == W_ALLOC) {
        tc_fence_after();
        tmem_dealloc(tmem, 512u);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_tiled_fn() {
    static std::atomic<EncodeTiledFn> cached{nullptr};
    EncodeTiledFn fn = cached.load(std::memory_order_acquire);
    if (!fn) {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            return nullptr;
        fn = reinterpret_cast<EncodeTiledFn>(sym);
        cached.store(fn, std::memory_order_release);
    }
    return fn;
}

// fp32 [rows][ld] row-major tensor, box = box_cols x box_rows starting anywhere; swizzle 128B for the 32-column boxes
inline bool make_map(CUtensorMap* m, const float* base, long long rows, int ld, int box_cols, int box_rows, bool swizzle128) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return false;
    const cuuint64_t gdim[2] = {(cuuint64_t)ld, (cuuint64_t)rows};
    const cuuint64_t gstr[1] = {(cuuint64_t)ld * 4};
    const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

cudaError_t launch(cudaStream_t st, const Weights& w, float* x, long long M, float eps, const StatsOut& so,
                   unsigned long long* trace) {
    const Plan& pl = w.plan;
    if (!pl.ok || !w.img) return cudaErrorInvalidValue;
    if (M <= 0) return cudaSuccess;
    if (M >= (1LL << 31) - BM) return cudaErrorInvalidValue;
    static std::atomic<bool> configured[tc::kMaxDevices];
    const int dev = tc::current_device();
    if (!configured[dev].load(std::memory_order_acquire)) {
        cudaError_t e = cudaFuncSetAttribute(mlp_fused_kernel<45>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_MAX);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(mlp_fused_kernel<72>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_MAX);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(mlp_fused_kernel<96>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_MAX);
        if (e != cudaSuccess) return e;
        configured[dev].store(true, std::memory_order_release);
    }
    CUtensorMap mLf, mLr, mSf, mSr;
    const int remc = pl.rem ? pl.rem : 32;                 // unused maps still have to be valid
    const int fullc = pl.nboxf ? 32 : pl.ld;
    if (!make_map(&mLf, x, M, pl.ld, fullc, BM, pl.nboxf > 0) || !make_map(&mSf, x, M, pl.ld, fullc, 32, pl.nboxf > 0) ||
        !make_map(&mLr, x, M, pl.ld, remc, BM, false) || !make_map(&mSr, x, M, pl.ld, remc, 32, false))
        return cudaErrorInvalidValue;
    Params p;
    p.C = pl.C; p.ld = pl.ld; p.Kp16 = pl.Kp16; p.ksteps1 = pl.ksteps1; p.nkb1 = pl.nkb1; p.nch = pl.nch; p.N2 = pl.N2;
    p.nx = pl.nx; p.na1 = pl.na1; p.nl = pl.nl; p.nacc = pl.nacc; p.resident = pl.resident; p.ns1 = pl.ns1; p.ns2 = pl.ns2;
    p.nboxf = pl.nboxf; p.rem = pl.rem;
    p.st2_bytes = pl.st2_bytes; p.slot_bytes = pl.slot_bytes; p.chunk_bytes = pl.chunk_bytes; p.xslot_bytes = pl.xslot_bytes;
    p.col_a1 = pl.col_a1; p.col_r = pl.col_r; p.col_l = pl.col_l; p.col_acc = pl.col_acc; p.corr2 = pl.corr2;
    p.w_img = w.img; p.b1 = w.b1; p.b2 = w.b2; p.gamma = w.gamma; p.beta = w.beta;
    p.eps = eps;
    p.M = M;
    p.ntiles = (int)((M + BM - 1) / BM);
    p.trace = trace;
    { const char* e = getenv("ESCB_MF_DBG"); p.dbg = e ? atoi(e) : 0; }
    p.stats_out = so.out;
    p.stat_geom = so.geom;
    p.H = so.H; p.W = so.W;
    p.dHW = FastDiv::make((unsigned)(so.H > 0 ? so.H * so.W : 1));
    p.dW_ = FastDiv::make((unsigned)(so.W > 0 ? so.W : 1));
    p.ng = so.ng;
    int grid = tc::sm_count();
    if (grid > p.ntiles) grid = p.ntiles;
    switch (pl.C) {
        case 45: mlp_fused_kernel<45><<<grid, THREADS, pl.smem_bytes, st>>>(mLf, mLr, mSf, mSr, p); break;
        case 72: mlp_fused_kernel<72><<<grid, THREADS, pl.smem_bytes, st>>>(mLf, mLr, mSf, mSr, p); break;
        case 96: mlp_fused_kernel<96><<<grid, THREADS, pl.smem_bytes, st>>>(mLf, mLr, mSf, mSr, p); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace mf
}  // namespace escb

#ifdef ESCB_TC_TRACE
// Debug (trace builds): arm the CTA-0 timeline for launches with `arm_C` channels / fetch what was recorded so far.
extern "C" __attribute__((visibility("default"))) int escb_debug_timeline(int arm_C, unsigned long long* out_host, int cap) {
    unsigned n = 0;
    cudaDeviceSynchronize();
    if (out_host) {
        cudaMemcpyFromSymbol(&n, escb::mf::g_tl_n, sizeof n);
        if (n > 8192u) n = 8192u;
        if ((int)n > cap) n = (unsigned)cap;
        cudaMemcpyFromSymbol(out_host, escb::mf::g_tl, (size_t)n * 8);
    }
    const unsigned zero = 0;
    cudaMemcpyToSymbol(escb::mf::g_tl_n, &zero, sizeof zero);
    cudaMemcpyToSymbol(escb::mf::g_tl_arm, &arm_C, sizeof arm_C);
    return (int)n;
}
#endif
