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
    int resident, ns;               // weights resident in smem | ring slots when streamed
    int nboxf, rem;                 // x tile = nboxf full 32-column boxes + a remainder of `rem` columns (dense rows)
    unsigned st2_bytes, slot_bytes, chunk_bytes, xslot_bytes;
    int col_a1, col_r, col_l, col_acc;
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
    float2* stats_out;
    int stat_geom;
    int H, W;
    FastDiv dHW, dW_;
    WindowGeom ng;
};

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

// The MMA warp's op order over the CTA's chunk stream c = 0 .. total-1 (chunk c belongs to tile c / nch): G2 of a chunk
// is issued two chunks behind G1, except that with a single A1 buffer the G2s of a tile are drained before the first G1 of
// the next one (its LayerNorm cannot start until the last G1 of this tile has released A1, and the in-order issuer
// must not sit on that wait with G2 work behind it).  The weight loader walks the same order.
template <class F1, class F2>
__device__ __forceinline__ void for_each_op(int total, int nch, int na1, F1 g1, F2 g2) {
    int pend = 0;
    for (int c = 0; c < total; ++c) {
        const int limit = (na1 == 1 && c % nch == 0) ? c - 1 : c - 2;
        while (pend <= limit) g2(pend++);
        g1(c);
    }
    while (pend < total) g2(pend++);
}

// address of the 16-byte group holding columns k..k+3 (k % 4 == 0, k < ld) of row r of an x slot
__device__ __forceinline__ uint32_t x_off(const Params& p, int r, int k) {
    const int f = k >> 5;
    if (f < p.nboxf) return (uint32_t)(f * BOX_BYTES + r * 128 + ((((k & 31) >> 2) ^ (r & 7)) << 4));
    return (uint32_t)(p.nboxf * BOX_BYTES + r * (p.rem * 4) + (k - 32 * p.nboxf) * 4);
}

__global__ void __launch_bounds__(THREADS, 1)
mlp_fused_kernel(const __grid_constant__ CUtensorMap mapLf, const __grid_constant__ CUtensorMap mapLr,
                 const __grid_constant__ CUtensorMap mapSf, const __grid_constant__ CUtensorMap mapSr, const Params p) {
    using namespace tc;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t sX = smem_u32(smem);
    const uint32_t sW = sX + (uint32_t)p.nx * p.xslot_bytes;
    const uint32_t w_bytes = p.resident ? (uint32_t)p.nch * p.chunk_bytes : (uint32_t)p.ns * p.slot_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.nx * p.xslot_bytes + w_bytes);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NBARS);
    const uint32_t bar0 = smem_u32(bars);
    auto bar = [&](int i) -> uint32_t { return bar0 + 8u * (uint32_t)i; };

    const int my_tiles = (int)blockIdx.x < p.ntiles ? (p.ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int total = my_tiles * p.nch;

    if (warp == W_ALLOC) tmem_alloc(smem_u32(tmem_slot), 512u);
    if (tid == 0) {
        for (int i = 0; i < MAX_NX; ++i) { mbar_init(bar(B_XFULL + i), 1); mbar_init(bar(B_XFREE + i), 4); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar(B_A1FULL + i), 4); mbar_init(bar(B_A1FREE + i), 1);
            mbar_init(bar(B_RFULL + i), 1); mbar_init(bar(B_HFULL + i), GELU_WARPS);
            mbar_init(bar(B_ACCFULL + i), 1); mbar_init(bar(B_ACCFREE + i), 4);
        }
        mbar_init(bar(B_LFREE), 1);
        for (int i = 0; i < MAX_WST; ++i) { mbar_init(bar(B_WFULL + i), 1); mbar_init(bar(B_WFREE + i), 1); }
        fence_barrier_init();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == W_XLOAD) {
        // ================================================================================ x tile loader (TMA)
        if (elect_one()) {
            const uint32_t tile_bytes = (uint32_t)(p.nboxf * BOX_BYTES + BM * p.rem * 4);
            for (int t = 0; t < my_tiles; ++t) {
                const int s = t % p.nx, use = t / p.nx;
                mbar_wait(bar(B_XFREE + s), (uint32_t)((use & 1) ^ 1));          // the store of the tile that used the slot has read it
                const int row0 = ((int)blockIdx.x + t * (int)gridDim.x) * BM;
                const uint32_t dst = sX + (uint32_t)s * p.xslot_bytes;
                mbar_expect_tx(bar(B_XFULL + s), tile_bytes);
                for (int f = 0; f < p.nboxf; ++f) tma_load_2d(dst + f * BOX_BYTES, &mapLf, 32 * f, row0, bar(B_XFULL + s));
                if (p.rem) tma_load_2d(dst + p.nboxf * BOX_BYTES, &mapLr, 32 * p.nboxf, row0, bar(B_XFULL + s));
            }
        }
        __syncwarp();
    } else if (warp == W_WLOAD) {
        // ================================================================================ weight loader
        if (elect_one() && my_tiles > 0) {
            const uint8_t* img = reinterpret_cast<const uint8_t*>(p.w_img);
            if (p.resident) {
                int st = 0;
                for (int j = 0; j < p.nch; ++j) {
                    uint32_t off = (uint32_t)j * p.chunk_bytes;
                    for (int kb = 0; kb < p.nkb1 + 2; ++kb, ++st) {
                        const uint32_t bytes = kb < p.nkb1 ? (uint32_t)ST1_BYTES : p.st2_bytes;
                        mbar_expect_tx(bar(B_WFULL + st), bytes);
                        bulk_g2s(sW + off, img + off, bytes, bar(B_WFULL + st));
                        off += bytes;
                    }
                }
            } else {
                uint32_t slot = 0, phase = 0;
                auto load = [&](uint32_t goff, uint32_t bytes) {
                    mbar_wait(bar(B_WFREE + slot), phase ^ 1);
                    mbar_expect_tx(bar(B_WFULL + slot), bytes);
                    bulk_g2s(sW + slot * p.slot_bytes, img + goff, bytes, bar(B_WFULL + slot));
                    if (++slot == (uint32_t)p.ns) { slot = 0; phase ^= 1; }
                };
                for_each_op(total, p.nch, p.na1,
                    [&](int c) {
                        const uint32_t base = (uint32_t)(c % p.nch) * p.chunk_bytes;
                        for (int kb = 0; kb < p.nkb1; ++kb) load(base + kb * ST1_BYTES, ST1_BYTES);
                    },
                    [&](int c) {
                        const uint32_t base = (uint32_t)(c % p.nch) * p.chunk_bytes + p.nkb1 * ST1_BYTES;
                        for (int kb = 0; kb < 2; ++kb) load(base + kb * p.st2_bytes, p.st2_bytes);
                    });
            }
        }
        __syncwarp();
    } else if (warp == W_MMA) {
        // ================================================================================ MMA issuer
        const uint32_t idesc1 = make_idesc(HC), idesc2 = make_idesc(p.N2);
        uint32_t wslot = 0, wphase = 0;
        // stage of the next weight block: resident -> fixed position `rst`, streamed -> ring slot
        auto stage_wait = [&](int rst, uint32_t& saddr, uint32_t& sbar) {
            if (p.resident) {
                const int j = rst / (p.nkb1 + 2), kb = rst % (p.nkb1 + 2);
                saddr = sW + (uint32_t)j * p.chunk_bytes + (kb < p.nkb1 ? (uint32_t)kb * ST1_BYTES : (uint32_t)p.nkb1 * ST1_BYTES + (uint32_t)(kb - p.nkb1) * p.st2_bytes);
                sbar = 0;
                mbar_wait(bar(B_WFULL + rst), 0u);
            } else {
                saddr = sW + wslot * p.slot_bytes;
                sbar = bar(B_WFREE + wslot);
                mbar_wait(bar(B_WFULL + wslot), wphase);
                if (++wslot == (uint32_t)p.ns) { wslot = 0; wphase ^= 1; }
            }
        };
        for_each_op(total, p.nch, p.na1,
            [&](int c) {                                                           // G1(c): R[c % 2] = A1 * W1[chunk]^T
                const int t = c / p.nch, j = c % p.nch, a = t % p.na1;
                if (j == 0) mbar_wait(bar(B_A1FULL + a), (uint32_t)((t / p.na1) & 1));
                const uint32_t a_hi = tmem + (uint32_t)(p.col_a1 + a * 2 * p.Kp16), a_lo = a_hi + (uint32_t)p.Kp16;
                const uint32_t d = tmem + (uint32_t)(p.col_r + (c & 1) * HC);
                for (int kb = 0; kb < p.nkb1; ++kb) {
                    uint32_t saddr, sbar;
                    stage_wait(j * (p.nkb1 + 2) + kb, saddr, sbar);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint64_t b_hi = make_desc(saddr), b_lo = make_desc(saddr + HC * 128);
                        const int ks_end = p.ksteps1 - 4 * kb < 4 ? p.ksteps1 - 4 * kb : 4;
                        for (int ks = 0; ks < ks_end; ++ks) {
                            const uint32_t ac = (uint32_t)((4 * kb + ks) * 8);
                            const uint64_t adv = (uint64_t)(ks * 2);
                            umma_ts_tf32(d, a_lo + ac, b_hi + adv, idesc1, (kb | ks) ? 1u : 0u);
                            umma_ts_tf32(d, a_hi + ac, b_lo + adv, idesc1, 1u);
                            umma_ts_tf32(d, a_hi + ac, b_hi + adv, idesc1, 1u);
                        }
                        if (sbar) umma_commit(sbar);
                        if (kb + 1 == p.nkb1) {
                            umma_commit(bar(B_RFULL + (c & 1)));
                            if (j + 1 == p.nch) umma_commit(bar(B_A1FREE + a));
                        }
                    }
                    __syncwarp();
                }
            },
            [&](int c) {                                                           // G2(c): ACC2 += GELU chunk * W2[:, chunk]^T
                const int t = c / p.nch, j = c % p.nch, ab = t % p.nacc;
                mbar_wait(bar(B_HFULL + (c & 1)), (uint32_t)((c >> 1) & 1));
                if (j == 0) mbar_wait(bar(B_ACCFREE + ab), (uint32_t)(((t / p.nacc) & 1) ^ 1));
                const uint32_t a_hi = tmem + (uint32_t)(p.col_r + (c & 1) * HC);
                const uint32_t a_lo = tmem + (uint32_t)(p.col_l + (p.nl == 2 ? (c & 1) : 0) * HC);
                const uint32_t d = tmem + (uint32_t)(p.col_acc + ab * p.N2);
                for (int kb = 0; kb < 2; ++kb) {
                    uint32_t saddr, sbar;
                    stage_wait(j * (p.nkb1 + 2) + p.nkb1 + kb, saddr, sbar);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint64_t b_hi = make_desc(saddr), b_lo = make_desc(saddr + (uint32_t)p.N2 * 128);
                        for (int ks = 0; ks < 4; ++ks) {
                            const uint32_t ac = (uint32_t)((4 * kb + ks) * 8);
                            const uint64_t adv = (uint64_t)(ks * 2);
                            umma_ts_tf32(d, a_lo + ac, b_hi + adv, idesc2, (j | kb | ks) ? 1u : 0u);
                            umma_ts_tf32(d, a_hi + ac, b_lo + adv, idesc2, 1u);
                            umma_ts_tf32(d, a_hi + ac, b_hi + adv, idesc2, 1u);
                        }
                        if (sbar) umma_commit(sbar);
                        if (kb == 1) {
                            if (p.nl == 1) umma_commit(bar(B_LFREE));
                            if (j + 1 == p.nch) umma_commit(bar(B_ACCFULL + ab));
                        }
                    }
                    __syncwarp();
                }
            });
    } else if (warp >= LN_BASE && warp < LN_BASE + 4) {
        // ================================================================================ LayerNorm -> A1 (TMEM)
        const int q = warp & 3, r = q * 32 + lane;
        const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
        const float invC = 1.0f / (float)p.C;
        for (int t = 0; t < my_tiles; ++t) {
            const int s = t % p.nx, a = t % p.na1;
            mbar_wait(bar(B_XFULL + s), (uint32_t)((t / p.nx) & 1));
            const uint8_t* xs = smem + (size_t)s * p.xslot_bytes;
            float sum = 0.f;
            for (int k = 0; k < p.ld; k += 4) {
                const float4 v = mask4(*reinterpret_cast<const float4*>(xs + x_off(p, r, k)), k, p.C);
                sum += (v.x + v.y) + (v.z + v.w);
            }
            const float mean = sum * invC;
            float sq = 0.f;
            for (int k = 0; k < p.ld; k += 4) {
                const float4 v = *reinterpret_cast<const float4*>(xs + x_off(p, r, k));
                { const float d = v.x - mean; sq = fmaf(d, d, sq); }
                if (k + 1 < p.C) { const float d = v.y - mean; sq = fmaf(d, d, sq); }
                if (k + 2 < p.C) { const float d = v.z - mean; sq = fmaf(d, d, sq); }
                if (k + 3 < p.C) { const float d = v.w - mean; sq = fmaf(d, d, sq); }
            }
            const float rstd = 1.0f / sqrtf(sq * invC + p.eps);
            mbar_wait(bar(B_A1FREE + a), (uint32_t)(((t / p.na1) & 1) ^ 1));     // the G1s of the tile that used this buffer are done
            tc_fence_after();
            const uint32_t t_hi = lane_base + (uint32_t)(p.col_a1 + a * 2 * p.Kp16), t_lo = t_hi + (uint32_t)p.Kp16;
            for (int g = 0; g < p.Kp16; g += 16) {
                float hi[16], lo[16];
#pragma unroll
                for (int c4 = 0; c4 < 4; ++c4) {
                    const int k = g + 4 * c4;
                    float4 v = zero4();
                    if (k < p.C) {
                        v = *reinterpret_cast<const float4*>(xs + x_off(p, r, k));
                        const float4 gm = __ldg(reinterpret_cast<const float4*>(p.gamma + k)), be = __ldg(reinterpret_cast<const float4*>(p.beta + k));
                        v.x = (v.x - mean) * rstd * gm.x + be.x;
                        v.y = (v.y - mean) * rstd * gm.y + be.y;
                        v.z = (v.z - mean) * rstd * gm.z + be.z;
                        v.w = (v.w - mean) * rstd * gm.w + be.w;
                        v = mask4(v, k, p.C);
                    }
                    float4 h4, l4;
                    split4(v, h4, l4);
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
        }
    } else if (warp >= GELU_BASE && warp < GELU_BASE + GELU_WARPS) {
        // ================================================================================ bias + GELU + split, in TMEM
        const int q = warp & 3, half = (warp - GELU_BASE) >> 2;
        const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
        for (int c = 0; c < total; ++c) {
            const int j = c % p.nch, b = c & 1;
            mbar_wait(bar(B_RFULL + b), (uint32_t)((c >> 1) & 1));
            if (p.nl == 1 && c > 0) mbar_wait(bar(B_LFREE), (uint32_t)((c - 1) & 1));   // G2(c - 1) has read the single lo buffer
            tc_fence_after();
            const uint32_t t_r = lane_base + (uint32_t)(p.col_r + b * HC + half * 32);
            const uint32_t t_l = lane_base + (uint32_t)(p.col_l + (p.nl == 2 ? b : 0) * HC + half * 32);
            const float* bias = p.b1 + j * HC + half * 32;
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                float v[16], lo[16];
                tmem_ld16(t_r + (uint32_t)(16 * g), v);
                add_bias<16>(v, bias + 16 * g);
#pragma unroll
                for (int i = 0; i < 16; i += 2) gelu_erf2(v[i], v[i + 1]);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float h = tf32_rna(v[i]);
                    lo[i] = tf32_rna(v[i] - h);
                    v[i] = h;
                }
                tmem_st16(t_r + (uint32_t)(16 * g), v);
                tmem_st16(t_l + (uint32_t)(16 * g), lo);
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(B_HFULL + b));
        }
    } else if (warp >= OUT_BASE) {
        // ================================================================================ bias + residual + store
        const int q = warp & 3, r = q * 32 + lane;
        const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
        const float invC = 1.0f / (float)p.C;
        const bool leader = elect_one();                   // bulk async-groups are per thread: one lane issues every store and waits
        for (int t = 0; t < my_tiles; ++t) {
            const int s = t % p.nx, ab = t % p.nacc;
            const int row0 = ((int)blockIdx.x + t * (int)gridDim.x) * BM;
            mbar_wait(bar(B_ACCFULL + ab), (uint32_t)((t / p.nacc) & 1));
            tc_fence_after();
            uint8_t* xs = smem + (size_t)s * p.xslot_bytes;
            const uint32_t t_acc = lane_base + (uint32_t)(p.col_acc + ab * p.N2);
            float sum = 0.f;
            for (int g = 0; g < p.N2; g += 16) {
                float v[16];
                tmem_ld16(t_acc + (uint32_t)g, v);
                add_bias<16>(v, p.b2 + g);
#pragma unroll
                for (int c4 = 0; c4 < 4; ++c4) {
                    const int k = g + 4 * c4;
                    if (k < p.ld) {
                        float4* px = reinterpret_cast<float4*>(xs + x_off(p, r, k));
                        const float4 x4 = *px;
                        float4 o;
                        o.x = x4.x + v[4 * c4]; o.y = x4.y + v[4 * c4 + 1]; o.z = x4.z + v[4 * c4 + 2]; o.w = x4.w + v[4 * c4 + 3];
                        *px = o;
                        const float4 m4 = mask4(o, k, p.C);
                        sum += (m4.x + m4.y) + (m4.z + m4.w);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(B_ACCFREE + ab));
            if (p.stats_out) {
                const float mean = sum * invC;
                float sq = 0.f;
                for (int k = 0; k < p.ld; k += 4) {
                    const float4 v = *reinterpret_cast<const float4*>(xs + x_off(p, r, k));
                    { const float d = v.x - mean; sq = fmaf(d, d, sq); }
                    if (k + 1 < p.C) { const float d = v.y - mean; sq = fmaf(d, d, sq); }
                    if (k + 2 < p.C) { const float d = v.z - mean; sq = fmaf(d, d, sq); }
                    if (k + 3 < p.C) { const float d = v.w - mean; sq = fmaf(d, d, sq); }
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
                for (int f = 0; f < p.nboxf; ++f) tma_store_2d(&mapSf, 32 * f, row0 + q * 32, src + f * BOX_BYTES + q * 32 * 128);
                if (p.rem) tma_store_2d(&mapSr, 32 * p.nboxf, row0 + q * 32, src + p.nboxf * BOX_BYTES + q * 32 * p.rem * 4);
                tma_store_commit();
                tma_store_wait_read();                    // shared memory has been read: the slot may be refilled
                mbar_arrive(bar(B_XFREE + s));
            }
            __syncwarp();
        }
        if (leader) tma_store_wait_all();                 // global writes complete before the CTA retires
        __syncwarp();
    }

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

cudaError_t launch(cudaStream_t st, const Weights& w, float* x, long long M, float eps, const StatsOut& so) {
    const Plan& pl = w.plan;
    if (!pl.ok || !w.img) return cudaErrorInvalidValue;
    if (M <= 0) return cudaSuccess;
    if (M >= (1LL << 31) - BM) return cudaErrorInvalidValue;
    static std::atomic<bool> configured[tc::kMaxDevices];
    const int dev = tc::current_device();
    if (!configured[dev].load(std::memory_order_acquire)) {
        const cudaError_t e = cudaFuncSetAttribute(mlp_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_MAX);
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
    p.nx = pl.nx; p.na1 = pl.na1; p.nl = pl.nl; p.nacc = pl.nacc; p.resident = pl.resident; p.ns = pl.ns;
    p.nboxf = pl.nboxf; p.rem = pl.rem;
    p.st2_bytes = pl.st2_bytes; p.slot_bytes = pl.slot_bytes; p.chunk_bytes = pl.chunk_bytes; p.xslot_bytes = pl.xslot_bytes;
    p.col_a1 = pl.col_a1; p.col_r = pl.col_r; p.col_l = pl.col_l; p.col_acc = pl.col_acc;
    p.w_img = w.img; p.b1 = w.b1; p.b2 = w.b2; p.gamma = w.gamma; p.beta = w.beta;
    p.eps = eps;
    p.M = M;
    p.ntiles = (int)((M + BM - 1) / BM);
    p.stats_out = so.out;
    p.stat_geom = so.geom;
    p.H = so.H; p.W = so.W;
    p.dHW = FastDiv::make((unsigned)(so.H > 0 ? so.H * so.W : 1));
    p.dW_ = FastDiv::make((unsigned)(so.W > 0 ? so.W : 1));
    p.ng = so.ng;
    int grid = tc::sm_count();
    if (grid > p.ntiles) grid = p.ntiles;
    mlp_fused_kernel<<<grid, THREADS, pl.smem_bytes, st>>>(mLf, mLr, mSf, mSr, p);
    return cudaGetLastError();
}

}  // namespace mf
}  // namespace escb
