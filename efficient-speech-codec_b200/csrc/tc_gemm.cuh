// tcgen05 GEMM engine: fp32-grade products on the 5th-gen tensor cores by error-compensated TF32 splitting.
//
//   Y = epilogue( LN?(gather(A)) [M,K] * W^T ),  W [N,K] (nn.Linear layout = K-major B operand)
//
// Why 3xTF32: the RVQ code indices must match the fp32 reference bit for bit, and plain TF32 / BF16 inputs flip
// 3..742 of 3600 codes per stream (hard part 1).  Every operand x is split into
// hi = tf32(x), lo = tf32(x - hi); D += A_lo*B_hi + A_hi*B_lo + A_hi*B_hi in the fp32 TMEM accumulator
// (the dropped lo*lo term is 2^-22 relative).  Weights are split once at escb_finalize(); activations are split
// by the A-producer threads after the fused gather / LayerNorm.
//
// Structure of one CTA (one 128 x BN output tile, 256 threads, 2 CTAs per SM so one CTA's epilogue overlaps the
// other's main loop):
//   * B operand: pre-swizzled smem images of the weights in HBM, fetched per 32-wide K block with ONE
//     cp.async.bulk (TMA bulk engine, SASS UBLKCP) into a 2-stage ring, completion on an mbarrier;
//   * A operand: each thread gathers 4 float4 of the logical rows (window partition + cyclic shift, frequency-row
//     pairing, im2col ... same loaders as the SIMT engine), applies LayerNorm, splits hi/lo and writes the
//     128-byte-swizzled K-major layout tcgen05 expects; the loads of block k+1 are in flight while block k runs;
//   * one elected thread issues tcgen05.mma.kind::tf32 (M=128, N=BN, K=8) x 3 per k-step and tcgen05.commit;
//   * epilogue: 8 warps read the accumulator with tcgen05.ld 32x32b (one row per thread, 16 columns a time)
//     and apply bias / GELU / residual / scatter through the same epilogue functors as the SIMT engine.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "gemm.cuh"

namespace escb {
namespace tc {

constexpr int BM = 128;               // UMMA M
constexpr int KB = 32;                // tf32 per 128-byte swizzle row = one K block
constexpr int THREADS = 256;
constexpr int MAX_BN = 144;
constexpr int A_BYTES = 2 * BM * 128; // hi + lo images of one A block

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
// TMA bulk copy global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
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

// K-major, 128-byte swizzle, 8-row groups 1024 bytes apart; version 1 (sm_100)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// c_format F32 (1 << 4), a/b format TF32 (2 << 7, 2 << 10), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
__device__ __forceinline__ uint32_t make_idesc(int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

__device__ __forceinline__ float tf32_rna(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ void split4(const float4 v, float4& hi, float4& lo) {
    hi.x = tf32_rna(v.x); lo.x = tf32_rna(v.x - hi.x);
    hi.y = tf32_rna(v.y); lo.y = tf32_rna(v.y - hi.y);
    hi.z = tf32_rna(v.z); lo.z = tf32_rna(v.z - hi.z);
    hi.w = tf32_rna(v.w); lo.w = tf32_rna(v.w - hi.w);
}

template <bool LN, class AL, class EP>
__global__ void __launch_bounds__(THREADS, 2)
tc_gemm_kernel(const AL al, const LnParams ln, const TcWeight w, const long long M, const EP ep, const int tmem_cols) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ typename AL::Row rows[BM];
    __shared__ __align__(8) uint64_t bars[3];          // [0],[1]: B stage landed; [2]: MMAs of a K block retired
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nt = blockIdx.x % w.ntn;
    const long long m0 = (long long)(blockIdx.x / w.ntn) * BM;
    const int K = w.K, BN = w.BN, nkb = w.nkb;

    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t sA = smem_u32(smem);                   // A hi image, then A lo image (16 KB each)
    const uint32_t sB = sA + A_BYTES;                     // 2 stages x (hi image, lo image), BN x 128 B each
    const uint32_t b_img = (uint32_t)BN * 128u, b_stage = 2u * b_img;
    const uint32_t bar0 = smem_u32(&bars[0]);

    if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), (uint32_t)tmem_cols);
    if (tid == 32) {
        mbar_init(bar0, 1); mbar_init(bar0 + 8, 1); mbar_init(bar0 + 16, 1);
        fence_barrier_init();
    }
    for (int r = tid; r < BM; r += THREADS) al.init(m0 + r, M, rows[r]);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;

    const uint8_t* wimg = (const uint8_t*)w.img + (size_t)nt * nkb * b_stage;
    if (tid == 0) {
        for (int s = 0; s < 2 && s < nkb; ++s) {
            mbar_expect_tx(bar0 + 8 * s, b_stage);
            bulk_g2s(sB + s * b_stage, wimg + (size_t)s * b_stage, b_stage, bar0 + 8 * s);
        }
    }

    // A producer mapping: 16-byte chunk c of the 128-byte row, rows r0 + 32 i (8 consecutive lanes share a row)
    const int c = tid & 7, r0 = tid >> 3;
    typename AL::Row myrow[4];
    bool vld[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { myrow[i] = rows[r0 + 32 * i]; vld[i] = al.valid(myrow[i]); }

    float mean[4] = {0.f, 0.f, 0.f, 0.f}, rstd[4] = {0.f, 0.f, 0.f, 0.f};
    if (LN) {   // per-row mean / rstd, two passes; every load of a pass is independent (second pass hits L1/L2)
        float s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
        for (int kb = 0; kb < nkb; ++kb) {
            const int k = kb * KB + c * 4;
            if (k < K) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (vld[i]) { const float4 v = al.load4(myrow[i], k, K); s[i] += (v.x + v.y) + (v.z + v.w); }
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            s[i] += __shfl_xor_sync(0xffffffffu, s[i], 1);
            s[i] += __shfl_xor_sync(0xffffffffu, s[i], 2);
            s[i] += __shfl_xor_sync(0xffffffffu, s[i], 4);
            mean[i] = s[i] / (float)K;
            s[i] = 0.f;
        }
#pragma unroll 4
        for (int kb = 0; kb < nkb; ++kb) {
            const int k = kb * KB + c * 4;
            if (k < K) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (vld[i]) {
                        const float4 v = al.load4(myrow[i], k, K);
                        float d = v.x - mean[i]; s[i] = fmaf(d, d, s[i]);
                        if (k + 1 < K) { d = v.y - mean[i]; s[i] = fmaf(d, d, s[i]); }
                        if (k + 2 < K) { d = v.z - mean[i]; s[i] = fmaf(d, d, s[i]); }
                        if (k + 3 < K) { d = v.w - mean[i]; s[i] = fmaf(d, d, s[i]); }
                    }
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            s[i] += __shfl_xor_sync(0xffffffffu, s[i], 1);
            s[i] += __shfl_xor_sync(0xffffffffu, s[i], 2);
            s[i] += __shfl_xor_sync(0xffffffffu, s[i], 4);
            rstd[i] = 1.0f / sqrtf(s[i] / (float)K + ln.eps);
        }
    }

    float4 a[4];
    auto load_a = [&](int kb) {
        const int k = kb * KB + c * 4;
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = (vld[i] && k < K) ? al.load4(myrow[i], k, K) : zero4();
    };
    load_a(0);

    const uint32_t idesc = make_idesc(BN);
    for (int kb = 0; kb < nkb; ++kb) {
        const int k = kb * KB + c * 4;
        float4 g = zero4(), be = zero4();
        if (LN && k < K) { g = ldg4(ln.gamma + k); be = ldg4(ln.beta + k); }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = r0 + 32 * i;
            float4 v = a[i];
            if (LN && k < K && vld[i]) {
                v.x = (v.x - mean[i]) * rstd[i] * g.x + be.x;
                v.y = (v.y - mean[i]) * rstd[i] * g.y + be.y;
                v.z = (v.z - mean[i]) * rstd[i] * g.z + be.z;
                v.w = (v.w - mean[i]) * rstd[i] * g.w + be.w;
                v = mask4(v, k, K);
            }
            float4 hi, lo;
            split4(v, hi, lo);
            const uint32_t off = (uint32_t)r * 128u + ((uint32_t)(c ^ (r & 7)) << 4);
            *reinterpret_cast<float4*>(smem + off) = hi;
            *reinterpret_cast<float4*>(smem + BM * 128 + off) = lo;
        }
        fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
            const int s = kb & 1;
            mbar_wait(bar0 + 8 * s, (uint32_t)((kb >> 1) & 1));
            tc_fence_after();
            int rem = K - kb * KB;
            if (rem > KB) rem = KB;
            const int ksteps = (rem + 7) >> 3;
            const uint64_t a_hi = make_desc(sA), a_lo = make_desc(sA + BM * 128);
            const uint64_t b_hi = make_desc(sB + s * b_stage), b_lo = make_desc(sB + s * b_stage + b_img);
            for (int ks = 0; ks < ksteps; ++ks) {
                const uint64_t adv = (uint64_t)(ks * 2);      // 32 bytes >> 4 inside the swizzle row
                umma_tf32(tmem, a_lo + adv, b_hi + adv, idesc, (kb | ks) ? 1u : 0u);
                umma_tf32(tmem, a_hi + adv, b_lo + adv, idesc, 1u);
                umma_tf32(tmem, a_hi + adv, b_hi + adv, idesc, 1u);
            }
            umma_commit(bar0 + 16);
        }
        if (kb + 1 < nkb) load_a(kb + 1);
        mbar_wait(bar0 + 16, (uint32_t)(kb & 1));
        if (tid == 0 && kb + 2 < nkb) {
            const int s = kb & 1;
            mbar_expect_tx(bar0 + 8 * s, b_stage);
            bulk_g2s(sB + s * b_stage, wimg + (size_t)(kb + 2) * b_stage, b_stage, bar0 + 8 * s);
        }
    }
    tc_fence_after();

    // epilogue: warp -> TMEM lane quadrant (warp & 3), column chunks of 16 interleaved over the two warp halves
    {
        const int q = warp & 3, half = warp >> 2;
        const int r = q * 32 + lane;
        const long long m = m0 + r;
        typename EP::Row er;
        bool ok = m < M;
        if (ok) ok = ep.row(m, er);
        const int n0 = nt * BN, N = w.N;
        for (int ch = half; ch < BN / 16; ch += 2) {
            float v[16];
            tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(ch * 16), v);
            if (ok) {
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                    const int n = n0 + ch * 16 + j;
                    if (n + 3 < N) {
                        ep.store4(er, n, make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
                    } else {
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            if (n + u < N) ep.store(er, n + u, v[j + u]);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, (uint32_t)tmem_cols);
}

template <bool LN, class AL, class EP>
inline cudaError_t launch(cudaStream_t st, const AL& al, const LnParams& ln, const GemmWeight& gw, long long M, const EP& ep) {
    const TcWeight& w = gw.tc;
    if (!w.img || M <= 0) return M <= 0 ? cudaSuccess : cudaErrorInvalidValue;
    const size_t smem = 1024 + A_BYTES + 2 * 2 * (size_t)w.BN * 128;
    static bool configured = false;     // per instantiation
    if (!configured) {
        const cudaError_t e = cudaFuncSetAttribute(tc_gemm_kernel<LN, AL, EP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                   (int)(1024 + A_BYTES + 4 * MAX_BN * 128));
        if (e != cudaSuccess) return e;
        configured = true;
    }
    const long long ntm = (M + BM - 1) / BM;
    const int cols = w.BN <= 32 ? 32 : (w.BN <= 64 ? 64 : (w.BN <= 128 ? 128 : 256));
    tc_gemm_kernel<LN, AL, EP><<<(unsigned)(ntm * w.ntn), THREADS, smem, st>>>(al, ln, w, M, ep, cols);
    return cudaGetLastError();
}

}  // namespace tc
}  // namespace escb
