// fp32 GEMM engine with fused gather/LayerNorm prologues and scatter/activation epilogues.
//
//   Y = epilogue( LN?(gather(A)) [M,K] * Wt [K,N] )
//
// Every dense contraction of the ESC hot path (QKV / proj / MLP linears, PatchMerge / PatchSplit, the
// product-VQ projections, the 5x5 de-embedding conv as implicit GEMM, the DFT / inverse-DFT) is one
// instantiation of this kernel: the "A loader" describes how a logical row is gathered from HBM (window
// partition + cyclic shift, frequency-row pairing, VQ frame layout, im2col, STFT framing ...) and the
// epilogue describes where each output element lands (window reverse, pixel shuffle, overlap-add ...),
// so none of the reference's permute/contiguous copies (21 % of its CPU time) exist here.
//
// Arithmetic is plain fp32 FMA on the CUDA cores: bit-exact RVQ indices need fp32-grade products
// (SURVEY.md §7 hard part 1).  Tile: (16*TM) x (16*TN) x 16, 256 threads, register double-buffered.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace escb {

constexpr int kBK = 16;
constexpr int kGemmThreads = 256;

struct LnParams {
    const float* gamma;   // padded with zeros to a multiple of 4
    const float* beta;
    float eps;
    float2* stats;        // [M] (mean, rstd) scratch of the tcgen05 engine's ln_stats_kernel (unused by the SIMT engine)
    unsigned long long* trace;   // ESCB_TC_TRACE builds: 16 counters of this launch (null otherwise)
    // post-GEMM LayerNorm (tc_gemm.cuh LNP): the weight image holds gamma-scaled rows W' = gamma o W and the epilogue
    // computes rstd * (acc - mean * cs) + f * bw (+ bias), cs[n] = sum_k W'[n][k], bw[n] = sum_k beta[k] W[n][k],
    // f = 0 for zero-padded rows (which bypass the norm in the reference)
    const float* cs;
    const float* bw;
};

// streaming 16-byte load: activations are read once per GEMM, so they bypass L1 allocation (ESCB_LDG_PLAIN: plain __ldg)
__device__ __forceinline__ float4 ldg4(const float* p) {
#ifdef ESCB_LDG_PLAIN
    return __ldg(reinterpret_cast<const float4*>(p));
#else
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::128B.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
#endif
}
__device__ __forceinline__ float4 zero4() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float4 mask4(float4 v, int k, int K) {
    if (k + 1 >= K) v.y = 0.f;
    if (k + 2 >= K) v.z = 0.f;
    if (k + 3 >= K) v.w = 0.f;
    return v;
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// L2 prefetch of elements [0, n) of a row, 128 bytes per probe, probes interleaved over four threads
__device__ __forceinline__ void prefetch_row(const float* p, int n, int part) {   // part in 0..3
    for (int k = part * 32; k < n; k += 128) prefetch_l2(p + k);
}
__device__ __forceinline__ float gelu_erf(float x) { return x * 0.5f * (1.0f + erff(x * 0.70710678118654752440f)); }

// packed fp32 pairs: Blackwell's FFMA2 / FMUL2 do two independent fp32 operations (scalar rounding) per issue slot
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

// Exact (erf) GELU of two values, same accuracy as gelu_erf (max abs error 3.8e-7 vs 4.5e-7 over [-8, 8], 1.5e-7
// for |x| < 2) at less than half its instruction count: with a = min(|x|, 6),
//   Phi(-a) = erfc(a / sqrt 2) / 2 = 2^Q(a),   Q(a) = -1 + a R(a),   R a degree-9 polynomial (Chebyshev fit of
//   (log2 Phi(-a) + 1) / a on [0, 6]),
// so gelu(x) = x * (x < 0 ? e : 1 - e) with e = 2^Q.  No range split (erff selects between two coefficient sets per
// element), the Horner chain runs on both lanes of FFMA2, and what matters downstream - the absolute error of
// x * Phi(x) - is that of the final rounding (Q is exact to ~1e-7 where e is not negligible).
__device__ __forceinline__ void gelu_erf2(float& x0, float& x1) {
    const float a0 = fminf(fabsf(x0), 6.0f), a1 = fminf(fabsf(x1), 6.0f);
    const unsigned long long a = pack2(a0, a1);
#define ESCB_C2(v) pack2(v, v)
    unsigned long long r = ESCB_C2(-1.945421602e-08f);
    r = ffma2(r, a, ESCB_C2(6.227343192e-07f));
    r = ffma2(r, a, ESCB_C2(-8.505132428e-06f));
    r = ffma2(r, a, ESCB_C2(6.270733866e-05f));
    r = ffma2(r, a, ESCB_C2(-2.353102609e-04f));
    r = ffma2(r, a, ESCB_C2(-8.306776726e-05f));
    r = ffma2(r, a, ESCB_C2(7.032935973e-03f));
    r = ffma2(r, a, ESCB_C2(-5.248502642e-02f));
    r = ffma2(r, a, ESCB_C2(-4.592124820e-01f));
    r = ffma2(r, a, ESCB_C2(-1.151104450e+00f));
    r = ffma2(r, a, ESCB_C2(-1.0f));
#undef ESCB_C2
    float q0, q1, e0, e1;
    unpack2(r, q0, q1);
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(q0));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(q1));
    x0 *= x0 < 0.f ? e0 : 1.0f - e0;
    x1 *= x1 < 0.f ? e1 : 1.0f - e1;
}

// ------------------------------------------------------------------------------------------------ kernel
template <int TM, int TN, bool LN, class AL, class EP>
__device__ __forceinline__ void gemm_tile(const AL& al, const LnParams& ln, const float* __restrict__ Wt, const int ldw,
                                          const long long M, const int N, const int K, const int Kpad, const int ntn,
                                          const EP& ep) {
    constexpr int BM = 16 * TM, BN = 16 * TN;
    constexpr int AP = BM + 4;
    constexpr int PA = BM / 64;                         // float4 A loads per thread per k-tile
    constexpr int NB4 = kBK * BN / 4;
    constexpr int PB = (NB4 + kGemmThreads - 1) / kGemmThreads;

    __shared__ __align__(16) float As[2][kBK][AP];
    __shared__ __align__(16) float Bs[2][kBK][BN];
    __shared__ typename AL::Row rows[BM];
    __shared__ float2 stats[LN ? BM : 1];

    const int tid = threadIdx.x;
    const long long m0 = (long long)(blockIdx.x / ntn) * BM;
    const int n0 = (blockIdx.x % ntn) * BN;

    for (int r = tid; r < BM; r += kGemmThreads) al.init(m0 + r, M, rows[r]);
    __syncthreads();

    if (LN) {   // per-row mean / rstd, two-pass, one warp per row
        const int warp = tid >> 5, lane = tid & 31;
        for (int r = warp; r < BM; r += kGemmThreads / 32) {
            const typename AL::Row row = rows[r];
            float mean = 0.f, rstd = 0.f;
            if (al.valid(row)) {
                float s = 0.f;
                for (int k = lane; k < K; k += 32) s += al.load1(row, k);
#pragma unroll
                for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                mean = s / (float)K;
                float q = 0.f;
                for (int k = lane; k < K; k += 32) { const float d = al.load1(row, k) - mean; q = fmaf(d, d, q); }
#pragma unroll
                for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
                rstd = 1.0f / sqrtf(q / (float)K + ln.eps);
            }
            if (lane == 0) stats[r] = make_float2(mean, rstd);
        }
        __syncthreads();
    }

    const int tx = tid & 15, ty = tid >> 4;
    const int a_kq = (tid & 3) * 4, a_r = tid >> 2;

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    float4 a_reg[PA], b_reg[PB];

    auto load_tiles = [&](int k0) {
#pragma unroll
        for (int p = 0; p < PA; ++p) {
            const int r = a_r + 64 * p;
            const typename AL::Row row = rows[r];
            const int k = k0 + a_kq;
            float4 v = zero4();
            if (al.valid(row) && k < K) {
                v = al.load4(row, k, K);
                if (LN) {
                    const float2 st = stats[r];
                    const float4 g = ldg4(ln.gamma + k), b = ldg4(ln.beta + k);
                    v.x = (v.x - st.x) * st.y * g.x + b.x;
                    v.y = (v.y - st.x) * st.y * g.y + b.y;
                    v.z = (v.z - st.x) * st.y * g.z + b.z;
                    v.w = (v.w - st.x) * st.y * g.w + b.w;
                    v = mask4(v, k, K);
                }
            }
            a_reg[p] = v;
        }
#pragma unroll
        for (int p = 0; p < PB; ++p) {
            const int idx = tid + kGemmThreads * p;
            float4 v = zero4();
            if (idx < NB4) {
                const int kk = idx / (BN / 4), nq = idx % (BN / 4);
                const int n = n0 + nq * 4;
                if (n < ldw) v = ldg4(Wt + (long long)(k0 + kk) * ldw + n);
            }
            b_reg[p] = v;
        }
    };
    auto store_tiles = [&](int buf) {
#pragma unroll
        for (int p = 0; p < PA; ++p) {
            const int r = a_r + 64 * p;
            As[buf][a_kq + 0][r] = a_reg[p].x;
            As[buf][a_kq + 1][r] = a_reg[p].y;
            As[buf][a_kq + 2][r] = a_reg[p].z;
            As[buf][a_kq + 3][r] = a_reg[p].w;
        }
#pragma unroll
        for (int p = 0; p < PB; ++p) {
            const int idx = tid + kGemmThreads * p;
            if (idx < NB4) {
                const int kk = idx / (BN / 4), nq = idx % (BN / 4);
                *reinterpret_cast<float4*>(&Bs[buf][kk][nq * 4]) = b_reg[p];
            }
        }
    };

    const int nk = Kpad / kBK;
    load_tiles(0);
    store_tiles(0);
    __syncthreads();
    for (int kt = 0; kt < nk; ++kt) {
        const int cur = kt & 1;
        if (kt + 1 < nk) load_tiles((kt + 1) * kBK);
#pragma unroll
        for (int kk = 0; kk < kBK; ++kk) {
            float a[TM], b[TN];
#pragma unroll
            for (int i = 0; i < TM; i += 4) {
                const float4 v = *reinterpret_cast<const float4*>(&As[cur][kk][ty * TM + i]);
                a[i] = v.x; a[i + 1] = v.y; a[i + 2] = v.z; a[i + 3] = v.w;
            }
#pragma unroll
            for (int j = 0; j < TN; ++j) b[j] = Bs[cur][kk][tx + 16 * j];
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kt + 1 < nk) {
            store_tiles(cur ^ 1);
            __syncthreads();
        }
    }

#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const long long m = m0 + ty * TM + i;
        if (m >= M) continue;
        typename EP::Row er;
        if (!ep.row(m, er)) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int n = n0 + tx + 16 * j;
            if (n < N) ep.store(er, n, acc[i][j]);
        }
    }
}

template <int TM, int TN, bool LN, class AL, class EP>
__global__ void __launch_bounds__(kGemmThreads)
gemm_kernel(const AL al, const LnParams ln, const float* __restrict__ Wt, const int ldw, const long long M,
            const int N, const int K, const int Kpad, const int ntn, const EP ep) {
    gemm_tile<TM, TN, LN, AL, EP>(al, ln, Wt, ldw, M, N, K, Kpad, ntn, ep);
}

// Three independent GEMMs of one shape in one grid (blockIdx.y picks the problem): the product-VQ down-projection
// is block diagonal over its 3 groups, so each group multiplies only its own third of the frame - a third of the
// flops of the stacked weight and three times the CTAs for a 43-row-tile problem.
template <class AL, class EP>
struct Gemm3 { AL al[3]; EP ep[3]; const float* wt[3]; };
template <int TM, int TN, class AL, class EP>
__global__ void __launch_bounds__(kGemmThreads)
gemm3_kernel(const __grid_constant__ Gemm3<AL, EP> g, const int ldw, const long long M, const int N, const int K,
             const int Kpad, const int ntn) {
    const LnParams ln{nullptr, nullptr, 0.f, nullptr, nullptr};
    const int y = blockIdx.y;
    gemm_tile<TM, TN, false, AL, EP>(g.al[y], ln, g.wt[y], ldw, M, N, K, Kpad, ntn, g.ep[y]);
}

// ------------------------------------------------------------------------------------------------ dispatch
// Operand images for the tcgen05 engine (tc_gemm.cuh): per (n tile, 32-wide K block) one contiguous chunk
// [hi image | lo image], each BN rows x 128 bytes in the 128-byte-swizzled K-major layout, hi = tf32(w),
// lo = tf32(w - hi).  img == nullptr: this weight has no tensor-core image.
struct TcWeight {
    const float* img;
    int N, K, BN, nsub, ntn, nkb;     // ntn output tiles of nsub sub-tiles of BN columns (tc::choose_tiling)
    int resident;       // the whole n-tile (nkb K blocks) stays in shared memory for the life of a CTA
    int wide;           // tiled for the 16-epilogue-warp role split (tc::Roles<4>)
    // Accumulator split (tc_gemm.cuh "accumulator split"): the hi*hi products of K block kb go to MAIN accumulator
    // kb % nmain; corr = 1 sends the lo*hi + hi*lo corrections to an accumulator of their own.  A sub-tile's region is
    // (nmain + corr) * BN TMEM columns [main 0 | ... | corr], summed by the epilogue.  nmain = 1, corr = 0: one accumulator.
    int nmain, corr;
};

struct GemmWeight {     // Wt [Kpad][ldw] row-major, zero padded; bias may be null
    const float* wt;
    const float* bias;
    int N, K, Kpad, ldw;
    TcWeight tc;
    TcWeight tc_alt;    // the same weight cut into 2-3x as many (narrower) output tiles; img null: none
    const float* cs;    // post-GEMM LayerNorm variants only (LnParams::cs / bw): column sums of the gamma-scaled weight
    const float* bw;    //                                                        and beta . W
};

inline int pick_tn(int N) {
    const int cand[6] = {3, 5, 6, 7, 8, 9};
    int best = 3;
    long long best_pad = -1;
    for (int c : cand) {
        const int bn = 16 * c;
        const long long pad = (long long)((N + bn - 1) / bn) * bn;
        if (best_pad < 0 || pad < best_pad || (pad == best_pad && c > best)) { best = c; best_pad = pad; }
    }
    return best;
}

template <int TM, int TN, bool LN, class AL, class EP>
inline cudaError_t launch_gemm_t(cudaStream_t st, const AL& al, const LnParams& ln, const GemmWeight& w,
                                 long long M, const EP& ep) {
    constexpr int BM = 16 * TM, BN = 16 * TN;
    const int ntn = (w.N + BN - 1) / BN;
    const long long ntm = (M + BM - 1) / BM;
    if (ntm <= 0) return cudaSuccess;
    gemm_kernel<TM, TN, LN, AL, EP><<<(unsigned)(ntm * ntn), kGemmThreads, 0, st>>>(
        al, ln, w.wt, w.ldw, M, w.N, w.K, w.Kpad, ntn, ep);
    return cudaGetLastError();
}

// Instantiates only the TN values listed in the TNS... pack (keeps compile time bounded); TN is chosen by
// pick_tn among them, TM by the number of row tiles (small problems get the 64-row tile to fill 148 SMs).
template <bool LN, class AL, class EP, int... TNS>
struct GemmLauncher {
    template <int TN>
    static cudaError_t go(cudaStream_t st, const AL& al, const LnParams& ln, const GemmWeight& w, long long M,
                          const EP& ep) {
        const int ntn = (w.N + 16 * TN - 1) / (16 * TN);
        if (((M + 127) / 128) * ntn >= 2 * 148) return launch_gemm_t<8, TN, LN, AL, EP>(st, al, ln, w, M, ep);
        return launch_gemm_t<4, TN, LN, AL, EP>(st, al, ln, w, M, ep);
    }
    static cudaError_t launch(cudaStream_t st, const AL& al, const LnParams& ln, const GemmWeight& w, long long M,
                              const EP& ep) {
        // best TN among the instantiated ones
        const int list[] = {TNS...};
        int best = list[0];
        long long best_pad = -1;
        // least column padding; on a tie the wider tile, unless the grid would leave SMs idle (skinny problems
        // such as the product-VQ down-projection: 43 row tiles), where the narrower tile doubles the CTAs
        const long long ntm64 = (M + 63) / 64;
        for (int c : list) {
            const int bn = 16 * c;
            const long long pad = (long long)((w.N + bn - 1) / bn) * bn;
            const int wider = c > best ? c : best;
            const bool small_grid = ntm64 * ((w.N + 16 * wider - 1) / (16 * wider)) < 148;
            if (best_pad < 0 || pad < best_pad || (pad == best_pad && (small_grid ? c < best : c > best))) { best = c; best_pad = pad; }
        }
        cudaError_t err = cudaErrorInvalidValue;
        const bool hit = ((best == TNS ? (err = go<TNS>(st, al, ln, w, M, ep), true) : false) || ...);
        (void)hit;
        return err;
    }
};

template <int TN, class AL, class EP>
inline cudaError_t launch_gemm3(cudaStream_t st, const Gemm3<AL, EP>& g, const GemmWeight& w, long long M) {
    const int ntn = (w.N + 16 * TN - 1) / (16 * TN);
    if (M <= 0) return cudaSuccess;
    if (3 * ((M + 127) / 128) * ntn >= 2 * 148)
        gemm3_kernel<8, TN, AL, EP><<<dim3((unsigned)(((M + 127) / 128) * ntn), 3), kGemmThreads, 0, st>>>(g, w.ldw, M, w.N, w.K, w.Kpad, ntn);
    else
        gemm3_kernel<4, TN, AL, EP><<<dim3((unsigned)(((M + 63) / 64) * ntn), 3), kGemmThreads, 0, st>>>(g, w.ldw, M, w.N, w.K, w.Kpad, ntn);
    return cudaGetLastError();
}

}  // namespace escb
