// Microbenchmark 5: error of a 3xTF32 GEMM row block (128 x N x K, fp32 operands split into hi + lo tf32 images, the
// engine's arithmetic) against float64, for different ways of laying the MMAs over TMEM accumulators.  tcgen05.mma
// truncates its fp32 accumulator (microbenchmark 4: ~0.6 ulp of bias toward zero per MMA), so the error grows with
// the number of MMAs chained into ONE accumulator and with that accumulator's magnitude:
//   S0  one accumulator, per k-step lo*hi, hi*lo, hi*hi                       (3 K/8 MMAs chained: the round-1 engine)
//   S1  hi*hi -> MAIN, lo*hi + hi*lo -> CORR (values ~2^-11 of MAIN)          (K/8 chained)
//   S2  as S1 with two MAIN accumulators taking alternate 32-wide K blocks    (K/16 chained)
//   S4  as S1 with four MAIN accumulators                                     (K/32 chained)
// The partial accumulators are added in fp32 (round to nearest) after the last MMA.  Next to them: a sequential fp32
// FMA chain on the host (what the SIMT engine and the reference's sgemm are made of).
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ uint32_t make_idesc(int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ float tf32_rna(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u); }
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;}" ::"r"(d), "l"(a),
                 "l"(b), "r"(idesc), "r"(acc) : "memory");
}
constexpr int N = 32;
// TMEM columns: S0 at 0; S1 main 32, corr 64; S2 main0 96, main1 128, corr 160; S4 mains 192..288+, corr 320
__global__ void __launch_bounds__(128, 1) k(const float* A, const float* B, float* D, int K) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    uint8_t* sAh = smem;                       // 128 x 128 B
    uint8_t* sAl = smem + 16384;
    uint8_t* sBh = smem + 32768;               // N x 128 B (padded to 1024-byte groups)
    uint8_t* sBl = smem + 32768 + 4096;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1u) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    const int nkb = (K + 31) / 32;
    uint32_t phase = 0;
    for (int kb = 0; kb < nkb; ++kb) {
        for (int i = tid; i < 128 * 32; i += 128) {
            const int r = i >> 5, kk = i & 31, kg = kb * 32 + kk;
            const float v = kg < K ? A[(size_t)r * K + kg] : 0.f;
            const float hi = tf32_rna(v), lo = tf32_rna(v - hi);
            const uint32_t off = r * 128 + ((((kk >> 2) ^ (r & 7)) << 4) | ((kk & 3) << 2));
            *reinterpret_cast<float*>(sAh + off) = hi;
            *reinterpret_cast<float*>(sAl + off) = lo;
        }
        for (int i = tid; i < N * 32; i += 128) {
            const int r = i >> 5, kk = i & 31, kg = kb * 32 + kk;
            const float v = kg < K ? B[(size_t)r * K + kg] : 0.f;
            const float hi = tf32_rna(v), lo = tf32_rna(v - hi);
            const uint32_t off = r * 128 + ((((kk >> 2) ^ (r & 7)) << 4) | ((kk & 3) << 2));
            *reinterpret_cast<float*>(sBh + off) = hi;
            *reinterpret_cast<float*>(sBl + off) = lo;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (tid == 0) {
            const uint32_t idesc = make_idesc(N);
            const uint64_t ah = make_desc(smem_u32(sAh)), al = make_desc(smem_u32(sAl)), bh = make_desc(smem_u32(sBh)),
                           bl = make_desc(smem_u32(sBl));
            for (int ks = 0; ks < 4; ++ks) {
                const uint64_t adv = (uint64_t)(ks * 2);
                const uint32_t first = (kb | ks) ? 1u : 0u;
                // S0
                mma(tmem + 0, al + adv, bh + adv, idesc, first);
                mma(tmem + 0, ah + adv, bl + adv, idesc, 1u);
                mma(tmem + 0, ah + adv, bh + adv, idesc, 1u);
                // S1
                mma(tmem + 64, al + adv, bh + adv, idesc, first);
                mma(tmem + 64, ah + adv, bl + adv, idesc, 1u);
                mma(tmem + 32, ah + adv, bh + adv, idesc, first);
                // S2
                mma(tmem + 160, al + adv, bh + adv, idesc, first);
                mma(tmem + 160, ah + adv, bl + adv, idesc, 1u);
                mma(tmem + 96 + 32 * (kb & 1), ah + adv, bh + adv, idesc, (kb >= 2 || ks) ? 1u : 0u);
                // S4
                mma(tmem + 320, al + adv, bh + adv, idesc, first);
                mma(tmem + 320, ah + adv, bl + adv, idesc, 1u);
                mma(tmem + 192 + 32 * (kb & 3), ah + adv, bh + adv, idesc, (kb >= 4 || ks) ? 1u : 0u);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        }
        uint32_t done = 0;
        while (!done) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p;}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(phase) : "memory");
        phase ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        __syncthreads();
    }
    // read all 352 columns of my lane
    for (int c0 = 0; c0 < 352; c0 += 16) {
        uint32_t r[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                       "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 16; ++j) D[(size_t)tid * 352 + c0 + j] = __uint_as_float(r[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}
static double gauss() {
    double u = (rand() + 1.0) / (RAND_MAX + 2.0), v = (rand() + 1.0) / (RAND_MAX + 2.0);
    return sqrt(-2.0 * log(u)) * cos(6.283185307179586 * v);
}
int main() {
    srand(7);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    printf("error of a 128 x %d x K product against float64: max|err| / max|ref| (rms err / rms ref)\n", N);
    for (int mode = 0; mode < 2; ++mode) {
        printf(mode == 0 ? "A ~ N(0,1) (LayerNorm output), B ~ N(0,1)/sqrt(K)\n" : "A = max(N(0,1), 0) (activation-like, one sign), B = |N(0,1)|/sqrt(K) (worst case: every product positive)\n");
        for (int K : {48, 96, 192, 384, 1536}) {
            float *A, *B, *D;
            cudaMallocManaged(&A, 128 * K * 4); cudaMallocManaged(&B, N * K * 4); cudaMallocManaged(&D, 128 * 352 * 4);
            for (int i = 0; i < 128 * K; ++i) { const double g = gauss(); A[i] = (float)(mode ? fmax(g, 0.0) : g); }
            for (int i = 0; i < N * K; ++i) { const double g = gauss() / sqrt((double)K); B[i] = (float)(mode ? fabs(g) : g); }
            k<<<1, 128, 64 * 1024>>>(A, B, D, K);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("K %d: %s\n", K, cudaGetErrorString(e)); return 1; }
            double mx[5] = {0, 0, 0, 0, 0}, sq[5] = {0, 0, 0, 0, 0}, rmax = 0, rsq = 0;
            for (int r = 0; r < 128; ++r)
                for (int n = 0; n < N; ++n) {
                    double ref = 0.0;
                    float chain = 0.f;
                    for (int kk = 0; kk < K; ++kk) {
                        ref += (double)A[(size_t)r * K + kk] * (double)B[(size_t)n * K + kk];
                        chain = fmaf(A[(size_t)r * K + kk], B[(size_t)n * K + kk], chain);
                    }
                    const float* d = D + (size_t)r * 352;
                    const int nkb = (K + 31) / 32;                   // accumulators no K block reached hold garbage
                    const float m2 = nkb > 1 ? d[96 + n] + d[128 + n] : d[96 + n];
                    float m4 = d[192 + n];
                    for (int j = 1; j < 4 && j < nkb; ++j) m4 += d[192 + 32 * j + n];
                    const float v[5] = {d[n], d[32 + n] + d[64 + n], m2 + d[160 + n], m4 + d[320 + n], chain};
                    for (int s = 0; s < 5; ++s) { const double er = fabs((double)v[s] - ref); mx[s] = fmax(mx[s], er); sq[s] += er * er; }
                    rmax = fmax(rmax, fabs(ref)); rsq += ref * ref;
                }
            printf("  K %4d   S0 %.2e (%.2e)   S1 %.2e (%.2e)   S2 %.2e (%.2e)   S4 %.2e (%.2e)   fp32 FMA chain %.2e (%.2e)\n", K,
                   mx[0] / rmax, sqrt(sq[0] / rsq), mx[1] / rmax, sqrt(sq[1] / rsq), mx[2] / rmax, sqrt(sq[2] / rsq), mx[3] / rmax,
                   sqrt(sq[3] / rsq), mx[4] / rmax, sqrt(sq[4] / rsq));
            cudaFree(A); cudaFree(B); cudaFree(D);
        }
    }
    return 0;
}
