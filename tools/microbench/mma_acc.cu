// Microbenchmark 4: how does tcgen05.mma kind::tf32 round its fp32 accumulator?  D (+)= A * B^T with tf32 operands whose
// products need 22 mantissa bits (A[m][k] = a_k, B[n][k] = v_n * b_k), issued `iters` times into the same accumulator: the
// exact result is iters * sum_k a_k b_k v_n.  Prints D's signed error in ulps of the result against double, next to what
// a chain of fp32 additions of the (exact) 8-term dot product gives on the host with ONE rounding per MMA, to nearest and
// toward zero.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <cfenv>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ uint32_t make_idesc(int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__global__ void __launch_bounds__(128, 1) k(const float* Av, const float* Bv, float* D, int N, int iters) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float* sA = reinterpret_cast<float*>(smem);                 // A image: 128 rows x 32 floats, all ones (swizzle irrelevant)
    float* sB = reinterpret_cast<float*>(smem + 16384);         // B image: row n = v_n everywhere
    // 128-byte-swizzled K-major images; only k = 0..7 (the first 32 bytes of every row) are used
    for (int i = tid; i < 128 * 32; i += 128) { const int r = i >> 5, kk = i & 31; sA[r * 32 + ((((kk >> 2) ^ (r & 7)) << 2) | (kk & 3))] = Av[kk & 7]; }
    for (int i = tid; i < N * 32; i += 128) { const int r = i >> 5, kk = i & 31; sB[r * 32 + ((((kk >> 2) ^ (r & 7)) << 2) | (kk & 3))] = Bv[r * 8 + (kk & 7)]; }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1u) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    if (tid == 0) {
        const uint32_t idesc = make_idesc(N);
        const uint64_t a = make_desc(smem_u32(sA)), b = make_desc(smem_u32(sB));
        for (int it = 0; it < iters; ++it)
            asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;}" ::"r"(tmem),
                         "l"(a), "l"(b), "r"(idesc), "r"(it ? 1u : 0u) : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t done = 0;
        while (!done) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p;}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    }
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t r[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                       "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (warp == 0 && lane == 0) for (int j = 0; j < 16; ++j) D[c0 + j] = __uint_as_float(r[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}
static float tf32(float x) { uint32_t u; memcpy(&u, &x, 4); u = (u + 0x1000u) & 0xffffe000u; memcpy(&x, &u, 4); return x; }
int main() {
    const int N = 16;
    float *Av, *Bv, *D;
    cudaMallocManaged(&Av, 8 * 4); cudaMallocManaged(&Bv, N * 8 * 4); cudaMallocManaged(&D, N * 4);
    const float vals[N] = {1.0f / 3, 0.7f, 1.9f, 0.0123f, -1.0f / 3, -0.7f, 2.0f / 7, 5.0f / 9, 0.11f, 0.93f, 1.37f, -1.37f, 3.3f, 0.57f, 0.81f, 0.29f};
    for (int kk = 0; kk < 8; ++kk) Av[kk] = tf32(1.0f + 0.137f * kk);
    for (int n = 0; n < N; ++n)
        for (int kk = 0; kk < 8; ++kk) Bv[n * 8 + kk] = tf32(vals[n] * (1.0f + 0.0731f * kk));
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    for (int iters : {1, 8, 64, 192, 576}) {
        k<<<1, 128, 64 * 1024>>>(Av, Bv, D, N, iters);
        cudaError_t e = cudaDeviceSynchronize();
        printf("iters %4d (%s): signed error in ulp of the result  [tensor core | host RN chain | host RZ chain]\n", iters, cudaGetErrorString(e));
        double stc = 0, srn = 0, srz = 0;
        for (int n = 0; n < N; ++n) {
            double dot = 0.0;
            for (int kk = 0; kk < 8; ++kk) dot += (double)Av[kk] * (double)Bv[n * 8 + kk];     // exact (22-bit products, 8 terms)
            const double exact = iters * dot;
            float rn = 0.f, rz = 0.f;
            for (int i = 0; i < iters; ++i) {
                fesetround(FE_TONEAREST); rn = (float)((double)rn + dot);
                fesetround(FE_TOWARDZERO); rz = (float)((double)rz + dot);
            }
            fesetround(FE_TONEAREST);
            const double ulp = ldexp(1.0, ilogb(fabs(exact)) - 23);
            if (iters == 576 || n < 4) printf("  v=% .6f  tc % 8.2f | rn % 8.2f | rz % 8.2f\n", vals[n], (D[n] - exact) / ulp, (rn - exact) / ulp, (rz - exact) / ulp);
            stc += fabs(D[n] - exact) / ulp; srn += fabs(rn - exact) / ulp; srz += fabs(rz - exact) / ulp;
        }
        printf("  mean |error| in ulp: tc %.2f  rn %.2f  rz %.2f\n", stc / N, srn / N, srz / N);
    }
    return 0;
}
