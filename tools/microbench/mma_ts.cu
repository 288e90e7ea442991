// Microbenchmark 3: tcgen05.mma kind::tf32 with the A operand in TENSOR MEMORY (written by tcgen05.st 32x32b),
// B in shared memory (K-major, 128-byte swizzle).  Verifies D = A * B^T against the host and times it.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ uint32_t make_idesc(int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
// A [128][32] row-major, B [N][32] row-major (both tf32-exact values), D [128][N]
__global__ void __launch_bounds__(128, 1) k(const float* A, const float* B, float* D, int N, int iters, long long* cyc) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // B image: row n at n*128 bytes, 16-byte chunk c stored at c ^ (n & 7)
    for (int i = tid; i < N * 8; i += 128) {
        const int n = i >> 3, c = i & 7;
        *reinterpret_cast<float4*>(smem + n * 128 + ((c ^ (n & 7)) << 4)) = *reinterpret_cast<const float4*>(B + n * 32 + c * 4);
    }
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
    // A -> TMEM columns [256, 288): lane = row (this warp's quadrant), column = k
    {
        const int row = warp * 32 + lane;
        uint32_t r[32];
        for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(A[row * 32 + j]);
        const uint32_t taddr = tmem + 256 + ((uint32_t)(warp * 32) << 16);
        asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
                     "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
                     "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
        asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr + 16),
                     "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]),
                     "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0) {
        const uint32_t idesc = make_idesc(N);
        const uint64_t b = make_desc(smem_u32(smem));
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it)
            for (int ks = 0; ks < 4; ++ks)
                asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;}" ::"r"(tmem),
                             "r"(tmem + 256 + ks * 8), "l"(b + (uint64_t)(ks * 2)), "r"(idesc), "r"((it | ks) ? 1u : 0u) : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t done = 0;
        while (!done) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p;}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
        cyc[0] = clock64() - t0;
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
        for (int j = 0; j < 16; ++j) D[(warp * 32 + lane) * N + c0 + j] = __uint_as_float(r[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}
int main() {
    for (int N : {64, 128, 256}) {
        float *A, *B, *D; long long* c;
        cudaMallocManaged(&A, 128 * 32 * 4); cudaMallocManaged(&B, N * 32 * 4); cudaMallocManaged(&D, 128 * N * 4); cudaMallocManaged(&c, 8);
        for (int i = 0; i < 128 * 32; ++i) A[i] = (float)((i * 7 + i / 32) % 13 - 6);
        for (int i = 0; i < N * 32; ++i) B[i] = (float)((i * 5 + i / 32 * 3) % 11 - 5);
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        for (int iters : {1, 1000}) {
            k<<<1, 128, 64 * 1024>>>(A, B, D, N, iters, c);
            cudaError_t e = cudaDeviceSynchronize();
            double maxerr = 0;
            for (int i = 0; i < 128; ++i)
                for (int n = 0; n < N; ++n) {
                    double ref = 0;
                    for (int kk = 0; kk < 32; ++kk) ref += (double)A[i * 32 + kk] * B[n * 32 + kk];
                    maxerr = fmax(maxerr, fabs(ref * iters - D[i * N + n]));
                }
            printf("TS mode N %3d iters %4d: max |err| %.3g, %.1f clk/MMA  %s\n", N, iters, maxerr, (double)c[0] / (iters * 4.0), cudaGetErrorString(e));
        }
    }
    return 0;
}
