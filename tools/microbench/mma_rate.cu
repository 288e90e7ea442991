// Microbenchmark: back-to-back tcgen05.mma.kind::tf32 (M=128, N, K=8) from fixed shared-memory operands.
// Prints cycles per MMA for several N; build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 mma_rate.cu -o mma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ uint32_t make_idesc(int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__global__ void __launch_bounds__(832, 1) k(int N, int iters, int same, long long* out, int noise, volatile int* stop) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 832) ((float*)smem)[i] = 1.0f;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1u) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    __shared__ volatile int done_flag;
    if (threadIdx.x == 0) done_flag = 0;
    __syncthreads();
    if (warp >= 1 && warp <= noise) {
        // noise warps: STS.128 into a private 32 KB region (like the A producers) until the MMA thread finishes
        float4* dst = reinterpret_cast<float4*>(smem + 128 * 1024) + (threadIdx.x - 32) % 2048;
        float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
        while (!done_flag) { *dst = v; v.x += 1.f; __nanosleep(0); }
    }
    if (threadIdx.x == 0) {
        const uint32_t idesc = make_idesc(N);
        const uint32_t sA = smem_u32(smem), sB = sA + 64 * 1024;
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            // 4 k-steps x 3 MMAs like the engine: (a_lo,b_hi) (a_hi,b_lo) (a_hi,b_hi); `same` = reuse one operand pair
            const uint32_t a = sA + (same ? 0 : (it & 1) * 32768), b = sB + (same ? 0 : (it & 1) * 32768);
            const uint64_t a_hi = make_desc(a), a_lo = make_desc(a + 16384), b_hi = make_desc(b), b_lo = make_desc(b + N * 128);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                const uint64_t adv = ks * 2;
                asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;}" ::"r"(tmem), "l"(a_lo + adv), "l"(b_hi + adv), "r"(idesc), "r"(1u) : "memory");
                asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;}" ::"r"(tmem), "l"(a_hi + adv), "l"(b_lo + adv), "r"(idesc), "r"(1u) : "memory");
                asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;}" ::"r"(tmem), "l"(a_hi + adv), "l"(b_hi + adv), "r"(idesc), "r"(1u) : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t done = 0;
        while (!done) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p;}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
        const long long t1 = clock64();
        if (blockIdx.x == 0) out[0] = t1 - t0;
        done_flag = 1;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}
int main() {
    long long* d; cudaMalloc(&d, 8);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int iters = 2000;
    int* stop; cudaMalloc(&stop, 4);
    for (int grid : {148})
        for (int noise : {0, 4, 8, 16, 24})
            for (int N : {96, 128, 256}) {
                const int same = 0;
                k<<<grid, 832, 200 * 1024>>>(N, iters, same, d, noise, stop);
                long long c = 0; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
                cudaError_t e = cudaDeviceSynchronize();
                printf("noise %2d grid %3d same %d N %3d: %.1f clk/MMA (model %.1f)  %s\n", noise, grid, same, N, (double)c / (iters * 12.0), N / 2.0, cudaGetErrorString(e));
            }
    return 0;
}
