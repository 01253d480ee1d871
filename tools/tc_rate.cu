// tc_rate.cu -- tcgen05 kind::tf32 issue-rate microbenchmark (debug aid): cycles per M128 x N x K8 MMA for
// N in {32, 64, 128, 256}, A from shared memory (SS) or TMEM (TS), one CTA per SM on every SM.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d: %s\n", #x, __LINE__, cudaGetErrorString(e)); return 1; } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo)
{
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo >> 4) & 0x3fff) << 16) | ((uint64_t)((sbo >> 4) & 0x3fff) << 32) | (1ull << 46);
}
__global__ void __launch_bounds__(128, 1) k_rate(int N, int ts, int iters, int nacc, int nw, long long* out)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint32_t tmem_base;
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 48 * 1024 / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 0.f;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(nw));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmem_base;
    if ((tid & 31) == 0 && warp < nw) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t ad = make_desc(smem_u32(smem), 128, 1536);
        const uint64_t bd = make_desc(smem_u32(smem) + 24576, 128, 256);
        long long t0 = clock64();
        const uint32_t nmask = (uint32_t)nacc - 1;
#pragma unroll 8
        for (int i = 0; i < iters; i++) {
            const uint32_t d = tm + 64 + (uint32_t)(warp * 96) + ((uint32_t)i & nmask) * (uint32_t)N;
            if (ts)
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d),
                             "r"(tm + (uint32_t)(8 * (i & 3))), "l"(bd), "r"(idesc), "r"(1)
                             : "memory");
            else
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d),
                             "l"(ad), "l"(bd), "r"(idesc), "r"(1)
                             : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
        long long t1 = clock64();
        if (blockIdx.x == 0 && warp == 0) *out = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512));
}
int main()
{
    long long* d_out;
    CK(cudaMalloc(&d_out, 8));
    CK(cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    const int iters = 4096;
    for (int ts = 0; ts < 2; ts++)
        for (int N = 32; N <= 64; N *= 2)
            for (int nw = 1; nw <= 4; nw *= 2) {
                k_rate<<<148, 128, 64 * 1024>>>(N, ts, iters, 1, nw, d_out);
                CK(cudaGetLastError());
                CK(cudaDeviceSynchronize());
                long long c = 0;
                CK(cudaMemcpy(&c, d_out, 8, cudaMemcpyDeviceToHost));
                printf("%s N=%3d, %d issuing warps: %.1f cycles per MMA overall (floor %d)\n", ts ? "TS" : "SS", N, nw, (double)c / (iters * nw), 128 * N / 256);
            }
    return 0;
}
