// tc_probe.cu -- stand-alone check of the tcgen05 conventions kernel 1T relies on (not part of the library):
//   * kind::tf32, M=128 N=128, operands in shared memory, K-major, SWIZZLE_NONE ("interleave") descriptors
//   * A plane:  addr(row, chunk q, e) = 16*(row%8) + 1536*(row/8) + 128*q + 4*e          (LBO 128, SBO 1536)
//   * B plane:  8 interleaved strips, addr(strip r, chunk w, e) = 16*r + 128*w + 4*e; MMA row n = 8g + r reads
//               chunks w = 4g + q  (LBO 128, SBO 512; K-step s starts 256*s bytes in) -- a Hankel operand
//   * a_negate, accumulate flag, 3xTF32 split accuracy, TMEM alloc / commit / 32x32b.x1 loads
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tc_probe tc_probe.cu ; run on a B200.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d: %s\n", #x, __LINE__, cudaGetErrorString(e)); return 1; } } while (0)

constexpr int M = 128, N = 128, K = 48;
constexpr int A_PLANE = 16 * 1536;      // bytes
constexpr int B_PLANE = 72 * 128;       // bytes

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo)
{
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo >> 4) & 0x3fff) << 16) | ((uint64_t)((sbo >> 4) & 0x3fff) << 32) |
           (1ull << 46);
}

__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ float tf32_rn(float x)
{
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// mode 0: single tf32 product; 1: 3xTF32; 2: 3xTF32 with A negated
__global__ void __launch_bounds__(128, 1) k_probe(const float* __restrict__ A, const float* __restrict__ Xs, float* __restrict__ D, int mode)
{
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* a_hi = smem;
    unsigned char* a_lo = a_hi + A_PLANE;
    unsigned char* b_hi = a_lo + A_PLANE;
    unsigned char* b_lo = b_hi + B_PLANE;
    __shared__ uint32_t tmem_base;
    __shared__ __align__(8) uint64_t bar;

    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < M * K; i += 128) {
        const int row = i / K, k = i % K;
        const float v = A[i];
        const float hi = tf32_rn(v), lo = tf32_rn(v - hi);
        const int off = 16 * (row & 7) + 1536 * (row >> 3) + 128 * (k >> 2) + 4 * (k & 3);
        *reinterpret_cast<float*>(a_hi + off) = hi;
        *reinterpret_cast<float*>(a_lo + off) = lo;
    }
    for (int i = tid; i < 8 * 288; i += 128) {
        const int r = i / 288, j = i % 288;
        const float v = Xs[i];
        const float hi = tf32_rn(v), lo = tf32_rn(v - hi);
        const int off = 16 * r + 128 * (j >> 2) + 4 * (j & 3);
        *reinterpret_cast<float*>(b_hi + off) = hi;
        *reinterpret_cast<float*>(b_lo + off) = lo;
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmem_base;

    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        const uint32_t idesc_neg = idesc | (1u << 13);
        uint32_t acc = 0;
        const int nterm = mode == 0 ? 1 : 3;
        for (int term = 0; term < nterm; term++) {
            // small terms first: lo*hi, hi*lo, then hi*hi
            const unsigned char* ap = (nterm == 1 || term == 2) ? a_hi : (term == 0 ? a_lo : a_hi);
            const unsigned char* bp = (nterm == 1 || term == 2) ? b_hi : (term == 0 ? b_hi : b_lo);
            for (int s = 0; s < K / 8; s++) {
                const uint64_t ad = make_desc(smem_u32(ap) + 256 * s, 128, 1536);
                const uint64_t bd = make_desc(smem_u32(bp) + 256 * s, 128, 512);
                mma_tf32(tm, ad, bd, mode == 2 ? idesc_neg : idesc, acc);
                acc = 1;
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    // wait for the MMAs
    {
        uint32_t done = 0;
        while (!done) {
            asm volatile(
                "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n"
                : "=r"(done)
                : "r"(smem_u32(&bar)), "r"(0)
                : "memory");
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t lane_base = tm + ((uint32_t)(warp * 32) << 16);
    for (int col = 0; col < N; col++) {
        uint32_t v;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(lane_base + col));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        D[(size_t)tid * N + col] = __uint_as_float(v);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(128));
}

int main()
{
    std::vector<float> A(M * K), Xs(8 * 288), D(M * N);
    srand(7);
    for (auto& v : A) v = (float)rand() / RAND_MAX * 2.f - 1.f;
    for (auto& v : Xs) v = ((float)rand() / RAND_MAX * 2.f - 1.f) * 1000.f;
    float *dA, *dX, *dD;
    CK(cudaMalloc(&dA, A.size() * 4));
    CK(cudaMalloc(&dX, Xs.size() * 4));
    CK(cudaMalloc(&dD, D.size() * 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dX, Xs.data(), Xs.size() * 4, cudaMemcpyHostToDevice));
    const int smem = 2 * A_PLANE + 2 * B_PLANE;
    CK(cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int bad = 0;
    for (int mode = 0; mode < 3; mode++) {
        CK(cudaMemset(dD, 0, D.size() * 4));
        k_probe<<<1, 128, smem>>>(dA, dX, dD, mode);
        CK(cudaGetLastError());
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
        double err2 = 0, ref2 = 0, worst = 0;
        for (int m = 0; m < M; m++)
            for (int n = 0; n < N; n++) {
                const int g = n >> 3, r = n & 7;
                double acc = 0;
                for (int k = 0; k < K; k++) acc += (double)A[m * K + k] * (double)Xs[r * 288 + 16 * g + k];
                if (mode == 2) acc = -acc;
                const double e = (double)D[m * N + n] - acc;
                err2 += e * e;
                ref2 += acc * acc;
                if (fabs(e) > worst) worst = fabs(e);
            }
        const double snr = 10 * log10(ref2 / (err2 + 1e-300));
        printf("mode %d: SNR %.1f dB, worst abs err %.3g (rms ref %.3g)\n", mode, snr, worst, sqrt(ref2 / (M * N)));
        const double need = mode == 0 ? 55.0 : 110.0;
        if (!(snr > need)) bad++;
    }
    printf(bad ? "PROBE FAILED\n" : "PROBE OK\n");
    return bad;
}
