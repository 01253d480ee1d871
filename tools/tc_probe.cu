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


// ---- probe 2: A in TMEM (tcgen05.st), B in natural time order (chunk (m,q) at 16 m + P q), N = 32, x16 loads ----
constexpr int N2 = 32, ROWS2 = N2 + 2, P2 = ROWS2 * 16;
__global__ void __launch_bounds__(128, 1) k_probe2(const float* __restrict__ A, const float* __restrict__ X, float* __restrict__ D, int neg)
{
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* b_hi = smem;                 // 4 q-subplanes of P2 bytes
    unsigned char* b_lo = smem + 4 * P2;
    __shared__ uint32_t tmem_base;
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 16 * ROWS2; i += 128) {
        const float v = X[i];
        const float hi = tf32_rn(v), lo = tf32_rn(v - hi);
        const int u = i >> 2, e = i & 3, m = u >> 2, q = u & 3;
        const int off = 16 * m + P2 * q + 4 * e;
        *reinterpret_cast<float*>(b_hi + off) = hi;
        *reinterpret_cast<float*>(b_lo + off) = lo;
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmem_base;
    // A row (= this thread's TMEM lane) -> columns 0..47 (hi) and 48..95 (lo); accumulator at columns 128..159
    {
        const uint32_t lane_base = tm + ((uint32_t)(warp * 32) << 16);
        for (int part = 0; part < 2; part++)
            for (int blk = 0; blk < 3; blk++) {
                uint32_t r[16];
                for (int j = 0; j < 16; j++) {
                    const float v = A[tid * K + blk * 16 + j];
                    const float hi = tf32_rn(v);
                    r[j] = __float_as_uint(part ? tf32_rn(v - hi) : hi);
                }
                asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(
                                 lane_base + (uint32_t)(part * 48 + blk * 16)),
                             "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
                             "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
                             : "memory");
            }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0) {
        uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N2 >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        if (neg) idesc |= (1u << 13);
        uint32_t acc = 0;
        for (int term = 0; term < 3; term++) {
            const uint32_t a_col = term == 0 ? 48 : 0;                       // lo*hi, hi*lo, hi*hi
            const unsigned char* bp = term == 1 ? b_lo : b_hi;
            for (int s = 0; s < K / 8; s++) {
                const uint64_t bd = make_desc(smem_u32(bp) + 16 * (s >> 1) + P2 * ((2 * s) & 3), P2, 128);
                asm volatile(
                    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tm + 128),
                    "r"(tm + a_col + 8 * s), "l"(bd), "r"(idesc), "r"(acc)
                    : "memory");
                acc = 1;
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
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
    const uint32_t lane_base = tm + ((uint32_t)(warp * 32) << 16) + 128;
    for (int half = 0; half < 2; half++) {
        uint32_t r[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                       "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(lane_base + 16 * half));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 16; j++) D[(size_t)tid * N2 + 16 * half + j] = __uint_as_float(r[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(256));
}

// ---- probe 3: kind::f16 (fp16 hi/lo split), A in TMEM as packed half pairs (column j = elements 2j, 2j+1), B natural order
// with 8-sample chunks: chunk (m, q) at 16 m + P q, q in {0,1}; one X row (16 samples) per K=16 MMA, row shift = 16 B ----
#include <cuda_fp16.h>
constexpr int N3 = 32, ROWS3 = N3 + 2, P3 = ROWS3 * 16;
__global__ void __launch_bounds__(128, 1) k_probe3(const float* __restrict__ A, const float* __restrict__ X, float* __restrict__ D)
{
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* b_hi = smem;                 // 2 q-subplanes of P3 bytes
    unsigned char* b_lo = smem + 2 * P3;
    __shared__ uint32_t tmem_base;
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 16 * ROWS3; i += 128) {
        const float v = X[i];
        const __half hi = __float2half_rn(v), lo = __float2half_rn(v - __half2float(hi));
        const int m = i >> 4, q = (i >> 3) & 1, e = i & 7;
        const int off = 16 * m + P3 * q + 2 * e;
        *reinterpret_cast<__half*>(b_hi + off) = hi;
        *reinterpret_cast<__half*>(b_lo + off) = lo;
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmem_base;
    {
        // A row -> columns 0..23 (hi pairs) and 24..47 (lo pairs): 48 columns = 3 x16 stores
        const uint32_t lane_base = tm + ((uint32_t)(warp * 32) << 16);
        uint32_t packed[48];
        for (int j = 0; j < 24; j++) {
            const float v0 = A[tid * K + 2 * j], v1 = A[tid * K + 2 * j + 1];
            const __half h0 = __float2half_rn(v0), h1 = __float2half_rn(v1);
            const __half l0 = __float2half_rn(v0 - __half2float(h0)), l1 = __float2half_rn(v1 - __half2float(h1));
            packed[j] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
            packed[24 + j] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
        }
        for (int blk = 0; blk < 3; blk++) {
            const uint32_t* r = packed + 16 * blk;
            asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(
                             lane_base + (uint32_t)(16 * blk)),
                         "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
                         "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
                         : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0) {
        // c_format F32 (1 << 4), a_format = b_format = F16 (0)
        const uint32_t idesc = (1u << 4) | ((uint32_t)(N3 >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        uint32_t acc = 0;
        for (int term = 0; term < 3; term++) {
            const uint32_t a_col = term == 0 ? 24 : 0;                       // lo*hi, hi*lo, hi*hi
            const unsigned char* bp = term == 1 ? b_lo : b_hi;
            for (int s = 0; s < K / 16; s++) {
                const uint64_t bd = make_desc(smem_u32(bp) + 16 * s, P3, 128);
                asm volatile(
                    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tm + 128),
                    "r"(tm + a_col + 8 * s), "l"(bd), "r"(idesc), "r"(acc)
                    : "memory");
                acc = 1;
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
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
    const uint32_t lane_base = tm + ((uint32_t)(warp * 32) << 16) + 128;
    for (int half = 0; half < 2; half++) {
        uint32_t r[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                       "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(lane_base + 16 * half));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 16; j++) D[(size_t)tid * N3 + 16 * half + j] = __uint_as_float(r[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(256));
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
    {
        std::vector<float> X(16 * ROWS2), D2(M * N2);
        for (auto& v : X) v = ((float)rand() / RAND_MAX * 2.f - 1.f) * 1000.f;
        float *dX2, *dD2;
        CK(cudaMalloc(&dX2, X.size() * 4));
        CK(cudaMalloc(&dD2, D2.size() * 4));
        CK(cudaMemcpy(dX2, X.data(), X.size() * 4, cudaMemcpyHostToDevice));
        for (int neg = 0; neg < 2; neg++) {
            k_probe2<<<1, 128, 8 * P2>>>(dA, dX2, dD2, neg);
            CK(cudaGetLastError());
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(D2.data(), dD2, D2.size() * 4, cudaMemcpyDeviceToHost));
            double err2 = 0, ref2 = 0;
            for (int m = 0; m < M; m++)
                for (int n = 0; n < N2; n++) {
                    double acc = 0;
                    for (int k = 0; k < K; k++) acc += (double)A[m * K + k] * (double)X[16 * n + k];
                    if (neg) acc = -acc;
                    const double e = (double)D2[m * N2 + n] - acc;
                    err2 += e * e;
                    ref2 += acc * acc;
                }
            const double snr = 10 * log10(ref2 / (err2 + 1e-300));
            printf("probe2 (A in TMEM, natural-order B, N=32, x16 ld) neg=%d: SNR %.1f dB\n", neg, snr);
            if (!(snr > 110.0)) bad++;
        }
    }
    {
        std::vector<float> X(16 * ROWS3), D3(M * N3);
        for (auto& v : X) v = ((float)rand() / RAND_MAX * 2.f - 1.f) * 1000.f;
        float *dX3, *dD3;
        CK(cudaMalloc(&dX3, X.size() * 4));
        CK(cudaMalloc(&dD3, D3.size() * 4));
        CK(cudaMemcpy(dX3, X.data(), X.size() * 4, cudaMemcpyHostToDevice));
        k_probe3<<<1, 128, 4 * P3>>>(dA, dX3, dD3);
        CK(cudaGetLastError());
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(D3.data(), dD3, D3.size() * 4, cudaMemcpyDeviceToHost));
        double err2 = 0, ref2 = 0;
        for (int m = 0; m < M; m++)
            for (int n = 0; n < N3; n++) {
                double acc = 0;
                for (int k = 0; k < K; k++) acc += (double)A[m * K + k] * (double)X[16 * n + k];
                const double e = (double)D3[m * N3 + n] - acc;
                err2 += e * e;
                ref2 += acc * acc;
            }
        const double snr = 10 * log10(ref2 / (err2 + 1e-300));
        printf("probe3 (kind::f16 hi/lo, A packed in TMEM, 8-sample chunks, N=32): SNR %.1f dB\n", snr);
        if (!(snr > 110.0)) bad++;
    }
    printf(bad ? "PROBE FAILED\n" : "PROBE OK\n");
    return bad;
}
