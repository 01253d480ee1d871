// peaks.cu -- device microbenchmarks behind bench.py's roofline denominators, measured live on the box the
// benchmark runs on (SURVEY.md section 8d asks for an FP32-FMA peak next to the HBM model; the tensor figures show
// what kind::tf32 / kind::f16 tcgen05 MMAs sustain when nothing else competes for the SM).
//   0  FP32 FMA issue peak (TFLOP/s, FMA = 2 flop): 8 independent chains per thread, every SM full
//   1  tcgen05.mma kind::tf32 dense peak (TFLOP/s): M128 N256 K8, operands in shared memory, one CTA per SM
//   2  tcgen05.mma kind::f16 dense peak (TFLOP/s):  M128 N256 K16
//   3  HBM copy bandwidth (GB/s, read + write bytes) over 1 GiB
#include "common.cuh"

namespace csdr {

__global__ void __launch_bounds__(256) k_peak_fma(float* out, int iters, float a, float b)
{
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = (float)(threadIdx.x + i) * 1e-3f;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
#pragma unroll
            for (int i = 0; i < 8; i++) v[i] = fmaf(v[i], a, b);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) s += v[i];
    if (s == 12345.678f) out[0] = s;          // never true; keeps the chains alive
}

__device__ __forceinline__ uint32_t pk_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t pk_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo)
{
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo >> 4) & 0x3fff) << 16) | ((uint64_t)((sbo >> 4) & 0x3fff) << 32) | (1ull << 46);
}

// one CTA per SM; warp 0's elected lane issues `iters` MMAs alternating between two 256-column accumulators
template <bool F16>
__global__ void __launch_bounds__(128, 1) k_peak_mma(int iters)
{
    extern __shared__ __align__(128) unsigned char pk_smem[];
    __shared__ uint32_t tmem_base;
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 48 * 1024 / 4; i += 128) reinterpret_cast<float*>(pk_smem)[i] = 0.f;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(pk_smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(pk_smem_u32(&tmem_base)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmem_base;
    if (tid == 0) {
        constexpr int N = 256;
        const uint32_t fmt = F16 ? 0u : ((2u << 7) | (2u << 10));
        const uint32_t idesc = (1u << 4) | fmt | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        // K-major, no swizzle: 8-row x 16-byte core matrices, LBO = 128 B between the two K chunks, SBO between row groups
        const uint64_t ad = pk_desc(pk_smem_u32(pk_smem), 128, 256);             // 128 rows x 32 B  = 4 KB
        const uint64_t bd = pk_desc(pk_smem_u32(pk_smem) + 8192, 128, 256);      // 256 rows x 32 B  = 8 KB
#pragma unroll 4
        for (int i = 0; i < iters; i++) {
            const uint32_t d = tm + (uint32_t)((i & 1) * N);
            if (F16)
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d),
                             "l"(ad), "l"(bd), "r"(idesc), "r"(1)
                             : "memory");
            else
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d),
                             "l"(ad), "l"(bd), "r"(idesc), "r"(1)
                             : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(pk_smem_u32(&bar)) : "memory");
        asm volatile(
            "{\n\t.reg .pred p;\n"
            "W_%=:\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0, 0x989680;\n\t"
            "@p bra D_%=;\n\t"
            "bra W_%=;\n"
            "D_%=:\n\t}\n" ::"r"(pk_smem_u32(&bar))
            : "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512));
}

__global__ void __launch_bounds__(256) k_peak_copy(const float4* __restrict__ in, float4* __restrict__ out, size_t n)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = in[i];
}

template <class F>
static int best_ms(F launch, int reps, float* best)
{
    cudaEvent_t a, b;
    CSDR_CK(cudaEventCreate(&a));
    CSDR_CK(cudaEventCreate(&b));
    float bm = 1e30f;
    for (int r = 0; r < reps + 1; r++) {
        CSDR_CK(cudaEventRecord(a, 0));
        launch();
        CSDR_CK(cudaEventRecord(b, 0));
        CSDR_CK(cudaEventSynchronize(b));
        CSDR_CK(cudaGetLastError());
        float ms = 0.f;
        CSDR_CK(cudaEventElapsedTime(&ms, a, b));
        if (r > 0 && ms < bm) bm = ms;      // first run warms up
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    *best = bm;
    return CUTESDR_OK;
}

}  // namespace csdr

using namespace csdr;

extern "C" int cutesdr_microbench(int device, int which, double* value)
{
    if (!value || which < 0 || which > 3) { set_error("microbench: bad arguments"); return CUTESDR_E_ARG; }
    CSDR_CK(cudaSetDevice(device));
    int sms = 148;
    CSDR_CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    float ms = 0.f;
    if (which == 0) {
        float* d = nullptr;
        CSDR_CK(cudaMalloc(&d, 64));
        const int iters = 4096, ctas = sms * 8;
        CSDR_TRY(best_ms([&]() { k_peak_fma<<<ctas, 256>>>(d, iters, 1.0000001f, 1e-7f); }, 5, &ms));
        cudaFree(d);
        *value = 2.0 * 64.0 * iters * 256.0 * ctas / (ms * 1e-3) / 1e12;
    } else if (which == 1 || which == 2) {
        const int iters = 16384;
        if (which == 1) {
            CSDR_CK(cudaFuncSetAttribute(k_peak_mma<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
            CSDR_TRY(best_ms([&]() { k_peak_mma<false><<<sms, 128, 64 * 1024>>>(iters); }, 5, &ms));
            *value = 2.0 * 128.0 * 256.0 * 8.0 * iters * sms / (ms * 1e-3) / 1e12;
        } else {
            CSDR_CK(cudaFuncSetAttribute(k_peak_mma<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
            CSDR_TRY(best_ms([&]() { k_peak_mma<true><<<sms, 128, 64 * 1024>>>(iters); }, 5, &ms));
            *value = 2.0 * 128.0 * 256.0 * 16.0 * iters * sms / (ms * 1e-3) / 1e12;
        }
    } else {
        const size_t bytes = (size_t)1 << 30;
        float4 *a = nullptr, *b = nullptr;
        CSDR_CK(cudaMalloc(&a, bytes));
        CSDR_CK(cudaMalloc(&b, bytes));
        CSDR_CK(cudaMemset(a, 1, bytes));
        CSDR_TRY(best_ms([&]() { k_peak_copy<<<sms * 16, 256>>>(a, b, bytes / sizeof(float4)); }, 5, &ms));
        cudaFree(a);
        cudaFree(b);
        *value = 2.0 * (double)bytes / (ms * 1e-3) / 1e9;
    }
    return CUTESDR_OK;
}
