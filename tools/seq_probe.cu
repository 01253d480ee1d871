// seq_probe.cu -- device microbenchmark behind the sequential burst kernels (post.cu): dependent-issue latency of the
// FP64 and integer operations a recurrence step is made of, on one warp per CTA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/seq_probe tools/seq_probe.cu && tools/bin/seq_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int N = 4096;

template <int OP>
__global__ void k_lat(double* out, long long* cyc, double a, double b, long long ia, int ib)
{
    double x = a + threadIdx.x * 1e-3, y = b;
    long long ix = ia + threadIdx.x;
    int i32 = ib;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) {
        if (OP == 0) x = fma(x, b, a);                       // DFMA chain
        else if (OP == 1) x = x + b;                         // DADD chain
        else if (OP == 2) x = (x > y) ? x * b : x + a;       // DSETP + select of two candidates
        else if (OP == 3) x = rint(x * b) + a;               // DMUL + DFRND + DADD
        else if (OP == 4) x = fmin(fmax(x + a, -b), b);      // DADD + 2 DMNMX
        else if (OP == 5) ix = ix + (ix >> 7) + ia;          // 64-bit integer adds
        else if (OP == 6) ix = (long long)(int)(ix >> 32) * (long long)ib + ix;   // IMAD.WIDE with 64-bit addend
        else if (OP == 7) ix = min(max(ix + ia, -(long long)ib), (long long)ib);  // 64-bit clamp
        else if (OP == 8) { float f = (float)x; x = (double)(f * 1.0001f) + a; }  // F2F round trips
        else if (OP == 9) x = (double)(long long)(x * b) + a;                     // D2I / I2D
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[OP] = t1 - t0;
    out[OP * 32 + threadIdx.x] = x + (double)ix + i32;
}

static void run_steps();
int main()
{
    double* d_out; long long* d_cyc;
    cudaMalloc(&d_out, 64 * 32 * sizeof(double));
    cudaMalloc(&d_cyc, 64 * sizeof(long long));
    cudaMemset(d_cyc, 0, 64 * sizeof(long long));
    const char* names[] = {"DFMA", "DADD", "DSETP+2cand+SEL", "DMUL+DFRND+DADD", "DADD+DMNMXx2", "IADD64 x2 + shift", "IMAD.WIDE+addend (hi word)",
                           "IADD64 + clamp64", "F2F.32<-64, FMUL, F2F.64<-32, DADD", "DMUL, D2I, I2D, DADD"};
#define RUN(OP) k_lat<OP><<<1, 32>>>(d_out, d_cyc, 1e-9, 0.999999, 12345, 77);
    for (int rep = 0; rep < 2; rep++) { RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9) }
    cudaDeviceSynchronize();
    long long h[64];
    cudaMemcpy(h, d_cyc, sizeof(h), cudaMemcpyDeviceToHost);
    for (int i = 0; i < 10; i++) printf("%-40s %7.1f cycles per iteration\n", names[i], (double)h[i] / N);
    run_steps();
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

// ---- the recurrence steps of post.cu as they are written there, one warp, inputs from shared memory
__device__ __forceinline__ double in_reg(double v) { asm volatile("" : "+d"(v)); return v; }
template <int WHICH>
__global__ void k_steps(double* out, long long* cyc, double alpha, double beta, double hi_, double lo_, double kdc, double gain,
                        double a1, double a2, double b0, double b1, double b2)
{
    __shared__ double th[32 * 17];
    for (int i = threadIdx.x; i < 32 * 17; i += blockDim.x) th[i] = 3.0 * sin(0.37 * i);
    __syncthreads();
    const double kTwoPi = 6.283185307179586, kInv2Pi = 1.0 / kTwoPi, kM = 6755399441055744.0;
    double phase = 0.1 * threadIdx.x, freq = 0, fm_dc = 0, w1 = 0, w2 = 0, lp = 0, acc = 0;
    double att = -5, dec = -5; int hang_timer = 0;
    double fprev = 0;
    const double hi = in_reg(hi_), lo = in_reg(lo_), qdc = 1.0 - kdc, na1 = -a1, na2 = -a2;
    const bool use_hang = (threadIdx.x & 7) == 0;
    long long t0 = clock64();
    for (int it = 0; it < N / 16; it++) {
#pragma unroll
        for (int j = 0; j < 16; j++) {
            const double v = th[(threadIdx.x & 31) * 17 + j];
            if (WHICH == 0 || WHICH == 1) {
                const double x = v + phase;
                const double r = fma(x, kInv2Pi, kM) - kM;
                const double err = fma(r, kTwoPi, -x);
                double f = fma(beta, err, freq);
                const double pa = fma(alpha, err, phase);
                f = f > hi ? hi : f;
                f = f < lo ? lo : f;
                freq = f;
                phase = pa + f;
                if (WHICH == 1) {
                    fm_dc = fma(qdc, fm_dc, kdc * f);
                    const double pre = (f - fm_dc) * gain;
                    const double w0 = fma(na1, w1, fma(na2, w2, pre));
                    lp = fma(b0, w0, fma(b1, w1, b2 * w2));
                    w2 = w1; w1 = w0;
                    acc += pre;
                }
            } else if (WHICH == 3) {
                // post-processing of the previous sample's frequency first (independent of this sample's PLL chain)
                {
                    const double f = fprev;
                    fm_dc = fma(qdc, fm_dc, kdc * f);
                    const double pre = (f - fm_dc) * gain;
                    const double w0 = fma(na1, w1, fma(na2, w2, pre));
                    lp = fma(b0, w0, fma(b1, w1, b2 * w2));
                    w2 = w1; w1 = w0;
                    acc += pre;
                }
                const double x = v + phase;
                const double r = fma(x, kInv2Pi, kM) - kM;
                const double err = fma(r, kTwoPi, -x);
                double f = fma(beta, err, freq);
                const double pa = fma(alpha, err, phase);
                f = f > hi ? hi : f;
                f = f < lo ? lo : f;
                freq = f;
                phase = pa + f;
                fprev = f;
            } else if (WHICH == 2) {
                const double peak = v;
                const double ar = fma(1 - alpha, att, alpha * peak), af = fma(1 - beta, att, beta * peak);
                const double dr = fma(1 - kdc, dec, kdc * peak), df = fma(1 - gain, dec, gain * peak);
                const bool ga = peak > att, gd = peak > dec;
                const bool hold = use_hang && !gd && hang_timer < 100;
                att = ga ? ar : af;
                dec = gd ? dr : (hold ? dec : df);
                hang_timer = use_hang ? (gd ? 0 : hang_timer + (hold ? 1 : 0)) : hang_timer;
                acc += att > dec ? att : dec;
            }
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[WHICH] = t1 - t0;
    out[threadIdx.x] = phase + freq + lp + acc + att + dec + hang_timer + fprev;
}

// FP64 pipe throughput of one SM: W warps, 8 independent DFMA chains each
__global__ void k_tput(double* out, long long* cyc, double a, double b)
{
    double x[8];
    for (int i = 0; i < 8; i++) x[i] = a * (i + 1) + threadIdx.x;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < N; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) x[i] = fma(x[i], b, a);
    }
    long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < 8; i++) s += x[i];
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

static void run_steps()
{
    {
        double* d_out; long long* d_cyc;
        cudaMalloc(&d_out, 2048 * sizeof(double));
        cudaMalloc(&d_cyc, 8 * sizeof(long long));
        for (int w = 1; w <= 16; w *= 2) {
            long long h = 0;
            for (int rep = 0; rep < 2; rep++) k_tput<<<1, 32 * w>>>(d_out, d_cyc, 1e-9, 0.999999);
            cudaMemcpy(&h, d_cyc, sizeof(h), cudaMemcpyDeviceToHost);
            printf("DFMA throughput, %2d warps on one SM:      %6.3f warp instructions per clock\n", w, (double)w * 8 * N / (double)h);
        }
        {
            long long h[8];
            for (int rep = 0; rep < 2; rep++) k_steps<3><<<1, 32>>>(d_out, d_cyc, 1.09, 0.6, 0.77, -0.77, 2e-3, 32000.0, -1.5, 0.6, 0.03, 0.06, 0.03);
            cudaMemcpy(h, d_cyc, sizeof(h), cudaMemcpyDeviceToHost);
            printf("FM step, post-processing one sample behind: %7.1f cycles per sample\n", (double)h[3] / N);
        }
        for (int w = 1; w <= 8; w *= 2) {
            long long h[8];
            for (int rep = 0; rep < 2; rep++) k_steps<1><<<1, 32 * w>>>(d_out, d_cyc, 1.09, 0.6, 0.77, -0.77, 2e-3, 32000.0, -1.5, 0.6, 0.03, 0.06, 0.03);
            cudaMemcpy(h, d_cyc, sizeof(h), cudaMemcpyDeviceToHost);
            printf("FM step, %d warps on one SM:               %7.1f cycles per sample\n", w, (double)h[1] / N);
        }
    }

    {
        double* d_out; long long* d_cyc;
        cudaMalloc(&d_out, 64 * sizeof(double));
        cudaMalloc(&d_cyc, 8 * sizeof(long long));
        for (int rep = 0; rep < 2; rep++) {
            k_steps<0><<<1, 32>>>(d_out, d_cyc, 1.09, 0.6, 0.77, -0.77, 2e-3, 32000.0, -1.5, 0.6, 0.03, 0.06, 0.03);
            k_steps<1><<<1, 32>>>(d_out, d_cyc, 1.09, 0.6, 0.77, -0.77, 2e-3, 32000.0, -1.5, 0.6, 0.03, 0.06, 0.03);
            k_steps<2><<<1, 32>>>(d_out, d_cyc, 0.01, 0.004, 0.77, -0.77, 2e-3, 1e-4, -1.5, 0.6, 0.03, 0.06, 0.03);
        }
        cudaDeviceSynchronize();
        long long h[8];
        cudaMemcpy(h, d_cyc, sizeof(h), cudaMemcpyDeviceToHost);
        printf("PLL step alone                           %7.1f cycles per sample\n", (double)h[0] / N);
        printf("FM step (PLL + DC tracker + biquad)      %7.1f cycles per sample\n", (double)h[1] / N);
        printf("AGC averager step                        %7.1f cycles per sample\n", (double)h[2] / N);
    }

}
