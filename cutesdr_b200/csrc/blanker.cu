// blanker.cu -- kernel 7: CNoiseProc::ProcessBlanker (dsp/noiseproc.cpp:121-176) on the shared
// wideband stream, restated as scans so a block of 10^6 samples runs in parallel:
//   mag[n]   = max(|I|,|Q|)
//   sum[n]   = sum of the last MagSamples+1 mags (the reference's ring holds MagSamples+1 entries,
//              :143-147)  = carry + prefix_sum(mag[n] - mag[n-(MagSamples+1)])        (double)
//   trig[n]  = mag[n]*Ratio > sum[n]                                                    (:152)
//   blank[n] = a trigger in (n-WidthSamples, n]   = n - prefix_max(trig ? n : -inf) < WidthSamples
//   out[n]   = blank ? 0 : x[n-(DelaySamples+1)]   (delay ring holds DelaySamples+1 entries, :149-151)
// Histories live in power-of-two rings addressed by absolute stream position.
#include "blanker.cuh"

#include <limits.h>

namespace csdr {

constexpr int kScanChunk = 2048;      // elements per CTA (256 threads x 8)

struct OpAdd { __device__ static double id() { return 0.0; } __device__ static double ap(double a, double b) { return a + b; } };
struct OpMax {
    __device__ static long long id() { return LLONG_MIN; }
    __device__ static long long ap(long long a, long long b) { return a > b ? a : b; }
};

template <typename T, typename Op>
__device__ __forceinline__ T block_exclusive(T v, T* total)
{
    // exclusive scan of one value per thread over a 256-thread CTA
    __shared__ T warp_tot[8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    T inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        T o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc = Op::ap(o, inc);
    }
    if (lane == 31) warp_tot[wid] = inc;
    __syncthreads();
    T base = Op::id();
    for (int w = 0; w < wid; w++) base = Op::ap(base, warp_tot[w]);
    T tot = Op::id();
    for (int w = 0; w < 8; w++) tot = Op::ap(tot, warp_tot[w]);
    if (total) *total = tot;
    T prev = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane == 0) prev = Op::id();
    __syncthreads();
    return Op::ap(base, prev);
}

template <typename T, typename Op>
__global__ void __launch_bounds__(256) k_scan_reduce(const T* __restrict__ in, T* __restrict__ part, int n)
{
    const int base = blockIdx.x * kScanChunk + threadIdx.x * 8;
    T acc = Op::id();
#pragma unroll
    for (int k = 0; k < 8; k++) if (base + k < n) acc = Op::ap(acc, in[base + k]);
    T tot;
    block_exclusive<T, Op>(acc, &tot);
    if (threadIdx.x == 0) part[blockIdx.x] = tot;
}

// single CTA: part[i] <- carry (+) part[0..i-1]
template <typename T, typename Op>
__global__ void __launch_bounds__(256) k_scan_parts(T* __restrict__ part, int nparts, const T* __restrict__ carry)
{
    __shared__ T running;
    if (threadIdx.x == 0) running = *carry;
    __syncthreads();
    for (int b0 = 0; b0 < nparts; b0 += 256) {
        const int i = b0 + threadIdx.x;
        T v = i < nparts ? part[i] : Op::id();
        T tot;
        T ex = block_exclusive<T, Op>(v, &tot);
        const T run = running;
        if (i < nparts) part[i] = Op::ap(run, ex);
        __syncthreads();
        if (threadIdx.x == 0) running = Op::ap(run, tot);
        __syncthreads();
    }
}

template <typename T, typename Op>
__global__ void __launch_bounds__(256) k_scan_final(const T* in, T* out, const T* __restrict__ part, int n)   // in may alias out
{
    const int base = blockIdx.x * kScanChunk + threadIdx.x * 8;
    T v[8];
    T acc = Op::id();
#pragma unroll
    for (int k = 0; k < 8; k++) { v[k] = base + k < n ? in[base + k] : Op::id(); acc = Op::ap(acc, v[k]); }
    T ex = block_exclusive<T, Op>(acc, nullptr);
    T run = Op::ap(part[blockIdx.x], ex);
#pragma unroll
    for (int k = 0; k < 8; k++) { run = Op::ap(run, v[k]); if (base + k < n) out[base + k] = run; }
}

__device__ __forceinline__ float peak_mag(float2 x) { return fmaxf(fabsf(x.x), fabsf(x.y)); }

__global__ void k_nb_prep(const float2* __restrict__ in, int n, long long pos, float2* __restrict__ xr, unsigned xmask,
                          float* __restrict__ magr, unsigned mmask, int wn, double* __restrict__ diff)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float2 x = in[i];
    const long long a = pos + i;
    xr[a & xmask] = x;
    const float mag = peak_mag(x);
    magr[a & mmask] = mag;
    float old = 0.f;
    if (a - wn >= 0) old = (i - wn >= 0) ? peak_mag(in[i - wn]) : magr[(a - wn) & mmask];
    diff[i] = (double)mag - (double)old;
}

__global__ void k_nb_trig(const float* __restrict__ magr, unsigned mmask, long long pos, int n, const double* __restrict__ sum,
                          double ratio, long long* __restrict__ tpos)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long a = pos + i;
    const double mag = (double)magr[a & mmask];
    tpos[i] = (mag * ratio > sum[i]) ? a : LLONG_MIN;
}

__global__ void k_nb_out(const float2* __restrict__ xr, unsigned xmask, long long pos, int n, const long long* __restrict__ last,
                         int width, int delay, float2* __restrict__ out, const double* __restrict__ sum,
                         double* __restrict__ carry_sum, long long* __restrict__ carry_last)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long a = pos + i;
    const long long lt = last[i];
    const bool blank = (lt != LLONG_MIN) && (a - lt < width);
    float2 v = make_float2(0.f, 0.f);
    if (!blank && a - delay >= 0) v = xr[(a - delay) & xmask];
    out[i] = v;
    if (i == n - 1) { *carry_sum = sum[i]; *carry_last = lt; }
}

Blanker::~Blanker()
{
    cudaFree(d_mag_); cudaFree(d_xh_); cudaFree(d_scan_); cudaFree(d_last_); cudaFree(d_part_); cudaFree(d_partl_);
    cudaFree(d_carry_); cudaFree(d_carryl_);
}

int Blanker::init(int max_block, cudaStream_t st, LaunchCounter* lc)
{
    max_block_ = max_block; st_ = st; lc_ = lc;
    CSDR_CK(cudaMalloc(&d_scan_, (size_t)max_block * sizeof(double)));
    CSDR_CK(cudaMalloc(&d_last_, (size_t)max_block * sizeof(long long)));
    const int nparts = (max_block + kScanChunk - 1) / kScanChunk;
    CSDR_CK(cudaMalloc(&d_part_, nparts * sizeof(double)));
    CSDR_CK(cudaMalloc(&d_partl_, nparts * sizeof(long long)));
    CSDR_CK(cudaMalloc(&d_carry_, sizeof(double)));
    CSDR_CK(cudaMalloc(&d_carryl_, sizeof(long long)));
    return CUTESDR_OK;
}

int Blanker::setup(bool on, double threshold, double width_us, double sample_rate)
{
    // CNoiseProc::SetupBlanker, dsp/noiseproc.cpp:77-119. Its change test compares SampleRate
    // with itself (:84), so a rate-only change is ignored there; the bank's rate is fixed anyway.
    if (configured_ && threshold == threshold_ && width_us == width_ && on == on_) return CUTESDR_OK;
    configured_ = true;
    on_ = on; threshold_ = threshold; width_ = width_us; rate_ = sample_rate;
    width_samples_ = (int)(width_us * 1e-6 * sample_rate);
    if (width_samples_ < 1) width_samples_ = 1;
    else if (width_samples_ > 4096) width_samples_ = 4096;          // MAX_WIDTH
    mag_samples_ = (int)(.005 * sample_rate);                      // MAGAVE_TIME
    ratio_ = .005 * threshold * (double)mag_samples_;
    delay_samples_ = width_samples_ / 2;
    // rings sized for history + one block; all state restarts (the reference zeroes its buffers)
    cudaFree(d_mag_); cudaFree(d_xh_);
    d_mag_ = nullptr; d_xh_ = nullptr;
    hist_mag_ = next_pow2((long long)mag_samples_ + 1 + max_block_);
    hist_x_ = next_pow2((long long)delay_samples_ + 1 + max_block_);
    CSDR_CK(cudaMalloc(&d_mag_, (size_t)hist_mag_ * sizeof(float)));
    CSDR_CK(cudaMalloc(&d_xh_, (size_t)hist_x_ * sizeof(float2)));
    CSDR_CK(cudaMemsetAsync(d_mag_, 0, (size_t)hist_mag_ * sizeof(float), st_));
    CSDR_CK(cudaMemsetAsync(d_xh_, 0, (size_t)hist_x_ * sizeof(float2), st_));
    CSDR_CK(cudaMemsetAsync(d_carry_, 0, sizeof(double), st_));
    const long long none = LLONG_MIN;
    CSDR_CK(cudaMemcpyAsync(d_carryl_, &none, sizeof(long long), cudaMemcpyHostToDevice, st_));
    CSDR_CK(cudaStreamSynchronize(st_));
    pos_ = 0;
    return CUTESDR_OK;
}

int Blanker::run(const float2* d_in, float2* d_out, int n)
{
    if (n <= 0) return CUTESDR_OK;
    if (n > max_block_) { set_error("blanker: block of %d exceeds capacity %d", n, max_block_); return CUTESDR_E_ARG; }
    if (!on_) {      // pass-through (dsp/noiseproc.cpp:125-129)
        if (d_out != d_in) CSDR_CK(cudaMemcpyAsync(d_out, d_in, (size_t)n * sizeof(float2), cudaMemcpyDeviceToDevice, st_));
        return CUTESDR_OK;
    }
    const int tb = 256, gb = (n + tb - 1) / tb;
    const int nparts = (n + kScanChunk - 1) / kScanChunk;
    k_nb_prep<<<gb, tb, 0, st_>>>(d_in, n, pos_, d_xh_, (unsigned)(hist_x_ - 1), d_mag_, (unsigned)(hist_mag_ - 1),
                                  mag_samples_ + 1, d_scan_);
    k_scan_reduce<double, OpAdd><<<nparts, 256, 0, st_>>>(d_scan_, d_part_, n);
    k_scan_parts<double, OpAdd><<<1, 256, 0, st_>>>(d_part_, nparts, d_carry_);
    k_scan_final<double, OpAdd><<<nparts, 256, 0, st_>>>(d_scan_, d_scan_, d_part_, n);
    k_nb_trig<<<gb, tb, 0, st_>>>(d_mag_, (unsigned)(hist_mag_ - 1), pos_, n, d_scan_, ratio_, d_last_);
    k_scan_reduce<long long, OpMax><<<nparts, 256, 0, st_>>>(d_last_, d_partl_, n);
    k_scan_parts<long long, OpMax><<<1, 256, 0, st_>>>(d_partl_, nparts, d_carryl_);
    k_scan_final<long long, OpMax><<<nparts, 256, 0, st_>>>(d_last_, d_last_, d_partl_, n);
    k_nb_out<<<gb, tb, 0, st_>>>(d_xh_, (unsigned)(hist_x_ - 1), pos_, n, d_last_, width_samples_, delay_samples_ + 1, d_out,
                                 d_scan_, d_carry_, d_carryl_);
    lc_->n += 9;
    CSDR_CK(cudaGetLastError());
    pos_ += n;
    return CUTESDR_OK;
}

}  // namespace csdr
