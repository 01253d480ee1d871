// resampler.cu -- kernel 5: batched CFractResampler (dsp/fractresampler.cpp:144-352).
//
// Every output is an independent 28-tap dot product once its fractional input time is known, so
// the grid is (outputs x rows). The time sequence itself is the reference's double accumulator
// (t += Rate; t -= InLength per call): it is identical for all rows of a bank, so the host steps
// it once per call and ships the times; the kernel turns (j - t)*10000 into the truncated table
// index with the same double operations the reference uses, so table lookups are identical.
#include "resampler.cuh"

namespace csdr {

void ResampleClock::advance(int n_in, double rate, std::vector<double>& times)
{
    times.clear();
    int it = (int)t_;
    while (it < n_in) {
        times.push_back(t_);
        t_ += rate;
        it = (int)t_;
    }
    t_ -= (double)n_in;
}

// Pass 1 (one small launch): every row of a bank shares the output times, hence the truncated table
// indices and the 28 weights of every output. They are looked up ONCE into wts[k][28] (+ the integer
// input position), turning the per-row work of pass 2 into a dense 28-tap dot product without gathers.
__global__ void __launch_bounds__(128) k_resample_weights(const double* __restrict__ times, int n_out,
                                                          const float* __restrict__ sinc, float* __restrict__ wts,
                                                          int* __restrict__ pos)
{
    const int k = blockIdx.x * 4 + (threadIdx.x >> 5);
    const int i = (threadIdx.x & 31) + 1;
    if (k >= n_out) return;
    const double t = times[k];
    const int it = (int)t;
    if (i == 1) pos[k] = it;
    if (i <= kRsPeriods) {
        const int j = it + i;
        const int s = (int)(((double)j - t) * (double)kRsPts);      // dsp/fractresampler.cpp:168
        wts[(size_t)k * kRsPeriods + (i - 1)] = __ldg(sinc + s);
    }
}

// Pass 2: w: [nrows][row_len], row = 28 carried samples then the new inputs (W floats per sample). Thread per
// (output, row); consecutive lanes take consecutive outputs of one row (inputs and weights stream, no table access).
template <int W>
__global__ void __launch_bounds__(128) k_resample(const float* __restrict__ w, int row_len, int nrows,
                                                  const float* __restrict__ wts, const int* __restrict__ pos, int n_out,
                                                  float* __restrict__ out, int out_stride,
                                                  int out_off, const int* __restrict__ row_map,
                                                  int16_t* __restrict__ out16, float gain, int interleave16)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (k >= n_out) return;
    const float* x = w + (size_t)r * row_len + (size_t)W * pos[k];
    const float4* c4 = reinterpret_cast<const float4*>(wts + (size_t)k * kRsPeriods);
    float acc[W];
#pragma unroll
    for (int e = 0; e < W; e++) acc[e] = 0.f;
#pragma unroll
    for (int q = 0; q < kRsPeriods / 4; q++) {
        const float4 c = __ldg(c4 + q);
        // same accumulation order as the reference loop i = 1..28
#pragma unroll
        for (int e = 0; e < W; e++) {
            acc[e] = fmaf(x[W * (4 * q + 1) + e], c.x, acc[e]);
            acc[e] = fmaf(x[W * (4 * q + 2) + e], c.y, acc[e]);
            acc[e] = fmaf(x[W * (4 * q + 3) + e], c.z, acc[e]);
            acc[e] = fmaf(x[W * (4 * q + 4) + e], c.w, acc[e]);
        }
    }
    if (out16) {
#pragma unroll
        for (int e = 0; e < W; e++) {
            float v = acc[e] * gain;                                // :228-239 gain, clip, truncate
            v = fminf(fmaxf(v, -32767.0f), 32767.0f);
            out16[(size_t)k * interleave16 + W * r + e] = (int16_t)v;
        }
    } else {
        const int row = row_map ? row_map[r] : r;
#pragma unroll
        for (int e = 0; e < W; e++) out[(size_t)row * out_stride + (size_t)W * (out_off + k) + e] = acc[e];
    }
}

// The output time sequence, stepped exactly like the reference's accumulator (same IEEE double adds
// as ResampleClock::advance on the host, which supplies the start time and the count).
__global__ void k_resample_times(double t0, double rate, int m, double* __restrict__ times)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double t = t0;
    for (int k = 0; k < m; k++) { times[k] = t; t += rate; }
}

// carry the last 28 inputs of every row to the row's front (dsp/fractresampler.cpp:180-182)
__global__ void k_resample_carry(float* w, int row_len, int nrows, int n_in, int width)
{
    const int r = blockIdx.x;
    const int i = threadIdx.x;
    float v = 0.f;
    if (i < width * kRsPeriods) v = w[(size_t)r * row_len + (size_t)width * n_in + i];
    __syncthreads();
    if (i < width * kRsPeriods) w[(size_t)r * row_len + i] = v;
}

ResamplerBank::~ResamplerBank()
{
    cudaFree(d_w_);
    cudaFree(d_sinc_);
    cudaFree(d_times_);
    cudaFree(d_wts_);
    cudaFree(d_pos_);
}

int ResamplerBank::init(int nrows, int max_in, cudaStream_t st, LaunchCounter* lc, int width)
{
    if (width != 1 && width != 2) { set_error("resampler: width %d", width); return CUTESDR_E_ARG; }
    nrows_ = nrows; max_in_ = max_in; st_ = st; lc_ = lc; width_ = width;
    row_len_ = round_up(width * (kRsPeriods + max_in), 4);
    CSDR_CK(cudaMalloc(&d_w_, (size_t)nrows * row_len_ * sizeof(float)));
    CSDR_CK(cudaMemsetAsync(d_w_, 0, (size_t)nrows * row_len_ * sizeof(float), st_));
    // window-sinc table, dsp/fractresampler.cpp:104-115 (computed in double, stored float32)
    std::vector<float> sinc(kRsLen);
    for (int i = 0; i < kRsLen; i++) {
        const double win = (0.35875 - 0.48829 * cos((kTwoPi * i) / (kRsLen - 1)) +
                            0.14128 * cos((2.0 * kTwoPi * i) / (kRsLen - 1)) -
                            0.01168 * cos((3.0 * kTwoPi * i) / (kRsLen - 1)));
        const double fi = kPi * (double)(i - kRsLen / 2) / (double)kRsPts;
        sinc[i] = (float)((i != kRsLen / 2) ? win * sin(fi) / fi : 1.0);
    }
    CSDR_CK(cudaMalloc(&d_sinc_, kRsLen * sizeof(float)));
    CSDR_CK(cudaMemcpyAsync(d_sinc_, sinc.data(), kRsLen * sizeof(float), cudaMemcpyHostToDevice, st_));
    CSDR_CK(cudaStreamSynchronize(st_));
    clk_.reset();
    return CUTESDR_OK;
}

int ResamplerBank::reset_row(int r)
{
    CSDR_CK(cudaMemsetAsync(d_w_ + (size_t)r * row_len_, 0, (size_t)width_ * kRsPeriods * sizeof(float), st_));
    return CUTESDR_OK;
}

int ResamplerBank::run(int n_in, double rate, float* d_out, int out_stride, int out_off, const int* d_row_map,
                       int* n_out, int16_t* d_out16, double gain, int interleave16)
{
    if (n_in < 0 || n_in > max_in_ || !(rate > 0)) { set_error("resampler: bad length %d / rate %g", n_in, rate); return CUTESDR_E_ARG; }
    const double t0 = clk_.now();
    clk_.advance(n_in, rate, times_);
    const int m = (int)times_.size();
    if (n_out) *n_out = m;
    if (m > 0) {
        if (m > times_cap_) {
            CSDR_CK(cudaStreamSynchronize(st_));
            cudaFree(d_times_);
            cudaFree(d_wts_);
            cudaFree(d_pos_);
            times_cap_ = std::max(m, 2 * times_cap_) + 64;
            CSDR_CK(cudaMalloc(&d_times_, times_cap_ * sizeof(double)));
            CSDR_CK(cudaMalloc(&d_wts_, (size_t)times_cap_ * kRsPeriods * sizeof(float)));
            CSDR_CK(cudaMalloc(&d_pos_, times_cap_ * sizeof(int)));
        }
        k_resample_times<<<1, 32, 0, st_>>>(t0, rate, m, d_times_);
        lc_->n++;
        if (d_out || d_out16) {
            dim3 grid((m + 127) / 128, nrows_);
            k_resample_weights<<<(m + 3) / 4, 128, 0, st_>>>(d_times_, m, d_sinc_, d_wts_, d_pos_);
            if (width_ == 2)
                k_resample<2><<<grid, 128, 0, st_>>>(d_w_, row_len_, nrows_, d_wts_, d_pos_, m, d_out, out_stride, out_off,
                                                     d_row_map, d_out16, (float)gain, interleave16);
            else
                k_resample<1><<<grid, 128, 0, st_>>>(d_w_, row_len_, nrows_, d_wts_, d_pos_, m, d_out, out_stride, out_off,
                                                     d_row_map, d_out16, (float)gain, interleave16);
            lc_->n += 2;
            CSDR_CK(cudaGetLastError());
        }
    }
    k_resample_carry<<<nrows_, 64, 0, st_>>>(d_w_, row_len_, nrows_, n_in, width_);
    lc_->n++;
    CSDR_CK(cudaGetLastError());
    return CUTESDR_OK;
}

}  // namespace csdr
