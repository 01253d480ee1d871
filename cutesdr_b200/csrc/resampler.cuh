// resampler.cuh -- kernel 5: batched CFractResampler (dsp/fractresampler.cpp:50-352), a 28-tap
// Blackman-Harris windowed-sinc interpolator read from a 280001-entry table (10000 points per
// zero crossing, index truncated -- no interpolation between table entries).
#pragma once
#include "common.cuh"

namespace csdr {

constexpr int kRsPeriods = 28;                 // SINC_PERIODS
constexpr int kRsPts = 10000;                  // SINC_PERIOD_PTS
constexpr int kRsLen = kRsPeriods * kRsPts + 1;

// Output time sequence of one Resample() call, computed exactly like the reference's double
// accumulator (t += Rate per output, t -= InLength at the end). Shared by all channels of a bank
// because they share Rate and start together.
class ResampleClock {
public:
    void reset() { t_ = 0.0; }
    double now() const { return t_; }
    // fills times with the fractional input time of every output produced for n_in inputs
    void advance(int n_in, double rate, std::vector<double>& times);
private:
    double t_ = 0.0;
};

class ResamplerBank {
public:
    ResamplerBank() {}
    ~ResamplerBank();
    ResamplerBank(const ResamplerBank&) = delete;
    ResamplerBank& operator=(const ResamplerBank&) = delete;

    // nrows independent streams, at most max_in inputs per call. width 1: real rows (a complex stream may be two
    // rows); width 2: rows of interleaved (re,im) / (left,right) float pairs -- the stereo form
    // Resample(.., TYPECPX*, TYPESTEREO16*/TYPECPX*), dsp/fractresampler.cpp:194-249,309-352: same clock, same
    // weights, two accumulators.
    int init(int nrows, int max_in, cudaStream_t st, LaunchCounter* lc, int width = 1);
    // rows in use (<= the nrows given to init); reset_row zeroes one row's 28 carried inputs in stream order
    void set_rows(int n) { nrows_ = n; }
    int reset_row(int r);
    // device buffer the producer writes new inputs into: row r starts at in_ptr() + r*in_stride()
    float* in_ptr() { return d_w_ + width_ * kRsPeriods; }
    int in_stride() const { return row_len_; }
    int max_out(int n_in, double rate) const { return (int)(n_in / rate) + 8; }
    // Resample n_in new inputs per row. Output k of row r goes to
    //   d_out[row_map ? row_map[r] : r][out_stride] + width * (out_off + k)   (float32, `width` floats), or, when d_out16 is
    // given, to int16 after gain + clip + truncation (dsp/fractresampler.cpp:228-239).
    // Returns the number of outputs per row through *n_out.
    int run(int n_in, double rate, float* d_out, int out_stride, int out_off, const int* d_row_map, int* n_out,
            int16_t* d_out16 = nullptr, double gain = 1.0, int interleave16 = 1);

private:
    int nrows_ = 0, max_in_ = 0, row_len_ = 0, width_ = 1;
    cudaStream_t st_ = 0;
    LaunchCounter* lc_ = nullptr;
    ResampleClock clk_;
    float* d_w_ = nullptr;        // [nrows][28 + max_in] : 28 carried inputs, then the new ones
    float* d_sinc_ = nullptr;     // [kRsLen]
    double* d_times_ = nullptr;
    float* d_wts_ = nullptr;      // [outputs][28] weights shared by all rows
    int* d_pos_ = nullptr;        // [outputs] integer input position
    int times_cap_ = 0;
    std::vector<double> times_;
};

}  // namespace csdr
