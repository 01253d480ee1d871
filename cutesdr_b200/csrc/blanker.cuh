// blanker.cuh -- kernel 7: CNoiseProc impulse blanker on the shared wideband stream
// (dsp/noiseproc.cpp:59-176), restated as exact scans so a whole block runs in parallel.
#pragma once
#include "common.cuh"

namespace csdr {

class Blanker {
public:
    Blanker() {}
    ~Blanker();
    Blanker(const Blanker&) = delete;
    Blanker& operator=(const Blanker&) = delete;

    int init(int max_block, cudaStream_t st, LaunchCounter* lc);
    // CNoiseProc::SetupBlanker
    int setup(bool on, double threshold, double width_us, double sample_rate);
    bool on() const { return on_; }
    // ProcessBlanker over n samples: d_in (complex64, device) -> d_out (may not alias d_in)
    int run(const float2* d_in, float2* d_out, int n);

private:
    int max_block_ = 0;
    cudaStream_t st_ = 0;
    LaunchCounter* lc_ = nullptr;
    bool on_ = false, configured_ = false;
    double threshold_ = 0, width_ = 0, rate_ = 0, ratio_ = 0;
    int width_samples_ = 1, mag_samples_ = 0, delay_samples_ = 0;
    long long pos_ = 0;             // absolute stream position of the next sample
    // history carried between calls
    float* d_mag_ = nullptr;        // [hist_mag | max_block]
    float2* d_xh_ = nullptr;        // [hist_x | max_block] raw input with delay history
    int hist_mag_ = 0, hist_x_ = 0;
    double* d_scan_ = nullptr;      // prefix of (mag - delayed mag)
    long long* d_last_ = nullptr;   // prefix max of trigger positions
    double* d_part_ = nullptr;
    long long* d_partl_ = nullptr;
    double* d_carry_ = nullptr;     // [0] = moving sum at block end
    long long* d_carryl_ = nullptr; // [0] = last trigger position (relative to next block start)
};

}  // namespace csdr
