// decimator.cuh -- batched CDownConvert: NCO mix + CIC3/half-band decimate-by-2 cascade for a
// group of channels that share one stage list (reference: dsp/downconvert.cpp:98-460).
#pragma once
#include "common.cuh"

namespace csdr {

// Stage ladder of CDownConvert::SetDataRate (dsp/downconvert.cpp:114-173).
// lens: 3 = CIC3, 11..51 = half-band tap count. Returns the output rate.
double plan_stages(double in_rate, double max_bw, std::vector<int>& lens);

// Amplitude of the reference's quadrature oscillator at stream sample n (its gain servo
// 1.95-|z|^2 settles |z|^2 at 0.95 from 1.0, dsp/downconvert.cpp:211-216), divided by the
// steady-state sqrt(0.95) the kernels fold into their output scale. Applied to the first
// kNcoStartup samples of a stream, in place, on the device copy of the wideband block.
constexpr int kNcoStartup = 512;
int apply_nco_startup_gain(float2* d_x, long long stream_pos, int n, cudaStream_t st, LaunchCounter* lc);
// raw radio samples (fmt 1/2, see Decimator::run_block) -> complex64
int unpack_samples(const void* d_raw, int fmt, float2* d_out, int n, cudaStream_t st, LaunchCounter* lc);
inline int sample_bytes(int fmt) { return fmt == 0 ? 8 : (fmt == 1 ? 4 : 6); }

// Tuning / ablation knobs, read from the environment ONCE per Decimator::init (never on the launch path).
// None is needed for correctness; DESIGN.md section 7 lists them.
struct Tuning {
    bool no_tc = false, no_hbchain = false, hbtail = false, no_overlap = false, debug_timing = false;
    bool no_hbstream = false, no_hbtail = false;
    int fuse_hb = -1, tile = 0, tc_seg = 0, hs_halo = 300, hs_ctas = 3, tc_spare = 0;
    static Tuning from_env();
};

struct NcoDev {
    unsigned long long inc;   // phase increment per input sample, turns * 2^64
    float w1c, w1s;           // e^{j inc}
    float wgc, wgs;           // e^{j G inc}, G = 2^ncic
    float wtc, wts;           // e^{j 16 inc} (kernel 1T's oscillator step at fs/16)
};

class Decimator {
public:
    Decimator() {}
    ~Decimator();
    Decimator(const Decimator&) = delete;
    Decimator& operator=(const Decimator&) = delete;

    // nch channels at in_rate with the stage list for max_bw; block_len wideband samples per
    // DSP block (multiple of 2^stages, dsp/downconvert.cpp:182-183).
    int init(int nch, double in_rate, double max_bw, int block_len, cudaStream_t st, LaunchCounter* lc);

    // CDownConvert::SetFrequency for local channel i (freq already includes the CW offset)
    void set_frequency(int i, double nco_freq);
    // A new CDownConvert chain for local channel i only (CDownConvert::SetDataRate rebuilds the stage objects,
    // dsp/downconvert.cpp:114-173): its stage histories, decimated ring and oscillator phase restart from zero, queued in
    // stream order; every other channel's state is untouched.
    int reset_channel(int i);

    // Run one block of L <= block_len samples (L a multiple of 2^stages; L < 0 = the full block).
    // d_x: the block (device, complex64, 16-byte aligned). halo_cur: kHaloMax samples that preceded
    // it in the stream (zeros at stream start); halo_next (a different buffer) receives the last
    // kHaloMax samples of [halo_cur | block] for the next call -- kernel 1 writes it itself, so the
    // caller's block is read in place and no copy node sits between consecutive launches.
    // fmt: 0 = complex64, 1 = int16 I,Q pairs, 2 = packed int24 I,Q (unpacked inside kernel 1's tile load)
    int run_block(const void* d_x, const float2* halo_cur, float2* halo_next, int L = -1, int fmt = 0);
    int block_len() const { return block_len_; }
    // Overlap mode (used by the bank): the HBM-bound half-band stages (kernel 2) run on an internal
    // second stream so they execute under the FP32-bound kernel 1 of the NEXT block. The first stage
    // ring then holds two blocks. done_event() is recorded after the last kernel that writes the
    // decimated ring; wait_before_output(ev) orders that last writer after `ev` (ring-reuse guard).
    // Without overlap everything is ordered on the caller's stream.
    int set_overlap(bool on);
    cudaEvent_t done_event() const { return ev_done_; }
    int wait_before_output(cudaEvent_t ev);
    int join_main();          // order the caller's stream after all internal work queued so far
    // CUDA-event timing of kernel 1 on the launching stream (bench.py's roofline line)
    void enable_timing(bool on) { timing_ = on; }
    int read_timing(double* ms_total, long long* launches);
    // kernel 1T in use? and the real multiply-adds (MAC = 2 flop) of its GEMM per full block, counted ONCE per product
    // (fp32-equivalent: the three tf32 partial products that emulate one fp32 product count as one)
    bool tensor_path() const { return tc_; }
    bool tensor_f16() const { return tc_ && tc_f16_; }        // the last block ran the fp16 form (int16 wire samples, kind::f16)
    double tensor_flops_per_block() const { return tc_ ? 2.0 * (128.0 * tc_groups_) * (block_len_ / 16.0) * 192.0 : 0.0; }

    int nch() const { return nch_; }
    int stride() const { return stride_; }
    double out_rate() const { return out_rate_; }
    int out_per_block() const { return n_out_; }
    long long total_out() const { return total_out_; }   // decimated samples produced so far
    const std::vector<int>& stages() const { return lens_; }
    // per-channel ring [c][kDecRing] of decimated samples; sample j lives at j & (kDecRing-1)
    const float2* ring() const { return d_ring_; }

private:
    int upload_dirty();
    int nch_ = 0, stride_ = 0, block_len_ = 0, ncic_ = 0, nhbf_ = 0, n_out_ = 0, tile_len_ = 0;
    int k1_stages() const { return ncic_ + nhbf_; }     // stages fused into kernel 1
    double in_rate_ = 0, out_rate_ = 0;
    std::vector<int> lens_;
    cudaStream_t st_ = 0;
    LaunchCounter* lc_ = nullptr;
    std::vector<NcoDev> h_nco_;
    PinnedStage stage_;
    bool dirty_ = true;
    // kernel 1T (tensor-core form of kernel 1, used when the ladder starts with >= 4 CIC3 stages)
    bool tc_ = false, tc_dirty_ = true, tc_f16_ = false;
    int tc_seg_len_ = 0, tc_groups_ = 0;
    float* d_tc_coef_ = nullptr;
    float* d_tc_coef16_ = nullptr;     // fp16 hi/lo form of the kernel-1T coefficient table (int16 blocks)
    NcoDev* d_nco_ = nullptr;
    unsigned long long* d_phase_[2] = {nullptr, nullptr};
    int phase_cur_ = 0;
    long long total_out_ = 0;
    std::vector<long long> stage_base_;   // absolute row index of the next row each stage ring receives
    // stage rings: ring s holds the INPUT rows of half-band stage s (time-major [row][stride])
    std::vector<float2*> d_stage_;
    std::vector<int> stage_rows_;      // power of two
    // kernel 2s (streaming fused half-bands): TMA descriptor of stage ring 0 and the number of stages it fuses
    alignas(64) unsigned char tmap0_[128];
    int hs_stages_ = 0;
    float2* d_ring_ = nullptr;
    bool overlap_ = false;
    Tuning tun_;
    cudaStream_t st_hb_ = 0;
    cudaEvent_t ev_k1_ = nullptr, ev_done_ = nullptr;
    cudaEvent_t ev_k2_[4] = {nullptr, nullptr, nullptr, nullptr};
    long long blocks_run_ = 0;
    bool timing_ = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_pool_;
    size_t ev_used_ = 0;
    double k1_ms_ = 0.0;
    long long k1_n_ = 0;
};

}  // namespace csdr
