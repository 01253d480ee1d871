// post.cuh -- kernel 4: the per-channel recurrent stages that follow CFastFIR:
// CSMeter -> CAgc -> CAmDemod | CSamDemod | CFmDemod | CSsbDemod, on 1024-sample bursts.
// References: dsp/smeter.cpp:62-93, dsp/agc.cpp:104-296, dsp/amdemod.cpp:50-82,
// dsp/samdemod.cpp:54-110, dsp/fmdemod.cpp:62-192, dsp/ssbdemod.cpp:48-53, dsp/fir.cpp, dsp/iir.cpp.
#pragma once
#include "common.cuh"

namespace csdr {

// Kaiser-window FIR designs of CFir (dsp/fir.cpp:173-367); return the tap count.
int design_kaiser_lp(double scale, double astop, double fpass, double fstop, double fs, double* coef);
int design_kaiser_hp(double scale, double astop, double fpass, double fstop, double fs, double* coef);

constexpr int kYHist = 2048;          // >= MAX_DELAY_BUF-1 samples of AGC signal delay

enum PostMode { POST_NONE = -1, POST_AM = 0, POST_SAM = 1, POST_FM = 2, POST_SSB = 3, POST_AGC_ONLY = 100 };

// values shared by every channel of a group (they depend on the sample rate only)
struct PostUniform {
    double sm_attack, sm_decay;                 // CSMeter alphas
    int agc_delay, agc_window;                  // CAgc m_DelaySamples / m_WindowSamples
    double sam_alpha, sam_beta, sam_lo, sam_hi; // CSamDemod PLL
    double fm_alpha, fm_beta, fm_lo, fm_hi, fm_gain, fm_dc_alpha, fm_sq_alpha;
    double lp_b0, lp_b1, lp_b2, lp_a1, lp_a2;   // 3 kHz Q=1 biquad of CFmDemod
    int stereo;                                 // 1: interleaved (L,R) output, CDemodulator::ProcessData(.., TYPECPX*)
    int sam_ntaps;                              // Hilbert pair of the stereo SAM demodulator
};

class PostBank {
public:
    PostBank() {}
    ~PostBank();
    PostBank(const PostBank&) = delete;
    PostBank& operator=(const PostBank&) = delete;

    int init(int nch, int stride, double rate, int max_samples, cudaStream_t st, LaunchCounter* lc);
    // new demodulator object for local channel i (zeroes its demod state); mode is a
    // CUTESDR_DEMOD_* value or POST_AGC_ONLY
    void set_mode(int i, int mode);
    // slots in use (<= stride); free_channel parks a slot (no demodulator, all state re-initialised at its next use)
    void set_nch(int n) { nch_ = n; }
    void free_channel(int i);
    // CAgc::SetParameters (rate is the group's)
    void set_agc(int i, int on, int hang, int thresh, int manual_gain, int slope, int decay);
    // CAmDemod::SetBandwidth -- re-designs AND zeroes the post filter (dsp/amdemod.cpp:56-60)
    void set_am_bandwidth(int i, double bw);
    // CFmDemod::SetSquelch and the FmBW handed to ProcessData (m_DemodInfo.HiCut)
    void set_fm(int i, int squelch_value, double fm_bw);

    // The FIR (or the caller) writes the burst's complex64 samples for channel c at
    //   y_in() + c * y_stride()   (contiguous in time).
    float2* y_in() { return d_y_ + kYHist; }
    int y_stride() const { return y_row_; }
    // Run n samples (a burst is 1024) already placed at y_in(). d_audio (float32) is indexed
    // [chan_map[c]][audio_stride] + audio_off. tap3() then holds the post-AGC complex stream
    // [c][tap3_stride()].
    int run(int n, float* d_audio, int audio_stride, int audio_off, const int* d_chan_map);
    const float2* tap3() const { return d_z_; }
    int tap3_stride() const { return max_n_; }
    // stereo output (dsp/demodulator.cpp:221-273): audio rows then hold interleaved L,R float pairs
    void set_stereo(bool on) { uni_.stereo = on ? 1 : 0; }
    bool stereo() const { return uni_.stereo != 0; }
    // S-meter readout for local channel i (synchronises the stream)
    int read_smeter(int i, double* peak, double* ave);
    double rate() const { return rate_; }

private:
    int upload();
    int nch_ = 0, stride_ = 0, max_n_ = 0;
    double rate_ = 0;
    cudaStream_t st_ = 0;
    LaunchCounter* lc_ = nullptr;
    PostUniform uni_{};
    struct AgcHost { int on = 1, hang = 0, thresh = 0, mgain = 0, decay = 0; double slope = 0; bool valid = false; };
    std::vector<AgcHost> agc_;
    std::vector<double> fm_bw_;
    std::vector<double> h_par_;     // [P_COUNT][stride]
    std::vector<double> h_taps_;    // [stride][kFirMax]
    std::vector<int> h_mode_, h_reset_;
    bool dirty_ = true, taps_dirty_ = true, need_reset_kernel_ = false;
    PinnedStage stage_;
    double* d_par_ = nullptr;
    double* d_taps_ = nullptr;
    int* d_mode_ = nullptr;
    int* d_reset_ = nullptr;
    double* d_state_ = nullptr;     // [S_COUNT][stride]
    int* d_istate_ = nullptr;       // [I_COUNT][stride]
    // channel-major work rows (row c = channel c)
    float2* d_y_ = nullptr;         // [nch][kYHist + max_n]   FIR output with the AGC delay history in front
    double* d_magh_ = nullptr;      // [nch][kAgcBuf]          last window-1 AGC log-magnitudes
    double* d_smag_ = nullptr;      // [nch][max_n]            S-meter dB values
    double* d_peak_ = nullptr;      // [nch][max_n]            sliding-window peak, then max(attack,decay)
    float2* d_z_ = nullptr;         // [nch][max_n]            AGC output (PROFILE_3 tap)
    double* d_u_ = nullptr;         // [nch][max_n]            envelope (AM/SAM)
    double* d_th_ = nullptr;        // [nch][max_n]            phase angle (SAM/FM)
    double* d_v_ = nullptr;         // [nch][kHist + max_n]    FIR input with history (AM post filter / FM squelch)
    double* d_v2_ = nullptr;        // [nch][kHist + max_n]    second FIR input row (stereo SAM imaginary part)
    double* d_sam_taps_ = nullptr;  // [2][kFirMax] I and Q coefficient sets of the stereo SAM Hilbert pair
    double* d_qpow_ = nullptr;      // [max_n+1] powers of (1 - squelch alpha)
    int y_row_ = 0, v_row_ = 0;
};

}  // namespace csdr
