// bank.cuh -- the receiver bank: N x CDemodulator (dsp/demodulator.cpp:47-215) on one wideband
// stream. Channels that share a decimation chain (same MaxBW -> same stage list and output
// rate) form a Group whose kernels run batched; groups run back to back on the bank's stream.
#pragma once
#include "common.cuh"
#include "decimator.cuh"
#include "fastfir.cuh"
#include "post.cuh"
#include "resampler.cuh"
#include "blanker.cuh"

namespace csdr {

struct ChanCfg {
    int mode = -1;                       // m_DemodMode
    cutesdr_demod_info info{};           // m_DemodInfo
    bool configured = false;
    double demod_cw = 0.0;               // CDemodulator::m_CW_Offset
    double dc_cw = 0.0;                  // CDownConvert::m_CW_Offset
    double dc_nco_freq = 0.0;            // CDownConvert::m_NcoFreq (already includes a CW offset)
    double dc_max_bw = 48000.0;          // CDownConvert::m_MaxBW after SetInputSampleRate
    double max_bw = 48000.0;             // m_DesiredMaxOutputBandwidth
    int group = -1, local = -1;
    unsigned tap_mask = 0;
    std::vector<float> tap[5];
};

// input slots of the pipelined entry points: the transfer of block k+3 (host -> GPU, and the NCCL broadcast on several
// GPUs) may run while block k is still being processed -- with two slots a transfer that takes longer than one block's
// kernels (8 ranks: 2 x ~0.1 ms of broadcast latency per block) stalled the pipeline
constexpr int kAsyncSlots = 4;

struct Group {
    double max_bw = 0;
    std::vector<int> chans;              // user channel index of each local slot in use (-1 = parked slot)
    int cap = 0;                         // slots allocated (the stage objects' stride)
    std::vector<int> h_chan_map;         // host copy of d_chan_map
    csdr::PinnedStage stage;
    Decimator dec;
    FirBank fir;
    PostBank post;
    std::unique_ptr<ResamplerBank> rs;
    float* d_demod = nullptr;            // demod output [local c][kMaxBurstSamples] when resampling
    int* d_chan_map = nullptr;
    int* d_local_map = nullptr;          // identity map (for resampler input rows)
    long long bursts_done = 0;
    int last_fir_n = 0;                  // FIR samples produced by the last block
    // The burst stages (FIR, AGC/demod, resampler) run on their own stream so the small sequential
    // kernels overlap the next blocks' decimation instead of idling the GPU.
    cudaStream_t st_post = 0;
    cudaEvent_t ev_dec = nullptr;        // decimator output of the current block is in the ring
    cudaEvent_t ev_post[8] = {};         // completion of recent burst chains (round robin)
    long long post_launches = 0;
    struct Pending { cudaEvent_t ev; long long block; };
    std::vector<Pending> pending;        // burst chains the main stream has not been ordered after yet
    bool any_tap = false;
    ~Group();
};

constexpr int kMaxBurstSamples = 2 * kBurst;   // a DSP block completes at most two FIR bursts

// Test-bench spectrum of one PROFILE tap of one channel: CTestBench's frequency-domain DisplayData branch
// (gui/testbench.cpp:583-611) kept on the device -- the tap's samples are appended to a 2048-sample frame in stream
// order and every m_DisplaySkipValue-th full frame goes through the attached CFft object's PutInDisplayFFT.
constexpr int kTestFftSize = 2048;             // TEST_FFTSIZE, gui/testbench.h:44
struct TapSpectrum {
    int ch = -1, profile = 0;
    cutesdr_fft* fft = nullptr;
    int display_rate = 10;                     // m_DisplayRate
    double rate = 0;                           // m_DisplaySampleRate
    int skip_value = 0, skip_counter = -2;     // m_DisplaySkipValue / m_DisplaySkipCounter (:570,574)
    int pos = 0;                               // m_FftBufPos
    long long frames = 0;                      // PutInDisplayFFT calls so far
    float2* d_frame = nullptr;                 // m_FftInBuf
    cudaEvent_t ev[8] = {};
    long long n_ev = 0;
    ~TapSpectrum();
};

}  // namespace csdr

struct cutesdr_bank {
    int nch = 0;
    double in_rate = 0;
    int device = 0;
    cudaStream_t st = 0;
    csdr::LaunchCounter lc;
    std::mutex mu;
    std::vector<csdr::ChanCfg> ch;
    std::vector<std::unique_ptr<csdr::Group>> groups;
    bool layout_dirty = true;
    int L = 0;                           // m_InBufLimit
    float2* d_x = nullptr;               // [L] staging of host / blanked blocks
    // pipelined host interface (process_async): two H2D staging slots on a copy stream, D2H on another
    float2* d_xs[csdr::kAsyncSlots] = {};
    cudaStream_t st_h2d = 0, st_d2h = 0;
    cudaEvent_t ev_h2d[csdr::kAsyncSlots] = {}, ev_free[csdr::kAsyncSlots] = {}, ev_host[csdr::kAsyncSlots] = {}, ev_d2h = nullptr;
    long long async_blocks = 0;
    bool d2h_pending = false;
    float2* d_halo[2] = {nullptr, nullptr};   // [kHaloMax] tail of the previous block (double buffer)
    int halo_cur = 0;
    const void* last_block = nullptr;    // device block of the most recent DSP block (after the blanker)
    int last_fmt = 0;
    float2* h_stage = nullptr;           // pinned staging of one block
    int h_fill = 0;
    int h_fmt = 0;                       // sample format of the partially filled staging block
    long long stream_pos = 0;
    long long block_index = 0;
    float* d_audio = nullptr;            // [nch][audio_cap]
    int audio_cap = 0;
    double audio_rate = 0.0;             // > 0: CFractResampler to this rate
    bool stereo = false;                 // CDemodulator::ProcessData(.., TYPECPX*) : interleaved L,R output
    std::unique_ptr<csdr::Blanker> nb;
    bool nb_on = false;
    double nb_thresh = 50.0, nb_width = 2.0;
    std::vector<int> blk_nout;           // per channel, samples produced by the last block
    // UDP packet front end (CUdpThread::OnreadyRead, interface/netiobase.cpp:464-534)
    unsigned short pkt_last_seq = 0;     // m_LastSeqNum
    long long missed_packets = 0;        // m_MissedPackets
    std::vector<unsigned char> pkt_payload;
    std::vector<std::unique_ptr<csdr::TapSpectrum>> tap_spectra;

    int rebuild();
    int create_group(double max_bw, const std::vector<int>& chans, int min_cap, int* index);
    int move_channel(int c);             // channel c's decimation chain changed on a running bank
    int run_block(const void* d_block, int fmt, float* d_audio_out, int audio_stride, const int* audio_off, int* n_out_max);
    int collect_taps();
    int feed_tap_spectra(csdr::Group& g, int gi, int n_burst, float* d_audio_out, int audio_stride, int audio_off);
    int join();                          // order the main stream after every outstanding burst chain
    int sync_all();
    ~cutesdr_bank();
};
