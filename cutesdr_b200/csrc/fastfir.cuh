// fastfir.cuh -- batched CFastFIR (reference: dsp/fastfir.cpp:55-321).
#pragma once
#include "common.cuh"

namespace csdr {

constexpr int kFirTapRow = 1028;      // kFirTaps time-domain taps per filter row, padded to a multiple of 4

class FirBank {
public:
    FirBank() {}
    ~FirBank();
    FirBank(const FirBank&) = delete;
    FirBank& operator=(const FirBank&) = delete;

    // nch channels in use out of `stride` slots; set_nch changes the number in use later (slots are all allocated)
    int init(int nch, int stride, cudaStream_t st, LaunchCounter* lc);
    void set_nch(int n) { nch_ = n; }
    // CFastFIR::SetupParameters for local channel i. Host bookkeeping only: the design itself (windowed sinc +
    // 2048-point FFT, dsp/fastfir.cpp:207-254) runs ON THE DEVICE, queued in stream order by the next run() --
    // a retune never synchronises the stream and never stalls the other channels.
    int setup(int i, double lo, double hi, double offset, double rate);
    // forget channel i's filter (its row is released when no other channel uses it)
    void release(int i);
    // nb overlap-save bursts starting at burst index first_burst: burst b filters ring samples
    // [b*1024-1024, b*1024+1024) and emits 1024 outputs into d_y[c*y_stride + k*1024 + t].
    // Burst 0 of a stream is computed in direct form (k_fir_first), see fastfir.cu.
    int run(const float2* d_ring, long long first_burst, int nb, float2* d_y, int y_stride);
    int num_filters() const { return rows_in_use_; }      // distinct responses in use (row 0 = all zeros)
    int capacity() const { return cap_; }

private:
    int flush();
    struct Params { double lo, hi, offset, rate; };
    struct Key {
        double lo, hi, rate;
        bool operator<(const Key& o) const {
            if (lo != o.lo) return lo < o.lo;
            if (hi != o.hi) return hi < o.hi;
            return rate < o.rate;
        }
    };
    struct Job { double lo, hi, rate; int row, pad; };
    int nch_ = 0, stride_ = 0, cap_ = 0, rows_in_use_ = 1;
    cudaStream_t st_ = 0;
    LaunchCounter* lc_ = nullptr;
    std::vector<Params> cur_;
    std::vector<int> h_id_;
    // filter rows are shared by channels with equal (lo, hi, rate), reference-counted and recycled: the table is
    // bounded by nch + 1 rows however many retunes a long-running bank sees
    std::map<Key, int> ids_;
    std::vector<Key> row_key_;
    std::vector<int> refs_, free_;
    std::vector<Job> jobs_;
    bool ids_dirty_ = true;
    float2* d_H_ = nullptr;       // [cap][2048] frequency responses
    float2* d_h_ = nullptr;       // [cap][kFirTapRow] the same filters' time-domain taps (start-up burst)
    Job* d_jobs_ = nullptr;
    int* d_id_ = nullptr;
    float2* d_tw_ = nullptr;
    PinnedStage stage_;
};

}  // namespace csdr
