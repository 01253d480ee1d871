// fastfir.cuh -- batched CFastFIR (reference: dsp/fastfir.cpp:55-321).
#pragma once
#include "common.cuh"

namespace csdr {

class FirBank {
public:
    FirBank() {}
    ~FirBank();
    FirBank(const FirBank&) = delete;
    FirBank& operator=(const FirBank&) = delete;

    int init(int nch, int stride, cudaStream_t st, LaunchCounter* lc);
    // CFastFIR::SetupParameters for local channel i
    int setup(int i, double lo, double hi, double offset, double rate);
    // nb overlap-save bursts starting at burst index first_burst: burst b filters ring samples
    // [b*1024-1024, b*1024+1024) and emits 1024 outputs into d_y[c*y_stride + k*1024 + t].
    int run(const float2* d_ring, long long first_burst, int nb, float2* d_y, int y_stride);
    int num_filters() const { return nfilt_; }

private:
    int upload();
    struct Params { double lo, hi, offset, rate; };
    struct Key {
        double lo, hi, rate;
        bool operator<(const Key& o) const {
            if (lo != o.lo) return lo < o.lo;
            if (hi != o.hi) return hi < o.hi;
            return rate < o.rate;
        }
    };
    int nch_ = 0, stride_ = 0, nfilt_ = 0, cap_ = 0;
    cudaStream_t st_ = 0;
    LaunchCounter* lc_ = nullptr;
    std::vector<Params> cur_;
    std::vector<int> h_id_;
    std::vector<float2> h_H_;
    std::map<Key, int> ids_;
    bool dirty_ = true;
    float2* d_H_ = nullptr;
    int* d_id_ = nullptr;
    float2* d_tw_ = nullptr;
};

}  // namespace csdr
