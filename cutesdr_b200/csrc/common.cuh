// common.cuh -- shared declarations for libcutesdr_cuda (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <math.h>
#include <mutex>
#include <string>
#include <vector>
#include <map>
#include <memory>

#include "../../include/cutesdr_cuda.h"

namespace csdr {

void set_error(const char* fmt, ...);

#define CSDR_CK(call)                                                                        \
    do {                                                                                     \
        cudaError_t e__ = (call);                                                            \
        if (e__ != cudaSuccess) {                                                            \
            csdr::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return CUTESDR_E_CUDA;                                                           \
        }                                                                                    \
    } while (0)

#define CSDR_TRY(expr)                    \
    do {                                  \
        int rc__ = (expr);                \
        if (rc__ < 0) return rc__;        \
    } while (0)

constexpr double kTwoPi = 2.0 * 3.14159265358979323846;   // K_2PI, dsp/datatypes.h:42
constexpr double kPi = 3.14159265358979323846;

constexpr int kHaloMax = 2048;      // complex samples kept in front of every wideband block (kernel-1 halo)
constexpr int kFirFft = 2048;       // CONV_FFT_SIZE, dsp/fastfir.cpp:55
constexpr int kFirTaps = 1025;      // CONV_FIR_SIZE, dsp/fastfir.cpp:56
constexpr int kBurst = 1024;        // samples CFastFIR emits per FFT
constexpr int kDecRing = 4096;      // per-channel ring of decimated samples feeding CFastFIR
constexpr int kMaxStages = 24;
constexpr int kAgcBuf = 2048;       // MAX_DELAY_BUF, dsp/agc.h:15
constexpr int kFirMax = 75;         // MAX_NUMCOEF, dsp/fir.h:15

// Every kernel launch of the library goes through this counter (reported as gpu_launches).
struct LaunchCounter {
    long long n = 0;
};

// Small parameter updates travel host -> device through a ring of PINNED buffers: the host fills a slot and queues
// an async copy; nothing ever synchronises a stream (a cudaMemcpyAsync from pageable memory first waits for the
// stream's earlier work). A slot is reused only after its previous copy has completed (host-side event wait, which
// in practice never blocks: the ring is deeper than the run-ahead of the pipeline).
class PinnedStage {
public:
    PinnedStage() {}
    ~PinnedStage()
    {
        for (auto& s : slots_) { if (s.ev) cudaEventDestroy(s.ev); if (s.p) cudaFreeHost(s.p); }
    }
    PinnedStage(const PinnedStage&) = delete;
    PinnedStage& operator=(const PinnedStage&) = delete;
    // returns a pinned buffer of at least `bytes`; the caller fills it and hands it to push()
    int acquire(size_t bytes, void** out)
    {
        if (slots_.empty()) slots_.resize(kDepth);
        Slot& s = slots_[cur_];
        if (s.ev && s.used) { CSDR_CK(cudaEventSynchronize(s.ev)); s.used = false; }
        if (bytes > s.cap) {
            if (s.p) cudaFreeHost(s.p);
            s.p = nullptr; s.cap = 0;
            size_t cap = 4096;
            while (cap < bytes) cap <<= 1;
            CSDR_CK(cudaHostAlloc(&s.p, cap, cudaHostAllocDefault));
            s.cap = cap;
        }
        if (!s.ev) CSDR_CK(cudaEventCreateWithFlags(&s.ev, cudaEventDisableTiming));
        *out = s.p;
        return CUTESDR_OK;
    }
    // async copy of the acquired buffer to dst on st
    int push(void* dst, size_t bytes, cudaStream_t st)
    {
        Slot& s = slots_[cur_];
        CSDR_CK(cudaMemcpyAsync(dst, s.p, bytes, cudaMemcpyHostToDevice, st));
        CSDR_CK(cudaEventRecord(s.ev, st));
        s.used = true;
        cur_ = (cur_ + 1) % kDepth;
        return CUTESDR_OK;
    }
    int upload(void* dst, const void* src, size_t bytes, cudaStream_t st)
    {
        void* p = nullptr;
        CSDR_TRY(acquire(bytes, &p));
        memcpy(p, src, bytes);
        return push(dst, bytes, st);
    }
private:
    static constexpr int kDepth = 8;
    struct Slot { void* p = nullptr; size_t cap = 0; cudaEvent_t ev = nullptr; bool used = false; };
    std::vector<Slot> slots_;
    int cur_ = 0;
};

inline int round_up(int v, int m) { return (v + m - 1) / m * m; }
inline int next_pow2(long long v) { int p = 1; while (p < v) p <<= 1; return p; }

}  // namespace csdr
