// fastfir.cu -- kernel 3: batched CFastFIR (1025-tap complex band-pass by overlap-save with a
// 2048-point FFT), one CTA per (channel, burst).
//
// Reference: CFastFIR::SetupParameters / ProcessData, dsp/fastfir.cpp:178-306. The reference
// transforms with Ooura's e^{+j} kernel and inverts with its conjugate (dsp/fft.cpp:416-426);
// circular convolution is the same for either sign pair, so this file uses the textbook pair.
//
// Layout: input windows come from the per-channel ring [c][kDecRing] the decimator fills;
// H lives as complex64 [n_filters][2048] (channels with equal (lo,hi,offset,rate) share a row);
// output goes channel-major [c][row_stride] (one contiguous 1024-sample run per burst and channel).
// Filters are designed on the device (k_fir_design) and the first burst of a stream is computed in direct form
// (k_fir_first).
// The FFT is a shared-memory Stockham autosort (five radix-4 passes + one radix-2, ping-pong buffers).
#include "fastfir.cuh"

namespace csdr {

// ------------------------------------------------------------------------------------------
// device
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 cmulf(float2 a, float2 b)
{
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// 2048 = 4^5 * 2: five radix-4 Stockham passes and one radix-2 pass (autosort: natural order in and
// out, ping-pong buffers). 256 threads; a radix-4 pass gives every thread two butterflies.
// tw[m] = e^{-2 pi i m / 2048}; CONJ selects the inverse kernel.
template <bool CONJ>
__device__ __forceinline__ float2 twiddle(const float2* __restrict__ tw, int m)
{
    // m in [0, 2048): the table holds the first half, the second half is its negation
    float2 w = tw[m & 1023];
    if (m & 1024) { w.x = -w.x; w.y = -w.y; }
    if (CONJ) w.y = -w.y;
    return w;
}

template <bool CONJ>
__device__ __forceinline__ void stockham_r4(const float2* __restrict__ src, float2* __restrict__ dst,
                                            const float2* __restrict__ tw, int ns)
{
    // butterflies j = 0..511; inputs j + r*512; k = j mod ns; outputs (j-k)*4 + k + r*ns
    const int tw_mul = 512 / ns;           // 2048 / (4 ns)
#pragma unroll
    for (int q = 0; q < 2; q++) {
        const int j = threadIdx.x + q * 256;
        const int k = j & (ns - 1);
        float2 a0 = src[j], a1 = src[j + 512], a2 = src[j + 1024], a3 = src[j + 1536];
        if (ns > 1) {
            a1 = cmulf(a1, twiddle<CONJ>(tw, k * tw_mul));
            a2 = cmulf(a2, twiddle<CONJ>(tw, 2 * k * tw_mul));
            a3 = cmulf(a3, twiddle<CONJ>(tw, 3 * k * tw_mul));
        }
        const float2 s02 = make_float2(a0.x + a2.x, a0.y + a2.y), d02 = make_float2(a0.x - a2.x, a0.y - a2.y);
        const float2 s13 = make_float2(a1.x + a3.x, a1.y + a3.y), d13 = make_float2(a1.x - a3.x, a1.y - a3.y);
        // forward: multiply d13 by -i; inverse: by +i
        const float2 r13 = CONJ ? make_float2(-d13.y, d13.x) : make_float2(d13.y, -d13.x);
        const int j0 = ((j - k) << 2) + k;
        dst[j0] = make_float2(s02.x + s13.x, s02.y + s13.y);
        dst[j0 + ns] = make_float2(d02.x + r13.x, d02.y + r13.y);
        dst[j0 + 2 * ns] = make_float2(s02.x - s13.x, s02.y - s13.y);
        dst[j0 + 3 * ns] = make_float2(d02.x - r13.x, d02.y - r13.y);
    }
}

template <bool CONJ>
__device__ __forceinline__ void stockham_r2_last(const float2* __restrict__ src, float2* __restrict__ dst,
                                                 const float2* __restrict__ tw)
{
    // final pass: ns = 1024, k = j, w = e^{-+2 pi i j / 2048}
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const int j = threadIdx.x + q * 256;
        const float2 a = src[j];
        const float2 b = cmulf(src[j + 1024], twiddle<CONJ>(tw, j));
        dst[j] = make_float2(a.x + b.x, a.y + b.y);
        dst[j + 1024] = make_float2(a.x - b.x, a.y - b.y);
    }
}

template <bool CONJ>
__device__ __forceinline__ float2* fft2048(float2* src, float2* dst, const float2* __restrict__ tw)
{
    for (int ns = 1; ns < 1024; ns <<= 2) {
        stockham_r4<CONJ>(src, dst, tw, ns);
        __syncthreads();
        float2* t = src; src = dst; dst = t;
    }
    stockham_r2_last<CONJ>(src, dst, tw);
    __syncthreads();
    return dst;
}

__global__ void __launch_bounds__(256) k_fastfir(const float2* __restrict__ ring, long long first_burst,
                                                 const float2* __restrict__ H, const int* __restrict__ filt_id,
                                                 const float2* __restrict__ tw_g, float2* __restrict__ y, int stride)
{
    __shared__ float2 bufA[kFirFft];
    __shared__ float2 bufB[kFirFft];
    __shared__ float2 tw[1024];
    const int c = blockIdx.x;
    const long long burst = first_burst + blockIdx.y;
    const long long w0 = burst * kBurst - kBurst;          // first sample of the 2048 window
    const float2* r = ring + (size_t)c * kDecRing;
    for (int i = threadIdx.x; i < 1024; i += 256) tw[i] = tw_g[i];
    for (int i = threadIdx.x; i < kFirFft; i += 256) {
        const long long j = w0 + i;
        // samples before the stream start are the reference's zero-initialised overlap buffer
        bufA[i] = j < 0 ? make_float2(0.f, 0.f) : r[(size_t)(j & (kDecRing - 1))];
    }
    __syncthreads();
    float2* X = fft2048<false>(bufA, bufB, tw);
    float2* other = (X == bufA) ? bufB : bufA;
    const float2* Hc = H + (size_t)filt_id[c] * kFirFft;
    for (int i = threadIdx.x; i < kFirFft; i += 256) X[i] = cmulf(Hc[i], X[i]);   // CpxMpy, dsp/fastfir.cpp:312-321
    __syncthreads();
    float2* Y = fft2048<true>(X, other, tw);
    // keep samples 1024..2047 (dsp/fastfir.cpp:291-294); channel-major rows: coalesced stores
    float2* yo = y + (size_t)c * stride + (size_t)blockIdx.y * kBurst;
    for (int i = threadIdx.x; i < kBurst; i += 256) yo[i] = Y[kBurst + i];
}

// ------------------------------------------------------------------------------------------
// Start-up burst. The very first overlap-save window of a stream is [1024 zeros | x[0..1023]]: output k is
// sum_{m<=k} h[m] x[k-m], i.e. the filter's leading tail (taps of 1e-7 relative size) applied to full-size
// samples. A float32 FFT convolution has an ABSOLUTE error floor of ~1e-7 x the window's peak, which swamps those
// outputs -- and the SAM/FM PLLs acquire on exactly these samples (atan2 only looks at the ratio im/re), so the
// whole acquisition transient would differ from the reference's double-precision one. This burst is therefore
// computed in direct form with double accumulation from the time-domain taps: every output is accurate RELATIVE
// TO ITS OWN size, and the demodulators can be compared with the reference from the first sample on.
// One CTA per channel; thread t owns outputs t, 511-t, 512+t, 1023-t (equal work per thread).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_fir_first(const float2* __restrict__ ring, const float2* __restrict__ taps,
                                                   const int* __restrict__ filt_id, float2* __restrict__ y, int stride)
{
    __shared__ float2 xs[kBurst];
    __shared__ float2 hs[kBurst];
    const int c = blockIdx.x;
    const float2* r = ring + (size_t)c * kDecRing;
    const float2* h = taps + (size_t)filt_id[c] * kFirTapRow;
    for (int i = threadIdx.x; i < kBurst; i += 256) { xs[i] = r[i]; hs[i] = h[i]; }
    __syncthreads();
    float2* yo = y + (size_t)c * stride;
    const int t = threadIdx.x;
    const int ks[4] = {t, 511 - t, 512 + t, 1023 - t};
#pragma unroll 1
    for (int q = 0; q < 4; q++) {
        const int k = ks[q];
        double ar = 0.0, ai = 0.0;
        for (int m = 0; m <= k; m++) {          // ascending m: the small leading taps first
            const float2 hh = hs[m], x = xs[k - m];
            ar += (double)hh.x * (double)x.x - (double)hh.y * (double)x.y;
            ai += (double)hh.x * (double)x.y + (double)hh.y * (double)x.x;
        }
        yo[k] = make_float2((float)ar, (float)ai);
    }
}

// ------------------------------------------------------------------------------------------
// Filter design on the device: CFastFIR::SetupParameters, dsp/fastfir.cpp:207-254 -- 1025-tap Blackman-Nuttall
// windowed sinc, shifted to the band centre, scaled by 1/2048, zero-padded and transformed. One CTA per filter,
// double precision throughout (taps and a radix-2 Stockham FFT in shared memory), rounded to float32 at the end.
// ------------------------------------------------------------------------------------------
struct FirJob { double lo, hi, rate; int row, pad; };

__global__ void __launch_bounds__(256) k_fir_design(const FirJob* __restrict__ jobs, float2* __restrict__ H, float2* __restrict__ taps)
{
    extern __shared__ double2 fd_sm[];
    double2* a = fd_sm;
    double2* b = fd_sm + kFirFft;
    const FirJob job = jobs[blockIdx.x];
    const double nFL = job.lo / job.rate, nFH = job.hi / job.rate;
    const double nFc = (nFH - nFL) / 2.0;
    const double nFs = kTwoPi * (nFH + nFL) / 2.0;
    const double centre = 0.5 * (double)(kFirTaps - 1);
    float2* trow = taps + (size_t)job.row * kFirTapRow;
    for (int i = threadIdx.x; i < kFirFft; i += 256) {
        double re = 0.0, im = 0.0;
        if (i < kFirTaps) {
            const double x = (double)i - centre;
            double z;
            if ((double)i == centre) z = 2.0 * nFc;
            else {
                const double w = (0.3635819 - 0.4891775 * cos((kTwoPi * i) / (kFirTaps - 1)) +
                                  0.1365995 * cos((2.0 * kTwoPi * i) / (kFirTaps - 1)) -
                                  0.0106411 * cos((3.0 * kTwoPi * i) / (kFirTaps - 1)));
                z = sin(kTwoPi * x * nFc) / (kPi * x) * w;
            }
            re = z * cos(nFs * x);
            im = z * sin(nFs * x);
            trow[i] = make_float2((float)re, (float)im);
        } else if (i < kFirTapRow) trow[i] = make_float2(0.f, 0.f);
        a[i] = make_double2(re / (double)kFirFft, im / (double)kFirFft);
    }
    __syncthreads();
    // radix-2 Stockham, e^{-j 2 pi nk/N}
    double2* src = a;
    double2* dst = b;
    for (int ns = 1; ns < kFirFft; ns <<= 1) {
        for (int j = threadIdx.x; j < kFirFft / 2; j += 256) {
            const int k = j & (ns - 1);
            double sn, cs;
            sincospi(-(double)k / (double)ns, &sn, &cs);
            const double2 u = src[j], v = src[j + kFirFft / 2];
            const double2 tv = make_double2(v.x * cs - v.y * sn, v.x * sn + v.y * cs);
            const int j0 = ((j - k) << 1) + k;
            dst[j0] = make_double2(u.x + tv.x, u.y + tv.y);
            dst[j0 + ns] = make_double2(u.x - tv.x, u.y - tv.y);
        }
        __syncthreads();
        double2* t = src; src = dst; dst = t;
    }
    float2* hrow = H + (size_t)job.row * kFirFft;
    for (int i = threadIdx.x; i < kFirFft; i += 256) hrow[i] = make_float2((float)src[i].x, (float)src[i].y);
}

// ------------------------------------------------------------------------------------------
// FirBank
// ------------------------------------------------------------------------------------------
FirBank::~FirBank()
{
    cudaFree(d_H_);
    cudaFree(d_h_);
    cudaFree(d_jobs_);
    cudaFree(d_id_);
    cudaFree(d_tw_);
}

int FirBank::init(int nch, int stride, cudaStream_t st, LaunchCounter* lc)
{
    static_assert(sizeof(FirJob) == sizeof(Job), "job layout");
    nch_ = nch; stride_ = stride; st_ = st; lc_ = lc;
    cur_.assign(stride, Params{-1.0, 1.0, 1.0, 1.0});     // CFastFIR ctor, dsp/fastfir.cpp:126-129
    h_id_.assign(stride, 0);
    // row 0 = all zeros (a channel that never had a valid SetupParameters); every other row is referenced by at
    // least one channel, so (slots + 1) rows always suffice
    cap_ = stride + 1;
    refs_.assign(cap_, 0);
    row_key_.assign(cap_, Key{0, 0, 0});
    free_.clear();
    for (int r = cap_ - 1; r >= 1; r--) free_.push_back(r);
    rows_in_use_ = 1;
    CSDR_CK(cudaMalloc(&d_H_, (size_t)cap_ * kFirFft * sizeof(float2)));
    CSDR_CK(cudaMalloc(&d_h_, (size_t)cap_ * kFirTapRow * sizeof(float2)));
    CSDR_CK(cudaMemsetAsync(d_H_, 0, (size_t)kFirFft * sizeof(float2), st_));
    CSDR_CK(cudaMemsetAsync(d_h_, 0, (size_t)kFirTapRow * sizeof(float2), st_));
    CSDR_CK(cudaMalloc(&d_jobs_, (size_t)cap_ * sizeof(Job)));
    CSDR_CK(cudaMalloc(&d_id_, stride * sizeof(int)));
    std::vector<float2> tw(1024);
    for (int m = 0; m < 1024; m++) tw[m] = make_float2((float)cos(-kTwoPi * m / 2048.0), (float)sin(-kTwoPi * m / 2048.0));
    CSDR_CK(cudaMalloc(&d_tw_, 1024 * sizeof(float2)));
    CSDR_TRY(stage_.upload(d_tw_, tw.data(), 1024 * sizeof(float2), st_));
    CSDR_CK(cudaFuncSetAttribute(k_fir_design, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * kFirFft * sizeof(double2))));
    ids_dirty_ = true;
    return CUTESDR_OK;
}

void FirBank::release(int i)
{
    const int old = h_id_[i];
    if (old > 0 && --refs_[old] == 0) {
        ids_.erase(row_key_[old]);
        free_.push_back(old);
        rows_in_use_--;
        // a design job that is still queued for this row would be wasted work but harmless; drop it
        for (size_t k = 0; k < jobs_.size(); k++) if (jobs_[k].row == old) { jobs_.erase(jobs_.begin() + k); break; }
    }
    h_id_[i] = 0;
    cur_[i] = Params{-1.0, 1.0, 1.0, 1.0};
    ids_dirty_ = true;
}

int FirBank::setup(int i, double lo, double hi, double offset, double rate)
{
    // CFastFIR::SetupParameters, dsp/fastfir.cpp:178-259
    Params& p = cur_[i];
    if (lo == p.lo && hi == p.hi && offset == p.offset && rate == p.rate) return CUTESDR_OK;
    p = Params{lo, hi, offset, rate};
    lo += offset;
    hi += offset;
    if (lo >= hi || lo >= rate / 2.0 || lo <= -rate / 2.0 || hi >= rate / 2.0 || hi <= -rate / 2.0) {
        // the reference logs "Filter Parameter error" and keeps filtering with the old response
        return CUTESDR_OK;
    }
    Key key{lo, hi, rate};
    auto it = ids_.find(key);
    int id;
    if (it != ids_.end()) id = it->second;
    else {
        if (free_.empty()) {
            // cannot happen while every row but 0 is referenced by a channel; release this channel's own row first
            const Params keep = p;
            release(i);
            cur_[i] = keep;
            if (free_.empty()) { set_error("FirBank: filter table exhausted"); return CUTESDR_E_STATE; }
        }
        id = free_.back();
        free_.pop_back();
        rows_in_use_++;
        row_key_[id] = key;
        ids_[key] = id;
        jobs_.push_back(Job{lo, hi, rate, id, 0});
    }
    if (id != h_id_[i]) {
        refs_[id]++;
        const int old = h_id_[i];
        if (old > 0 && --refs_[old] == 0) {
            ids_.erase(row_key_[old]);
            free_.push_back(old);
            rows_in_use_--;
            for (size_t k = 0; k < jobs_.size(); k++) if (jobs_[k].row == old) { jobs_.erase(jobs_.begin() + k); break; }
        }
        h_id_[i] = id;
        ids_dirty_ = true;
    }
    return CUTESDR_OK;
}

// queue pending designs and the channel -> row table in stream order (no synchronisation)
int FirBank::flush()
{
    if (!jobs_.empty()) {
        const int n = (int)jobs_.size();
        CSDR_TRY(stage_.upload(d_jobs_, jobs_.data(), (size_t)n * sizeof(Job), st_));
        k_fir_design<<<n, 256, 2 * kFirFft * sizeof(double2), st_>>>(reinterpret_cast<const FirJob*>(d_jobs_), d_H_, d_h_);
        lc_->n++;
        CSDR_CK(cudaGetLastError());
        jobs_.clear();
    }
    if (ids_dirty_) {
        CSDR_TRY(stage_.upload(d_id_, h_id_.data(), (size_t)stride_ * sizeof(int), st_));
        ids_dirty_ = false;
    }
    return CUTESDR_OK;
}

int FirBank::run(const float2* d_ring, long long first_burst, int nb, float2* d_y, int y_stride)
{
    if (nb <= 0) return CUTESDR_OK;
    CSDR_TRY(flush());
    if (first_burst == 0) {
        k_fir_first<<<nch_, 256, 0, st_>>>(d_ring, d_h_, d_id_, d_y, y_stride);
        lc_->n++;
        CSDR_CK(cudaGetLastError());
        first_burst++;
        nb--;
        d_y += kBurst;
        if (nb == 0) return CUTESDR_OK;
    }
    dim3 grid(nch_, nb);
    k_fastfir<<<grid, 256, 0, st_>>>(d_ring, first_burst, d_H_, d_id_, d_tw_, d_y, y_stride);
    lc_->n++;
    CSDR_CK(cudaGetLastError());
    return CUTESDR_OK;
}

}  // namespace csdr
