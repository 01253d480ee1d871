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
// The FFT is a Stockham autosort in three passes (radix 16, 16, 8) with the butterflies in registers.
#include "fastfir.cuh"

namespace csdr {

// ------------------------------------------------------------------------------------------
// device
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 cmulf(float2 a, float2 b)
{
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// 2048 = 16 x 16 x 8: three Stockham passes (radix 16, 16, 8) with the butterflies in REGISTERS. 128 threads, thread j
// owns butterfly j of the radix-16 passes (inputs j + 128 r, r = 0..15) and butterflies j, j + 128 of the radix-8 pass
// (inputs jj + 256 r). Between passes the data crosses shared memory once (one padded buffer: element i at i + i / 16,
// which makes the stride-16 writes of the first pass and every other access pattern below conflict-free). The last
// forward pass leaves thread j with elements j + 128 m, m = 0..15 -- exactly the inputs of the first inverse pass -- so
// the spectrum is multiplied by H in registers and never written out: 4 shared-memory crossings per overlap-save
// window instead of 14, straight from the decimator ring to the burst row.
// tw[m] = e^{-2 pi i m / 2048}, m < 1024; CONJ selects the inverse kernel.
template <bool CONJ>
__device__ __forceinline__ float2 twiddle(const float2* __restrict__ tw, int m)
{
    // m in [0, 2048): the table holds the first half, the second half is its negation
    float2 w = __ldg(tw + (m & 1023));
    if (m & 1024) { w.x = -w.x; w.y = -w.y; }
    if (CONJ) w.y = -w.y;
    return w;
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
// multiply by -i (forward kernels) or +i (inverse)
template <bool CONJ>
__device__ __forceinline__ float2 rot90(float2 a) { return CONJ ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x); }
// multiply by e^{-+2 pi i m / 16}
template <bool CONJ, int M>
__device__ __forceinline__ float2 mul_w16(float2 a)
{
    constexpr float c1 = 0.92387953251128674f, s1 = 0.38268343236508977f, h = 0.70710678118654752f;
    constexpr int m = M & 15;
    if (m == 0) return a;
    if (m == 4) return rot90<CONJ>(a);
    if (m == 8) return make_float2(-a.x, -a.y);
    if (m == 12) return rot90<!CONJ>(a);
    // w = (wr, -+wi)
    constexpr float wr = (m == 1 || m == 15) ? c1 : (m == 2 || m == 14) ? h : (m == 3 || m == 13) ? s1 : (m == 5 || m == 11) ? -s1
                        : (m == 6 || m == 10) ? -h : -c1;
    constexpr float wi0 = (m == 1 || m == 7) ? s1 : (m == 2 || m == 6) ? h : (m == 3 || m == 5) ? c1 : (m == 9 || m == 15) ? -s1
                         : (m == 10 || m == 14) ? -h : -c1;       // sin(2 pi m / 16)
    const float wi = CONJ ? wi0 : -wi0;
    return make_float2(a.x * wr - a.y * wi, a.x * wi + a.y * wr);
}
// 4-point DFT in place: x[q] = sum_r a[r] e^{-+2 pi i r q / 4}
template <bool CONJ>
__device__ __forceinline__ void dft4(float2& a0, float2& a1, float2& a2, float2& a3)
{
    const float2 s02 = cadd(a0, a2), d02 = csub(a0, a2), s13 = cadd(a1, a3), r13 = rot90<CONJ>(csub(a1, a3));
    a0 = cadd(s02, s13);
    a1 = cadd(d02, r13);
    a2 = csub(s02, s13);
    a3 = csub(d02, r13);
}
// 16-point DFT: input a[r], output a[q] (natural order). r = 4 r1 + r0, q = q0 + 4 q1:
// w16^(rq) = w4^(r1 q0) w16^(r0 q0) w4^(r0 q1)
template <bool CONJ>
__device__ __forceinline__ void dft16(float2* a)
{
    // step A: over r1 for every r0 -> B[r0][q0] stored at a[4 q0 + r0]
    dft4<CONJ>(a[0], a[4], a[8], a[12]);
    dft4<CONJ>(a[1], a[5], a[9], a[13]);
    dft4<CONJ>(a[2], a[6], a[10], a[14]);
    dft4<CONJ>(a[3], a[7], a[11], a[15]);
    // step B: B[r0][q0] *= w16^(r0 q0)
    a[5] = mul_w16<CONJ, 1>(a[5]);   a[6] = mul_w16<CONJ, 2>(a[6]);   a[7] = mul_w16<CONJ, 3>(a[7]);
    a[9] = mul_w16<CONJ, 2>(a[9]);   a[10] = mul_w16<CONJ, 4>(a[10]); a[11] = mul_w16<CONJ, 6>(a[11]);
    a[13] = mul_w16<CONJ, 3>(a[13]); a[14] = mul_w16<CONJ, 6>(a[14]); a[15] = mul_w16<CONJ, 9>(a[15]);
    // step C: over r0 for every q0 -> X[q0 + 4 q1] lands at a[4 q0 + q1]
    dft4<CONJ>(a[0], a[1], a[2], a[3]);
    dft4<CONJ>(a[4], a[5], a[6], a[7]);
    dft4<CONJ>(a[8], a[9], a[10], a[11]);
    dft4<CONJ>(a[12], a[13], a[14], a[15]);
}
// X[q], q = q0 + 4 q1, sits at a[4 q0 + q1] after dft16
__device__ __forceinline__ constexpr int dft16_slot(int q) { return 4 * (q & 3) + (q >> 2); }
// 8-point DFT: r = 4 r1 + r0, q = q0 + 2 q1: w8^(rq) = (-1)^(r1 q0) w8^(r0 q0) w4^(r0 q1). Output X[q0 + 2 q1] at a[4 q0 + q1].
template <bool CONJ>
__device__ __forceinline__ void dft8(float2* a)
{
#pragma unroll
    for (int r0 = 0; r0 < 4; r0++) { const float2 s = cadd(a[r0], a[r0 + 4]), d = csub(a[r0], a[r0 + 4]); a[r0] = s; a[r0 + 4] = d; }
    a[5] = mul_w16<CONJ, 2>(a[5]);
    a[6] = mul_w16<CONJ, 4>(a[6]);
    a[7] = mul_w16<CONJ, 6>(a[7]);
    dft4<CONJ>(a[0], a[1], a[2], a[3]);
    dft4<CONJ>(a[4], a[5], a[6], a[7]);
}
__device__ __forceinline__ constexpr int dft8_slot(int q) { return 4 * (q & 1) + (q >> 1); }

__device__ __forceinline__ int fpad(int i) { return i + (i >> 4); }
constexpr int kFftPadded = kFirFft + kFirFft / 16;

// passes 1 and 2 of a transform: a[r] = element j + 128 r on entry; on exit the data is in `buf` (pass-2 output order)
template <bool CONJ>
__device__ __forceinline__ void fft2048_head(float2* a, float2* buf, const float2* __restrict__ tw, int j)
{
    dft16<CONJ>(a);                                            // ns = 1: no twiddles, outputs 16 j + q
#pragma unroll
    for (int q = 0; q < 16; q++) buf[17 * j + q] = a[dft16_slot(q)];          // fpad(16 j + q) = 17 j + q
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 16; r++) a[r] = buf[fpad(j + 128 * r)];
    __syncthreads();
    const int k = j & 15;                                      // ns = 16: twiddle unit 2048 / 256 = 8
#pragma unroll
    for (int r = 1; r < 16; r++) a[r] = cmulf(a[r], twiddle<CONJ>(tw, 8 * k * r));
    dft16<CONJ>(a);
    const int o = ((j - k) << 4) + k;                          // outputs o + 16 q
#pragma unroll
    for (int q = 0; q < 16; q++) buf[fpad(o + 16 * q)] = a[dft16_slot(q)];
    __syncthreads();
}
// pass 3 (radix 8, ns = 256) for butterfly jj: reads buf, leaves X[jj + 256 q] in a[dft8_slot(q)]
template <bool CONJ>
__device__ __forceinline__ void fft2048_tail(float2* a, const float2* buf, const float2* __restrict__ tw, int jj)
{
#pragma unroll
    for (int r = 0; r < 8; r++) a[r] = buf[fpad(jj + 256 * r)];
#pragma unroll
    for (int r = 1; r < 8; r++) a[r] = cmulf(a[r], twiddle<CONJ>(tw, jj * r));
    dft8<CONJ>(a);
}

__global__ void __launch_bounds__(128) k_fastfir(const float2* __restrict__ ring, long long first_burst,
                                                 const float2* __restrict__ H, const int* __restrict__ filt_id,
                                                 const float2* __restrict__ tw, float2* __restrict__ y, int stride)
{
    __shared__ float2 buf[kFftPadded];
    const int c = blockIdx.x, j = threadIdx.x;
    const long long burst = first_burst + blockIdx.y;
    const long long w0 = burst * kBurst - kBurst;          // first sample of the 2048 window
    const float2* r = ring + (size_t)c * kDecRing;
    float2 a[16], lo[8], hi[8];
#pragma unroll
    for (int m = 0; m < 16; m++) {
        const long long i = w0 + j + 128 * m;
        // samples before the stream start are the reference's zero-initialised overlap buffer
        a[m] = i < 0 ? make_float2(0.f, 0.f) : r[(size_t)(i & (kDecRing - 1))];
    }
    fft2048_head<false>(a, buf, tw, j);
    fft2048_tail<false>(lo, buf, tw, j);                   // X[j + 256 q]
    fft2048_tail<false>(hi, buf, tw, j + 128);             // X[j + 128 + 256 q]
    __syncthreads();                                       // every thread has read buf
    // CpxMpy, dsp/fastfir.cpp:312-321, on the registers: a[m] = H[j + 128 m] X[j + 128 m]
    const float2* Hc = H + (size_t)filt_id[c] * kFirFft;
#pragma unroll
    for (int q = 0; q < 8; q++) {
        a[2 * q] = cmulf(__ldg(Hc + j + 256 * q), lo[dft8_slot(q)]);
        a[2 * q + 1] = cmulf(__ldg(Hc + j + 128 + 256 * q), hi[dft8_slot(q)]);
    }
    fft2048_head<true>(a, buf, tw, j);
    fft2048_tail<true>(lo, buf, tw, j);
    fft2048_tail<true>(hi, buf, tw, j + 128);
    // keep samples 1024..2047 (dsp/fastfir.cpp:291-294); channel-major rows: coalesced stores
    float2* yo = y + (size_t)c * stride + (size_t)blockIdx.y * kBurst;
#pragma unroll
    for (int q = 4; q < 8; q++) {
        yo[j + 256 * q - kBurst] = lo[dft8_slot(q)];
        yo[j + 128 + 256 * q - kBurst] = hi[dft8_slot(q)];
    }
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
    k_fastfir<<<grid, 128, 0, st_>>>(d_ring, first_burst, d_H_, d_id_, d_tw_, d_y, y_stride);
    lc_->n++;
    CSDR_CK(cudaGetLastError());
    return CUTESDR_OK;
}

}  // namespace csdr
