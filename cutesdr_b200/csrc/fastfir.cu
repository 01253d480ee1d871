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
// The FFT is a shared-memory Stockham autosort (five radix-4 passes + one radix-2, ping-pong buffers).
#include "fastfir.cuh"

namespace csdr {

// ------------------------------------------------------------------------------------------
// host-side filter design (double precision)
// ------------------------------------------------------------------------------------------
namespace {
struct cd { double re, im; };

void fft_host(std::vector<cd>& a, int sign)
{
    const int n = (int)a.size();
    for (int i = 1, j = 0; i < n; i++) {
        int bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) std::swap(a[i], a[j]);
    }
    for (int len = 2; len <= n; len <<= 1) {
        const double ang = sign * kTwoPi / len;
        const int half = len >> 1;
        for (int k = 0; k < half; k++) {
            const double wr = cos(ang * k), wi = sin(ang * k);
            for (int i = k; i < n; i += len) {
                cd u = a[i], v = a[i + half];
                double tr = v.re * wr - v.im * wi, ti = v.re * wi + v.im * wr;
                a[i] = {u.re + tr, u.im + ti};
                a[i + half] = {u.re - tr, u.im - ti};
            }
        }
    }
}
}  // namespace

// Frequency response of the band-pass designed by CFastFIR::SetupParameters
// (Blackman-Nuttall windowed sinc, shifted to the band centre, scaled by 1/2048).
static void design_filter(double lo, double hi, double rate, std::vector<float2>& H)
{
    static std::vector<double> window;
    static std::once_flag once;
    std::call_once(once, []() {
        window.resize(kFirTaps);
        for (int i = 0; i < kFirTaps; i++)
            window[i] = (0.3635819 - 0.4891775 * cos((kTwoPi * i) / (kFirTaps - 1)) +
                         0.1365995 * cos((2.0 * kTwoPi * i) / (kFirTaps - 1)) -
                         0.0106411 * cos((3.0 * kTwoPi * i) / (kFirTaps - 1)));
    });
    const double nFL = lo / rate, nFH = hi / rate;
    const double nFc = (nFH - nFL) / 2.0;
    const double nFs = kTwoPi * (nFH + nFL) / 2.0;
    const double centre = 0.5 * (double)(kFirTaps - 1);
    std::vector<cd> a(kFirFft, cd{0.0, 0.0});
    for (int i = 0; i < kFirTaps; i++) {
        const double x = (double)i - centre;
        double z;
        if ((double)i == centre) z = 2.0 * nFc;
        else z = sin(kTwoPi * x * nFc) / (kPi * x) * window[i];
        a[i].re = z * cos(nFs * x) / (double)kFirFft;
        a[i].im = z * sin(nFs * x) / (double)kFirFft;
    }
    fft_host(a, -1);
    H.resize(kFirFft);
    for (int k = 0; k < kFirFft; k++) H[k] = make_float2((float)a[k].re, (float)a[k].im);
}

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
// FirBank
// ------------------------------------------------------------------------------------------
FirBank::~FirBank()
{
    cudaFree(d_H_);
    cudaFree(d_id_);
    cudaFree(d_tw_);
}

int FirBank::init(int nch, int stride, cudaStream_t st, LaunchCounter* lc)
{
    nch_ = nch; stride_ = stride; st_ = st; lc_ = lc;
    cur_.assign(nch, Params{-1.0, 1.0, 1.0, 1.0});        // CFastFIR ctor, dsp/fastfir.cpp:126-129
    h_id_.assign(stride, 0);
    // filter 0 = all zeros (a channel that never had a valid SetupParameters)
    h_H_.assign(kFirFft, make_float2(0.f, 0.f));
    nfilt_ = 1;
    CSDR_CK(cudaMalloc(&d_id_, stride * sizeof(int)));
    std::vector<float2> tw(1024);
    for (int m = 0; m < 1024; m++) tw[m] = make_float2((float)cos(-kTwoPi * m / 2048.0), (float)sin(-kTwoPi * m / 2048.0));
    CSDR_CK(cudaMalloc(&d_tw_, 1024 * sizeof(float2)));
    CSDR_CK(cudaMemcpyAsync(d_tw_, tw.data(), 1024 * sizeof(float2), cudaMemcpyHostToDevice, st_));
    CSDR_CK(cudaStreamSynchronize(st_));
    dirty_ = true;
    return CUTESDR_OK;
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
        std::vector<float2> H;
        design_filter(lo, hi, rate, H);
        id = nfilt_++;
        h_H_.insert(h_H_.end(), H.begin(), H.end());
        ids_[key] = id;
    }
    h_id_[i] = id;
    dirty_ = true;
    return CUTESDR_OK;
}

int FirBank::upload()
{
    if (!dirty_) return CUTESDR_OK;
    if (nfilt_ > cap_) {
        cudaFree(d_H_);
        cap_ = std::max(nfilt_, cap_ * 2);
        CSDR_CK(cudaMalloc(&d_H_, (size_t)cap_ * kFirFft * sizeof(float2)));
    }
    CSDR_CK(cudaMemcpyAsync(d_H_, h_H_.data(), (size_t)nfilt_ * kFirFft * sizeof(float2), cudaMemcpyHostToDevice, st_));
    CSDR_CK(cudaMemcpyAsync(d_id_, h_id_.data(), stride_ * sizeof(int), cudaMemcpyHostToDevice, st_));
    CSDR_CK(cudaStreamSynchronize(st_));
    dirty_ = false;
    return CUTESDR_OK;
}

int FirBank::run(const float2* d_ring, long long first_burst, int nb, float2* d_y, int y_stride)
{
    if (nb <= 0) return CUTESDR_OK;
    CSDR_TRY(upload());
    dim3 grid(nch_, nb);
    k_fastfir<<<grid, 256, 0, st_>>>(d_ring, first_burst, d_H_, d_id_, d_tw_, d_y, y_stride);
    lc_->n++;
    CSDR_CK(cudaGetLastError());
    return CUTESDR_OK;
}

}  // namespace csdr
