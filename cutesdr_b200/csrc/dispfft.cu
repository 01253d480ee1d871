// dispfft.cu -- kernel 3 of the north star: CFft's display path (dsp/fft.cpp:118-410, 560-589):
// Hann window -> FFT -> |X|^2 -> moving/exponential average -> log10 -> integer screen mapping.
//
// N <= 4096: one CTA does window + FFT (shared-memory Stockham) + power + average + log in ONE
//            kernel; the frame is read once from HBM and only the N averaged bins are written.
// N  > 4096: four-step FFT, N = 256 x N2: kernel A (N2 CTAs: windowed column FFTs of length 256 +
//            twiddle), kernel B (256 CTAs: row FFTs of length N2 fused with power/average/log).
// The reference swaps I and Q and runs its e^{+j} kernel (:280-281); |.|^2 of that equals |DFT|^2
// of the un-swapped input, which is what is computed here. Averages are kept in double (the
// per-bin recursion Sum <- Sum - PwrAve + p runs for the life of the display).
#include "common.cuh"

namespace csdr {

constexpr int kTwLen = 2048;     // e^{-2 pi i m / 4096}, m < 2048

__device__ __forceinline__ float2 cmulf2(float2 a, float2 b)
{
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// In-place-by-ping-pong radix-2 Stockham FFT of length n (power of two, 2..4096) in shared memory.
// Returns the buffer that holds the result. tw: global table e^{-2 pi i m/4096}.
__device__ float2* smem_fft(float2* src, float2* dst, int n, const float2* __restrict__ tw)
{
    const int half = n >> 1;
    int tw_stride = 2048;       // 4096/(2*ns) with ns = 1
    for (int ns = 1; ns < n; ns <<= 1, tw_stride >>= 1) {
        for (int j = threadIdx.x; j < half; j += blockDim.x) {
            const int k = j & (ns - 1);
            const float2 w = __ldg(tw + k * tw_stride);
            const float2 a = src[j];
            const float2 b = cmulf2(src[j + half], w);
            const int j0 = ((j - k) << 1) + k;
            dst[j0] = make_float2(a.x + b.x, a.y + b.y);
            dst[j0 + ns] = make_float2(a.x - b.x, a.y - b.y);
        }
        __syncthreads();
        float2* t = src; src = dst; dst = t;
    }
    return src;
}

struct AveParams {
    int N, total_count, ave_size, ave_count;
    double k_b, k_c;
};

// CpxFFT power section, dsp/fft.cpp:566-589, for FFT bin k (fft-shifted storage: index N/2 = DC)
__device__ __forceinline__ void average_bin(int k, float2 X, const AveParams& p, double* sum, double* pwr_ave, double* ave)
{
    const int j = (k < p.N / 2) ? k + p.N / 2 : k - p.N / 2;
    const double pw = (double)X.x * (double)X.x + (double)X.y * (double)X.y;
    double s;
    if (p.total_count <= p.ave_size) s = sum[j] + pw;
    else s = sum[j] - pwr_ave[j] + pw;
    sum[j] = s;
    const double pa = s / (double)p.ave_count;
    pwr_ave[j] = pa;
    ave[j] = log10(pa + p.k_c) + p.k_b;
}

// ---- N <= 4096: everything in one CTA
__global__ void __launch_bounds__(512) k_dispfft_small(const float2* __restrict__ x, float2 dc, const float* __restrict__ win,
                                                       const float2* __restrict__ tw, AveParams p, double* sum,
                                                       double* pwr_ave, double* ave, int* overload)
{
    extern __shared__ float2 sm[];
    float2* a = sm;
    float2* b = sm + p.N;
    int ov = 0;
    for (int i = threadIdx.x; i < p.N; i += blockDim.x) {
        float2 v = x[i];
        v.x -= dc.x;                           // I/Q DC-offset correction of the display copy, interface/sdrinterface.cpp:889-894
        v.y -= dc.y;
        if (v.x > 32000.0f) ov = 1;            // OVER_LIMIT on I only, dsp/fft.cpp:275-276
        const float w = win[i];
        a[i] = make_float2(w * v.x, w * v.y);
    }
    if (ov) atomicOr(overload, 1);
    __syncthreads();
    float2* r = smem_fft(a, b, p.N, tw);
    for (int k = threadIdx.x; k < p.N; k += blockDim.x) average_bin(k, r[k], p, sum, pwr_ave, ave);
}

// ---- N > 4096, step A: column n2 -> FFT over n1 (length N1), times W_N^{n2 k1}; out[k1][n2]
__global__ void __launch_bounds__(128) k_dispfft_cols(const float2* __restrict__ x, float2 dc, const float* __restrict__ win,
                                                      const float2* __restrict__ tw, int N, int N1, int N2,
                                                      float2* __restrict__ mid, int* overload)
{
    extern __shared__ float2 sm[];
    float2* a = sm;
    float2* b = sm + N1;
    const int n2 = blockIdx.x;
    int ov = 0;
    for (int n1 = threadIdx.x; n1 < N1; n1 += blockDim.x) {
        const int n = n1 * N2 + n2;
        float2 v = x[n];
        v.x -= dc.x;
        v.y -= dc.y;
        if (v.x > 32000.0f) ov = 1;
        const float w = win[n];
        a[n1] = make_float2(w * v.x, w * v.y);
    }
    if (ov) atomicOr(overload, 1);
    __syncthreads();
    float2* r = smem_fft(a, b, N1, tw);
    for (int k1 = threadIdx.x; k1 < N1; k1 += blockDim.x) {
        const int m = (int)(((long long)n2 * k1) % N);
        float s, c;
        sincospif(-2.0f * (float)m / (float)N, &s, &c);      // m/N exact in float (both < 2^24)
        mid[(size_t)k1 * N2 + n2] = cmulf2(r[k1], make_float2(c, s));
    }
}

// ---- step B: row k1 -> FFT over n2 (length N2); bin k = k1 + N1*k2; fused power/average/log
__global__ void __launch_bounds__(128) k_dispfft_rows(const float2* __restrict__ mid, const float2* __restrict__ tw, int N1,
                                                      int N2, AveParams p, double* sum, double* pwr_ave, double* ave)
{
    extern __shared__ float2 sm[];
    float2* a = sm;
    float2* b = sm + N2;
    const int k1 = blockIdx.x;
    for (int n2 = threadIdx.x; n2 < N2; n2 += blockDim.x) a[n2] = mid[(size_t)k1 * N2 + n2];
    __syncthreads();
    float2* r = smem_fft(a, b, N2, tw);
    for (int k2 = threadIdx.x; k2 < N2; k2 += blockDim.x) average_bin(k1 + N1 * k2, r[k2], p, sum, pwr_ave, ave);
}

// ---- CFft::FwdFFT / RevFFT (dsp/fft.cpp:416-426): plain complex transforms of the object's size.
// The reference's "forward" uses the e^{+j} kernel and its "reverse" the e^{-j} kernel, neither
// normalises. e^{+j} transform = conj(DFT(conj x)).
__global__ void __launch_bounds__(512) k_fft_plain_small(float2* __restrict__ x, const float2* __restrict__ tw, int N, int plus_j)
{
    extern __shared__ float2 sm[];
    float2* a = sm;
    float2* b = sm + N;
    for (int i = threadIdx.x; i < N; i += blockDim.x) { float2 v = x[i]; if (plus_j) v.y = -v.y; a[i] = v; }
    __syncthreads();
    float2* r = smem_fft(a, b, N, tw);
    for (int k = threadIdx.x; k < N; k += blockDim.x) { float2 v = r[k]; if (plus_j) v.y = -v.y; x[k] = v; }
}

__global__ void __launch_bounds__(128) k_fft_plain_cols(const float2* __restrict__ x, const float2* __restrict__ tw, int N, int N1,
                                                        int N2, float2* __restrict__ mid, int plus_j)
{
    extern __shared__ float2 sm[];
    float2* a = sm;
    float2* b = sm + N1;
    const int n2 = blockIdx.x;
    for (int n1 = threadIdx.x; n1 < N1; n1 += blockDim.x) { float2 v = x[n1 * N2 + n2]; if (plus_j) v.y = -v.y; a[n1] = v; }
    __syncthreads();
    float2* r = smem_fft(a, b, N1, tw);
    for (int k1 = threadIdx.x; k1 < N1; k1 += blockDim.x) {
        const int m = (int)(((long long)n2 * k1) % N);
        float s, c;
        sincospif(-2.0f * (float)m / (float)N, &s, &c);
        mid[(size_t)k1 * N2 + n2] = cmulf2(r[k1], make_float2(c, s));
    }
}

__global__ void __launch_bounds__(128) k_fft_plain_rows(const float2* __restrict__ mid, const float2* __restrict__ tw, int N1, int N2,
                                                        float2* __restrict__ out, int plus_j)
{
    extern __shared__ float2 sm[];
    float2* a = sm;
    float2* b = sm + N2;
    const int k1 = blockIdx.x;
    for (int n2 = threadIdx.x; n2 < N2; n2 += blockDim.x) a[n2] = mid[(size_t)k1 * N2 + n2];
    __syncthreads();
    float2* r = smem_fft(a, b, N2, tw);
    for (int k2 = threadIdx.x; k2 < N2; k2 += blockDim.x) { float2 v = r[k2]; if (plus_j) v.y = -v.y; out[k1 + N1 * k2] = v; }
}

// ---- GetScreenIntegerFFTData, dsp/fft.cpp:365-407: one thread per pixel
// out2 / max_height2 (optional): a second mapping of the same bins with another height, so the plotter's waterfall
// row (255 levels) and its 2-D trace (gui/plotter.cpp:429-456) come out of one pass over the averaged spectrum.
__global__ void k_screen(const double* __restrict__ ave, int N, int invert, int bin_min, int bin_max, int width,
                         int max_height, double gain, double off, int32_t* __restrict__ out, int max_height2,
                         int32_t* __restrict__ out2)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= width) return;
    const int span = bin_max - bin_min;
    auto ypix_h = [&](int i, int mh) {
        int idx = invert ? (N - i) : i;
        if (idx >= N) idx = N - 1;          // the reference reads one past the end here (i == 0)
        int y = (int)((double)mh * gain * (ave[idx] - off));
        if (y < 0) y = 0;
        if (y > mh) y = mh;
        return y;
    };
    if (out2) {
        if (span > width) {
            const long long lo = ((long long)x * span + width - 1) / width;
            long long hi = ((long long)(x + 1) * span + width - 1) / width - 1;
            if (hi > span) hi = span;
            int best = 0x7fffffff;
            for (long long d = lo; d <= hi; d++) { const int y = ypix_h(bin_min + (int)d, max_height2); if (y < best) best = y; }
            out2[x] = best == 0x7fffffff ? 0 : best;
        } else {
            out2[x] = ypix_h(bin_min + (x * span) / width, max_height2);
        }
    }
    auto ypix = [&](int i) { return ypix_h(i, max_height); };
    if (span > width) {
        // bins i with ((i-bin_min)*width)/span == x are contiguous; the reference keeps the smallest
        // y (strongest signal) among them
        const long long lo = ((long long)x * span + width - 1) / width;
        long long hi = ((long long)(x + 1) * span + width - 1) / width - 1;
        if (hi > span) hi = span;
        int best = 0x7fffffff;
        for (long long d = lo; d <= hi; d++) { const int y = ypix(bin_min + (int)d); if (y < best) best = y; }
        out[x] = best == 0x7fffffff ? 0 : best;
    } else {
        out[x] = ypix(bin_min + (x * span) / width);
    }
}

}  // namespace csdr

using namespace csdr;

struct cutesdr_fft {
    int device = 0;
    cudaStream_t st = 0;
    LaunchCounter lc;
    std::mutex mu;
    // CFft members (dsp/fft.h:58-79)
    bool overload = false, invert = false;
    int ave_count = 0, total_count = 0, size = 1024, last_size = 0, ave_size = 1;
    double k_c = 0, k_b = 0, db_comp = 0, sample_freq = 1000;
    float2 dc = {0.f, 0.f};      // m_NCOSpurOffsetI/Q applied to the display copy (interface/sdrinterface.cpp:889-894)
    // device
    float* d_win = nullptr;
    float2* d_tw = nullptr;
    float2* d_x = nullptr;
    float2* d_mid = nullptr;
    float2* d_plain = nullptr;
    double *d_sum = nullptr, *d_pwr = nullptr, *d_ave = nullptr;
    int* d_flag = nullptr;
    int32_t* d_screen = nullptr;
    int screen_cap = 0;
    float2* h_x = nullptr;      // pinned
    int* h_flag = nullptr;      // pinned landing spot of the overload flag
    bool flag_pending = false;
    cudaEvent_t ev_in = nullptr, ev_done = nullptr;   // put_device_async: frame copied / frame's kernels finished
    bool async_used = false;

    void free_bufs()
    {
        cudaFree(d_win); cudaFree(d_x); cudaFree(d_mid); cudaFree(d_sum); cudaFree(d_pwr); cudaFree(d_ave);
        if (h_x) cudaFreeHost(h_x);
        d_win = nullptr; d_x = nullptr; d_mid = nullptr; d_sum = d_pwr = d_ave = nullptr; h_x = nullptr;
    }
    ~cutesdr_fft()
    {
        if (st) cudaStreamSynchronize(st);
        free_bufs();
        cudaFree(d_tw); cudaFree(d_flag); cudaFree(d_screen); cudaFree(d_plain);
        if (h_flag) cudaFreeHost(h_flag);
        if (ev_in) cudaEventDestroy(ev_in);
        if (ev_done) cudaEventDestroy(ev_done);
        if (st) cudaStreamDestroy(st);
    }
    int reset()
    {
        // ResetFFT, dsp/fft.cpp:248-259 (PwrAve is not cleared there either)
        CSDR_CK(cudaMemsetAsync(d_ave, 0, size * sizeof(double), st));
        CSDR_CK(cudaMemsetAsync(d_sum, 0, size * sizeof(double), st));
        ave_count = 0;
        total_count = 0;
        return CUTESDR_OK;
    }
    int set_params(int sz, bool inv, double dbc, double fs)
    {
        // SetFFTParams, dsp/fft.cpp:118-243
        if (sz == 0) return CUTESDR_OK;
        invert = inv;
        sample_freq = fs;
        if (db_comp != dbc) { last_size = 0; db_comp = dbc; }
        if (sz < 512) size = 512;
        else if (sz > 65536) size = 65536;
        else size = sz;
        if (size & (size - 1)) { set_error("FFT size %d is not a power of two", size); return CUTESDR_E_ARG; }
        if (last_size != size) {
            last_size = size;
            free_bufs();
            const int N = size;
            CSDR_CK(cudaMalloc(&d_win, N * sizeof(float)));
            CSDR_CK(cudaMalloc(&d_x, N * sizeof(float2)));
            CSDR_CK(cudaMalloc(&d_mid, N * sizeof(float2)));
            CSDR_CK(cudaMalloc(&d_sum, N * sizeof(double)));
            CSDR_CK(cudaMalloc(&d_pwr, N * sizeof(double)));
            CSDR_CK(cudaMalloc(&d_ave, N * sizeof(double)));
            CSDR_CK(cudaHostAlloc(&h_x, N * sizeof(float2), cudaHostAllocDefault));
            CSDR_CK(cudaMemsetAsync(d_pwr, 0, N * sizeof(double), st));
            k_b = db_comp - 20 * log10((double)N * 32767.0 / 2.0);      // :186-188
            k_c = pow(10.0, (-220.0 - k_b) / 10.0);
            k_b = k_b / 10.0;
            std::vector<float> w(N);
            for (int i = 0; i < N; i++) w[i] = (float)(2.0 * (.5 - .5 * cos((kTwoPi * i) / (N - 1))));   // Hann x2, :196-198
            CSDR_CK(cudaMemcpyAsync(d_win, w.data(), N * sizeof(float), cudaMemcpyHostToDevice, st));
            CSDR_CK(cudaStreamSynchronize(st));
        }
        return reset();
    }
    int put_device(const float2* d_in, int n, int* total)
    {
        CSDR_TRY(put_kernels(d_in, n, total));
        CSDR_CK(cudaStreamSynchronize(st));
        overload = *h_flag != 0;
        flag_pending = false;
        return CUTESDR_OK;
    }
    // the frame's kernels, queued on st; the overload flag lands in pinned memory and is picked up by the next
    // synchronising call (put_device, GetScreenIntegerFFTData)
    int put_kernels(const float2* d_in, int n, int* total)
    {
        // PutInDisplayFFT, dsp/fft.cpp:267-288
        if (n != size) { set_error("PutInDisplayFFT needs exactly %d samples (got %d)", size, n); return CUTESDR_E_ARG; }
        total_count++;
        if (ave_count < ave_size) ave_count++;
        CSDR_CK(cudaMemsetAsync(d_flag, 0, sizeof(int), st));
        AveParams p{size, total_count, ave_size, ave_count, k_b, k_c};
        if (size <= 4096) {
            size_t smem = 2 * (size_t)size * sizeof(float2);
            if (smem > 48 * 1024) CSDR_CK(cudaFuncSetAttribute(k_dispfft_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_dispfft_small<<<1, 512, smem, st>>>(d_in, dc, d_win, d_tw, p, d_sum, d_pwr, d_ave, d_flag);
            lc.n++;
        } else {
            const int N1 = 256, N2 = size / 256;
            k_dispfft_cols<<<N2, 128, 2 * N1 * sizeof(float2), st>>>(d_in, dc, d_win, d_tw, size, N1, N2, d_mid, d_flag);
            k_dispfft_rows<<<N1, 128, 2 * N2 * sizeof(float2), st>>>(d_mid, d_tw, N1, N2, p, d_sum, d_pwr, d_ave);
            lc.n += 2;
        }
        CSDR_CK(cudaGetLastError());
        CSDR_CK(cudaMemcpyAsync(h_flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
        flag_pending = true;
        if (total) *total = total_count;
        return CUTESDR_OK;
    }
};

extern "C" {

int cutesdr_fft_create(cutesdr_fft** out, int device)
{
    if (!out) { set_error("fft_create: bad arguments"); return CUTESDR_E_ARG; }
    *out = nullptr;
    CSDR_CK(cudaSetDevice(device));
    std::unique_ptr<cutesdr_fft> f(new cutesdr_fft());
    f->device = device;
    CSDR_CK(cudaStreamCreateWithFlags(&f->st, cudaStreamNonBlocking));
    std::vector<float2> tw(kTwLen);
    for (int m = 0; m < kTwLen; m++) tw[m] = make_float2((float)cos(-kTwoPi * m / 4096.0), (float)sin(-kTwoPi * m / 4096.0));
    CSDR_CK(cudaMalloc(&f->d_tw, kTwLen * sizeof(float2)));
    CSDR_CK(cudaMemcpy(f->d_tw, tw.data(), kTwLen * sizeof(float2), cudaMemcpyHostToDevice));
    CSDR_CK(cudaMalloc(&f->d_flag, sizeof(int)));
    CSDR_CK(cudaHostAlloc(&f->h_flag, sizeof(int), cudaHostAllocDefault));
    *f->h_flag = 0;
    CSDR_CK(cudaEventCreateWithFlags(&f->ev_in, cudaEventDisableTiming));
    CSDR_CK(cudaEventCreateWithFlags(&f->ev_done, cudaEventDisableTiming));
    // CFft::CFft(): SetFFTParams(2048, FALSE, 0.0, 1000); SetFFTAve(1)  (dsp/fft.cpp:29-50)
    CSDR_TRY(f->set_params(2048, false, 0.0, 1000));
    *out = f.release();
    return CUTESDR_OK;
}

void cutesdr_fft_destroy(cutesdr_fft* h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    delete h;
}

int cutesdr_fft_set_params(cutesdr_fft* h, int size, int invert, double db_compensation, double sample_freq)
{
    if (!h) { set_error("fft_set_params: bad handle"); return CUTESDR_E_ARG; }
    std::lock_guard<std::mutex> lk(h->mu);
    CSDR_CK(cudaSetDevice(h->device));
    return h->set_params(size, invert != 0, db_compensation, sample_freq);
}

int cutesdr_fft_set_ave(cutesdr_fft* h, int ave)
{
    if (!h) { set_error("fft_set_ave: bad handle"); return CUTESDR_E_ARG; }
    std::lock_guard<std::mutex> lk(h->mu);
    CSDR_CK(cudaSetDevice(h->device));
    if (h->ave_size != ave) h->ave_size = ave > 0 ? ave : 1;     // dsp/fft.cpp:103-113
    return h->reset();
}

int cutesdr_fft_reset(cutesdr_fft* h)
{
    if (!h) { set_error("fft_reset: bad handle"); return CUTESDR_E_ARG; }
    std::lock_guard<std::mutex> lk(h->mu);
    CSDR_CK(cudaSetDevice(h->device));
    return h->reset();
}

int cutesdr_fft_put_f32(cutesdr_fft* h, int n, const float* in, int* total_count)
{
    if (!h || !in) { set_error("fft_put: bad arguments"); return CUTESDR_E_ARG; }
    std::lock_guard<std::mutex> lk(h->mu);
    CSDR_CK(cudaSetDevice(h->device));
    if (n != h->size) { set_error("PutInDisplayFFT needs exactly %d samples (got %d)", h->size, n); return CUTESDR_E_ARG; }
    memcpy(h->h_x, in, (size_t)n * sizeof(float2));
    CSDR_CK(cudaMemcpyAsync(h->d_x, h->h_x, (size_t)n * sizeof(float2), cudaMemcpyHostToDevice, h->st));
    return h->put_device(h->d_x, n, total_count);
}

int cutesdr_fft_put(cutesdr_fft* h, int n, const double* in, int* total_count)
{
    if (!h || !in) { set_error("fft_put: bad arguments"); return CUTESDR_E_ARG; }
    std::lock_guard<std::mutex> lk(h->mu);
    CSDR_CK(cudaSetDevice(h->device));
    if (n != h->size) { set_error("PutInDisplayFFT needs exactly %d samples (got %d)", h->size, n); return CUTESDR_E_ARG; }
    for (int i = 0; i < n; i++) h->h_x[i] = make_float2((float)in[2 * i], (float)in[2 * i + 1]);
    CSDR_CK(cudaMemcpyAsync(h->d_x, h->h_x, (size_t)n * sizeof(float2), cudaMemcpyHostToDevice, h->st));
    return h->put_device(h->d_x, n, total_count);
}

int cutesdr_fft_put_device(cutesdr_fft* h, int n, const void* d_in, int* total_count)
{
    if (!h || !d_in) { set_error("fft_put_device: bad arguments"); return CUTESDR_E_ARG; }
    std::lock_guard<std::mutex> lk(h->mu);
    CSDR_CK(cudaSetDevice(h->device));
    return h->put_device(reinterpret_cast<const float2*>(d_in), n, total_count);
}

int cutesdr_fft_put_device_async(cutesdr_fft* h, int n, const void* d_in, void* src_stream, int* total_count)
{
    if (!h || !d_in) { set_error("fft_put_device_async: bad arguments"); return CUTESDR_E_ARG; }
    std::lock_guard<std::mutex> lk(h->mu);
    CSDR_CK(cudaSetDevice(h->device));
    if (n != h->size) { set_error("PutInDisplayFFT needs exactly %d samples (got %d)", h->size, n); return CUTESDR_E_ARG; }
    cudaStream_t src = reinterpret_cast<cudaStream_t>(src_stream);
    // the frame is copied on the producer's stream (stream-ordered with whatever wrote it, and it may be overwritten
    // right after), the transform runs on this object's own stream beside the producer's later work
    if (h->async_used) CSDR_CK(cudaStreamWaitEvent(src, h->ev_done, 0));
    CSDR_CK(cudaMemcpyAsync(h->d_x, d_in, (size_t)n * sizeof(float2), cudaMemcpyDeviceToDevice, src));
    CSDR_CK(cudaEventRecord(h->ev_in, src));
    CSDR_CK(cudaStreamWaitEvent(h->st, h->ev_in, 0));
    CSDR_TRY(h->put_kernels(h->d_x, n, total_count));
    CSDR_CK(cudaEventRecord(h->ev_done, h->st));
    h->async_used = true;
    return CUTESDR_OK;
}

static int fft_screen(cutesdr_fft* h, int max_height, int max_width, double max_db, double min_db, int start_freq, int stop_freq,
                      int32_t* out, int max_height2, int32_t* out2, int* overload)
{
    if (!h || !out || max_width <= 0) { set_error("fft_get_screen: bad arguments"); return CUTESDR_E_ARG; }
    std::lock_guard<std::mutex> lk(h->mu);
    CSDR_CK(cudaSetDevice(h->device));
    // dsp/fft.cpp:329-345
    const int N = h->size, maxbin = N - 1;
    int bin_min = (int)((double)start_freq * (double)N / h->sample_freq) + N / 2;
    int bin_max = (int)((double)stop_freq * (double)N / h->sample_freq) + N / 2;
    if (bin_min < 0) bin_min = 0;
    if (bin_min >= maxbin) bin_min = maxbin;
    if (bin_max < 0) bin_max = 0;
    if (bin_max >= maxbin) bin_max = maxbin;
    const double off = max_db / 10.0;
    const double gain = -10.0 / (max_db - min_db);
    if (max_width > h->screen_cap) {
        cudaFree(h->d_screen);
        h->d_screen = nullptr;
        CSDR_CK(cudaMalloc(&h->d_screen, 2 * (size_t)max_width * sizeof(int32_t)));
        h->screen_cap = max_width;
    }
    k_screen<<<(max_width + 127) / 128, 128, 0, h->st>>>(h->d_ave, N, h->invert ? 1 : 0, bin_min, bin_max, max_width, max_height,
                                                         gain, off, h->d_screen, max_height2, out2 ? h->d_screen + max_width : nullptr);
    h->lc.n++;
    CSDR_CK(cudaGetLastError());
    CSDR_CK(cudaMemcpyAsync(out, h->d_screen, (size_t)max_width * sizeof(int32_t), cudaMemcpyDeviceToHost, h->st));
    if (out2) CSDR_CK(cudaMemcpyAsync(out2, h->d_screen + max_width, (size_t)max_width * sizeof(int32_t), cudaMemcpyDeviceToHost, h->st));
    CSDR_CK(cudaStreamSynchronize(h->st));
    if (h->flag_pending) { h->overload = *h->h_flag != 0; h->flag_pending = false; }
    if (overload) *overload = h->overload ? 1 : 0;
    return CUTESDR_OK;
}

int cutesdr_fft_get_screen(cutesdr_fft* h, int max_height, int max_width, double max_db, double min_db, int start_freq,
                           int stop_freq, int32_t* out, int* overload)
{
    return fft_screen(h, max_height, max_width, max_db, min_db, start_freq, stop_freq, out, 0, nullptr, overload);
}

int cutesdr_fft_get_plot(cutesdr_fft* h, int trace_height, int width, double max_db, double min_db, int start_freq, int stop_freq,
                         int32_t* waterfall_row, int32_t* trace, int* overload)
{
    if (!waterfall_row || !trace) { set_error("fft_get_plot: bad arguments"); return CUTESDR_E_ARG; }
    return fft_screen(h, 255, width, max_db, min_db, start_freq, stop_freq, waterfall_row, trace_height, trace, overload);
}

int cutesdr_fft_set_dc_offset(cutesdr_fft* h, double off_i, double off_q)
{
    if (!h) { set_error("fft_set_dc_offset: bad handle"); return CUTESDR_E_ARG; }
    std::lock_guard<std::mutex> lk(h->mu);
    h->dc = make_float2((float)off_i, (float)off_q);
    return CUTESDR_OK;
}

int cutesdr_fft_get_ave(cutesdr_fft* h, float* out, int cap)
{
    if (!h || !out) { set_error("fft_get_ave: bad arguments"); return CUTESDR_E_ARG; }
    std::lock_guard<std::mutex> lk(h->mu);
    CSDR_CK(cudaSetDevice(h->device));
    const int n = std::min(cap, h->size);
    std::vector<double> tmp(n);
    CSDR_CK(cudaMemcpyAsync(tmp.data(), h->d_ave, n * sizeof(double), cudaMemcpyDeviceToHost, h->st));
    CSDR_CK(cudaStreamSynchronize(h->st));
    for (int i = 0; i < n; i++) out[i] = (float)tmp[i];
    return n;
}

static int fft_plain(cutesdr_fft* h, double* io, int plus_j)
{
    if (!h || !io) { set_error("fft fwd/rev: bad arguments"); return CUTESDR_E_ARG; }
    std::lock_guard<std::mutex> lk(h->mu);
    CSDR_CK(cudaSetDevice(h->device));
    const int N = h->size;
    for (int i = 0; i < N; i++) h->h_x[i] = make_float2((float)io[2 * i], (float)io[2 * i + 1]);
    CSDR_CK(cudaMemcpyAsync(h->d_x, h->h_x, (size_t)N * sizeof(float2), cudaMemcpyHostToDevice, h->st));
    float2* result = h->d_x;
    if (N <= 4096) {
        size_t smem = 2 * (size_t)N * sizeof(float2);
        if (smem > 48 * 1024) CSDR_CK(cudaFuncSetAttribute(k_fft_plain_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_fft_plain_small<<<1, 512, smem, h->st>>>(h->d_x, h->d_tw, N, plus_j);
        h->lc.n++;
    } else {
        const int N1 = 256, N2 = N / 256;
        if (!h->d_plain) CSDR_CK(cudaMalloc(&h->d_plain, (size_t)65536 * sizeof(float2)));
        k_fft_plain_cols<<<N2, 128, 2 * N1 * sizeof(float2), h->st>>>(h->d_x, h->d_tw, N, N1, N2, h->d_mid, plus_j);
        k_fft_plain_rows<<<N1, 128, 2 * N2 * sizeof(float2), h->st>>>(h->d_mid, h->d_tw, N1, N2, h->d_plain, plus_j);
        h->lc.n += 2;
        result = h->d_plain;
    }
    CSDR_CK(cudaGetLastError());
    CSDR_CK(cudaMemcpyAsync(h->h_x, result, (size_t)N * sizeof(float2), cudaMemcpyDeviceToHost, h->st));
    CSDR_CK(cudaStreamSynchronize(h->st));
    for (int i = 0; i < N; i++) { io[2 * i] = h->h_x[i].x; io[2 * i + 1] = h->h_x[i].y; }
    return CUTESDR_OK;
}

int cutesdr_fft_fwd(cutesdr_fft* h, double* io) { return fft_plain(h, io, 1); }
int cutesdr_fft_rev(cutesdr_fft* h, double* io) { return fft_plain(h, io, 0); }

int cutesdr_fft_launch_count(cutesdr_fft* h, long long* n)
{
    if (!h || !n) { set_error("fft_launch_count: bad arguments"); return CUTESDR_E_ARG; }
    *n = h->lc.n;
    return CUTESDR_OK;
}

}  // extern "C"
