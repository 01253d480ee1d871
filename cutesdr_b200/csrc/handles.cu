// handles.cu -- single-object entry points of the C ABI: one reference object each
// (CDownConvert, CFastFIR, CAgc, CFractResampler, CNoiseProc, CDemodulator), built from the same
// batched kernels as the bank with a batch of one. They exist for drop-in compatibility and for
// stage-by-stage parity tests; throughput comes from the bank.
#include "bank.cuh"

#include <algorithm>

using namespace csdr;

namespace {

struct HandleBase {
    int device = 0;
    cudaStream_t st = 0;
    LaunchCounter lc;
    std::mutex mu;
    int open(int dev)
    {
        device = dev;
        CSDR_CK(cudaSetDevice(dev));
        CSDR_CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        return CUTESDR_OK;
    }
    void close()
    {
        if (st) { cudaStreamSynchronize(st); cudaStreamDestroy(st); st = 0; }
    }
};

// host double pairs -> device float2 (through a growing pinned buffer)
struct Stager {
    float2* h = nullptr;
    float2* d = nullptr;
    int cap = 0, front = 0;
    ~Stager() { if (h) cudaFreeHost(h); cudaFree(d); }
    // device buffer = [front | n]; keeps `front` elements before the data pointer
    int ensure(int n, int front_elems)
    {
        if (n <= cap && front_elems == front) return CUTESDR_OK;
        if (h) cudaFreeHost(h);
        cudaFree(d);
        h = nullptr; d = nullptr;
        cap = std::max(n, cap);
        front = front_elems;
        CSDR_CK(cudaHostAlloc(&h, (size_t)cap * sizeof(float2), cudaHostAllocDefault));
        CSDR_CK(cudaMalloc(&d, (size_t)(front + cap) * sizeof(float2)));
        CSDR_CK(cudaMemset(d, 0, (size_t)(front + cap) * sizeof(float2)));
        return CUTESDR_OK;
    }
    float2* data() { return d + front; }
};

}  // namespace

// ================================================================================================
// CDownConvert
// ================================================================================================
struct cutesdr_downconvert : HandleBase {
    // CDownConvert members, dsp/downconvert.cpp:60-73
    double nco_freq = 0.0, cw_offset = 0.0, in_rate = 100000.0, max_bw = 10000.0, out_rate = 0.0;
    std::unique_ptr<Decimator> dec;
    std::vector<int> lens;
    Stager in;
    long long stream_pos = 0;
    int chunk = 0;
    float2* d_halo[2] = {nullptr, nullptr};
    int halo_cur = 0;
    ~cutesdr_downconvert() { if (st) cudaStreamSynchronize(st); dec.reset(); cudaFree(d_halo[0]); cudaFree(d_halo[1]); close(); }

    int build()
    {
        // chain for (in_rate, max_bw); state of all stage filters restarts (DeleteFilters + new)
        out_rate = plan_stages(in_rate, max_bw, lens);
        const int dec_by = 1 << lens.size();
        chunk = 2048 * dec_by;                     // at most 2048 outputs per device pass (ring size)
        dec.reset(new Decimator());
        CSDR_TRY(dec->init(1, in_rate, max_bw, chunk, st, &lc));
        dec->set_frequency(0, nco_freq);
        CSDR_TRY(in.ensure(chunk, 0));
        for (int k = 0; k < 2; k++) {
            if (!d_halo[k]) CSDR_CK(cudaMalloc(&d_halo[k], kHaloMax * sizeof(float2)));
            CSDR_CK(cudaMemsetAsync(d_halo[k], 0, kHaloMax * sizeof(float2), st));
        }
        halo_cur = 0;
        return CUTESDR_OK;
    }

    template <typename TI, typename TO> int process(int n_in, const TI* src, TO* dst)
    {
        if (!dec) CSDR_TRY(build());
        const int dec_by = 1 << lens.size();
        if (n_in % dec_by != 0) {
            set_error("InLength %d must be a multiple of 2^%d (dsp/downconvert.cpp:182-183)", n_in, (int)lens.size());
            return CUTESDR_E_ARG;
        }
        int done = 0, nout = 0;
        std::vector<float2> tmp;
        while (done < n_in) {
            const int m = std::min(chunk, n_in - done);
            for (int i = 0; i < m; i++) in.h[i] = make_float2((float)src[2 * (done + i)], (float)src[2 * (done + i) + 1]);
            CSDR_CK(cudaMemcpyAsync(in.data(), in.h, (size_t)m * sizeof(float2), cudaMemcpyHostToDevice, st));
            CSDR_TRY(apply_nco_startup_gain(in.data(), stream_pos, m, st, &lc));
            CSDR_TRY(dec->run_block(in.data(), d_halo[halo_cur], d_halo[halo_cur ^ 1], m));
            halo_cur ^= 1;
            stream_pos += m;
            const int k = m / dec_by;
            tmp.resize(k);
            const long long first = dec->total_out() - k;
            for (int j = 0; j < k;) {
                const int pos = (int)((first + j) & (kDecRing - 1));
                const int run = std::min(k - j, kDecRing - pos);
                CSDR_CK(cudaMemcpyAsync(tmp.data() + j, dec->ring() + pos, run * sizeof(float2), cudaMemcpyDeviceToHost, st));
                j += run;
            }
            CSDR_CK(cudaStreamSynchronize(st));
            for (int j = 0; j < k; j++) { dst[2 * (nout + j)] = (TO)tmp[j].x; dst[2 * (nout + j) + 1] = (TO)tmp[j].y; }
            nout += k;
            done += m;
        }
        return nout;
    }
};

extern "C" {

int cutesdr_downconvert_create(cutesdr_downconvert** out, int device)
{
    if (!out) { set_error("downconvert_create: bad arguments"); return CUTESDR_E_ARG; }
    *out = nullptr;
    std::unique_ptr<cutesdr_downconvert> h(new cutesdr_downconvert());
    CSDR_TRY(h->open(device));
    *out = h.release();
    return CUTESDR_OK;
}

void cutesdr_downconvert_destroy(cutesdr_downconvert* h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    delete h;
}

int cutesdr_downconvert_set_frequency(cutesdr_downconvert* h, double nco_freq)
{
    if (!h) { set_error("downconvert: bad handle"); return CUTESDR_E_ARG; }
    std::lock_guard<std::mutex> lk(h->mu);
    h->nco_freq = nco_freq + h->cw_offset;          // dsp/downconvert.cpp:100-102
    if (h->dec) h->dec->set_frequency(0, h->nco_freq);
    return CUTESDR_OK;
}

int cutesdr_downconvert_set_cw_offset(cutesdr_downconvert* h, double offset)
{
    if (!h) { set_error("downconvert: bad handle"); return CUTESDR_E_ARG; }
    std::lock_guard<std::mutex> lk(h->mu);
    h->cw_offset = offset;
    return CUTESDR_OK;
}

int cutesdr_downconvert_set_data_rate(cutesdr_downconvert* h, double in_rate, double max_bw, double* out_rate)
{
    if (!h || !(in_rate > 0)) { set_error("downconvert_set_data_rate: bad arguments"); return CUTESDR_E_ARG; }
    std::lock_guard<std::mutex> lk(h->mu);
    CSDR_CK(cudaSetDevice(h->device));
    if (h->in_rate != in_rate || h->max_bw != max_bw || !h->dec) {       // dsp/downconvert.cpp:118-170
        const bool changed = (h->in_rate != in_rate || h->max_bw != max_bw);
        h->in_rate = in_rate;
        h->max_bw = max_bw;
        if (changed) h->nco_freq = h->nco_freq + h->cw_offset;          // SetFrequency(m_NcoFreq), :168
        // NOTE: the reference keeps the oscillator phasor across a chain rebuild; this handle
        // restarts it (documented in DESIGN.md).
        CSDR_TRY(h->build());
    }
    if (out_rate) *out_rate = h->out_rate;
    return CUTESDR_OK;
}

int cutesdr_downconvert_stages(cutesdr_downconvert* h, int* lens, int cap, int* n)
{
    if (!h || !n) { set_error("downconvert_stages: bad arguments"); return CUTESDR_E_ARG; }
    std::lock_guard<std::mutex> lk(h->mu);
    CSDR_CK(cudaSetDevice(h->device));
    if (!h->dec) CSDR_TRY(h->build());
    *n = (int)h->lens.size();
    for (int i = 0; i < *n && i < cap; i++) lens[i] = h->lens[i];
    return CUTESDR_OK;
}

int cutesdr_downconvert_process(cutesdr_downconvert* h, int n_in, const double* in, double* out)
{
    if (!h || n_in < 0 || (n_in && (!in || !out))) { set_error("downconvert_process: bad arguments"); return CUTESDR_E_ARG; }
    std::lock_guard<std::mutex> lk(h->mu);
    CSDR_CK(cudaSetDevice(h->device));
    return h->process(n_in, in, out);
}

int cutesdr_downconvert_process_f32(cutesdr_downconvert* h, int n_in, const float* in, float* out)
{
    if (!h || n_in < 0 || (n_in && (!in || !out))) { set_error("downconvert_process: bad arguments"); return CUTESDR_E_ARG; }
    std::lock_guard<std::mutex> lk(h->mu);
    CSDR_CK(cudaSetDevice(h->device));
    return h->process(n_in, in, out);
}

}  // extern "C"

// ================================================================================================
// CFastFIR
// ================================================================================================
struct cutesdr_fastfir : HandleBase {
    FirBank fir;
    float2* d_ring = nullptr;      // [kDecRing]
    float2* d_y = nullptr;         // [kBurst]
    float2* h_buf = nullptr;       // pinned [kBurst]
    long long total_in = 0, bursts_done = 0;
    ~cutesdr_fastfir()
    {
        if (st) cudaStreamSynchronize(st);
        cudaFree(d_ring); cudaFree(d_y);
        if (h_buf) cudaFreeHost(h_buf);
        close();
    }
    template <typename TI, typename TO> int process(int n_in, const TI* src, TO* dst)
    {
        // dsp/fastfir.cpp:268-306: an FFT fires whenever 1024 new samples have accumulated
        int done = 0, nout = 0;
        while (done < n_in) {
            const int room = (int)(kBurst - (total_in % kBurst));
            const int m = std::min(room, n_in - done);
            for (int i = 0; i < m; i++) h_buf[i] = make_float2((float)src[2 * (done + i)], (float)src[2 * (done + i) + 1]);
            const int pos = (int)(total_in & (kDecRing - 1));
            CSDR_CK(cudaMemcpyAsync(d_ring + pos, h_buf, (size_t)m * sizeof(float2), cudaMemcpyHostToDevice, st));   // m <= room never wraps
            total_in += m;
            done += m;
            if (total_in / kBurst > bursts_done) {
                CSDR_TRY(fir.run(d_ring, bursts_done, 1, d_y, kBurst));
                bursts_done++;
                CSDR_CK(cudaMemcpyAsync(h_buf, d_y, kBurst * sizeof(float2), cudaMemcpyDeviceToHost, st));
                CSDR_CK(cudaStreamSynchronize(st));
                for (int j = 0; j < kBurst; j++) { dst[2 * (nout + j)] = (TO)h_buf[j].x; dst[2 * (nout + j) + 1] = (TO)h_buf[j].y; }
                nout += kBurst;
            } else {
                CSDR_CK(cudaStreamSynchronize(st));
            }
        }
        return nout;
    }
};

extern "C" {

int cutesdr_fastfir_create(cutesdr_fastfir** out, int device)
{
    if (!out) { set_error("fastfir_create: bad arguments"); return CUTESDR_E_ARG; }
    *out = nullptr;
    std::unique_ptr<cutesdr_fastfir> h(new cutesdr_fastfir());
    CSDR_TRY(h->open(device));
    CSDR_TRY(h->fir.init(1, 1, h->st, &h->lc));
    CSDR_CK(cudaMalloc(&h->d_ring, kDecRing * sizeof(float2)));
    CSDR_CK(cudaMemset(h->d_ring, 0, kDecRing * sizeof(float2)));
    CSDR_CK(cudaMalloc(&h->d_y, kBurst * sizeof(float2)));
    CSDR_CK(cudaHostAlloc(&h->h_buf, kBurst * sizeof(float2), cudaHostAllocDefault));
    *out = h.release();
    return CUTESDR_OK;
}

void cutesdr_fastfir_destroy(cutesdr_fastfir* h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    delete h;
}

int cutesdr_fastfir_setup(cutesdr_fastfir* h, double lo_cut, double hi_cut, double offset, double sample_rate)
{
    if (!h) { set_error("fastfir_setup: bad handle"); return CUTESDR_E_ARG; }
    std::lock_guard<std::mutex> lk(h->mu);
    CSDR_CK(cudaSetDevice(h->device));
    return h->fir.setup(0, lo_cut, hi_cut, offset, sample_rate);
}

int cutesdr_fastfir_process(cutesdr_fastfir* h, int n_in, const double* in, double* out)
{
    if (!h || n_in < 0 || (n_in && (!in || !out))) { set_error("fastfir_process: bad arguments"); return CUTESDR_E_ARG; }
    std::lock_guard<std::mutex> lk(h->mu);
    CSDR_CK(cudaSetDevice(h->device));
    return h->process(n_in, in, out);
}

int cutesdr_fastfir_process_f32(cutesdr_fastfir* h, int n_in, const float* in, float* out)
{
    if (!h || n_in < 0 || (n_in && (!in || !out))) { set_error("fastfir_process: bad arguments"); return CUTESDR_E_ARG; }
    std::lock_guard<std::mutex> lk(h->mu);
    CSDR_CK(cudaSetDevice(h->device));
    return h->process(n_in, in, out);
}

}  // extern "C"

// ================================================================================================
// CAgc (complex path)
// ================================================================================================
struct cutesdr_agc : HandleBase {
    std::unique_ptr<PostBank> post;
    double rate = 100.0;           // CAgc ctor, dsp/agc.cpp:88
    float2* h_buf = nullptr;
    int* d_map = nullptr;
    static constexpr int kChunk = 4096;
    ~cutesdr_agc()
    {
        if (st) cudaStreamSynchronize(st);
        post.reset();
        cudaFree(d_map);
        if (h_buf) cudaFreeHost(h_buf);
        close();
    }
};

extern "C" {

int cutesdr_agc_create(cutesdr_agc** out, int device)
{
    if (!out) { set_error("agc_create: bad arguments"); return CUTESDR_E_ARG; }
    *out = nullptr;
    std::unique_ptr<cutesdr_agc> h(new cutesdr_agc());
    CSDR_TRY(h->open(device));
    CSDR_CK(cudaMalloc(&h->d_map, sizeof(int)));
    CSDR_CK(cudaMemset(h->d_map, 0, sizeof(int)));
    CSDR_CK(cudaHostAlloc(&h->h_buf, cutesdr_agc::kChunk * sizeof(float2), cudaHostAllocDefault));
    *out = h.release();
    return CUTESDR_OK;
}

void cutesdr_agc_destroy(cutesdr_agc* h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    delete h;
}

int cutesdr_agc_set_parameters(cutesdr_agc* h, int agc_on, int use_hang, int threshold, int manual_gain, int slope,
                               int decay, double sample_rate)
{
    if (!h || !(sample_rate > 0)) { set_error("agc_set_parameters: bad arguments"); return CUTESDR_E_ARG; }
    std::lock_guard<std::mutex> lk(h->mu);
    CSDR_CK(cudaSetDevice(h->device));
    if (!h->post || h->rate != sample_rate) {
        // a sample-rate change clears the delay line and the averagers (dsp/agc.cpp:121-136)
        h->rate = sample_rate;
        h->post.reset(new PostBank());
        CSDR_TRY(h->post->init(1, 1, sample_rate, cutesdr_agc::kChunk, h->st, &h->lc));
        h->post->set_mode(0, POST_AGC_ONLY);
    }
    h->post->set_agc(0, agc_on, use_hang, threshold, manual_gain, slope, decay);
    return CUTESDR_OK;
}

int cutesdr_agc_process(cutesdr_agc* h, int n, const double* in, double* out)
{
    if (!h || n < 0 || (n && (!in || !out))) { set_error("agc_process: bad arguments"); return CUTESDR_E_ARG; }
    std::lock_guard<std::mutex> lk(h->mu);
    CSDR_CK(cudaSetDevice(h->device));
    if (!h->post) { set_error("agc_process before SetParameters"); return CUTESDR_E_STATE; }
    for (int done = 0; done < n;) {
        const int m = std::min(cutesdr_agc::kChunk, n - done);
        for (int i = 0; i < m; i++) h->h_buf[i] = make_float2((float)in[2 * (done + i)], (float)in[2 * (done + i) + 1]);
        CSDR_CK(cudaMemcpyAsync(h->post->y_in(), h->h_buf, (size_t)m * sizeof(float2), cudaMemcpyHostToDevice, h->st));
        CSDR_TRY(h->post->run(m, nullptr, 0, 0, h->d_map));
        CSDR_CK(cudaMemcpyAsync(h->h_buf, h->post->tap3(), (size_t)m * sizeof(float2), cudaMemcpyDeviceToHost, h->st));
        CSDR_CK(cudaStreamSynchronize(h->st));
        for (int i = 0; i < m; i++) { out[2 * (done + i)] = h->h_buf[i].x; out[2 * (done + i) + 1] = h->h_buf[i].y; }
        done += m;
    }
    return CUTESDR_OK;
}

}  // extern "C"

// ================================================================================================
// CFractResampler
// ================================================================================================
struct cutesdr_resampler : HandleBase {
    std::unique_ptr<ResamplerBank> real_rs, cpx_rs;   // the reference shares one buffer; re/im rows here
    int max_in = 0;
    float* h_in = nullptr;       // pinned [2][max_in]
    float* d_out = nullptr;
    float* h_out = nullptr;
    int16_t* d_out16 = nullptr;
    int16_t* h_out16 = nullptr;
    int out_cap = 0;
    ResamplerBank* rs2 = nullptr;
    ~cutesdr_resampler()
    {
        if (st) cudaStreamSynchronize(st);
        real_rs.reset(); cpx_rs.reset();
        if (h_in) cudaFreeHost(h_in);
        if (h_out) cudaFreeHost(h_out);
        if (h_out16) cudaFreeHost(h_out16);
        cudaFree(d_out); cudaFree(d_out16);
        close();
    }
    int ensure_out(int n)
    {
        if (n <= out_cap) return CUTESDR_OK;
        if (h_out) cudaFreeHost(h_out);
        if (h_out16) cudaFreeHost(h_out16);
        cudaFree(d_out); cudaFree(d_out16);
        out_cap = n + 256;
        CSDR_CK(cudaMalloc(&d_out, (size_t)2 * out_cap * sizeof(float)));
        CSDR_CK(cudaMalloc(&d_out16, (size_t)2 * out_cap * sizeof(int16_t)));
        CSDR_CK(cudaHostAlloc(&h_out, (size_t)2 * out_cap * sizeof(float), cudaHostAllocDefault));
        CSDR_CK(cudaHostAlloc(&h_out16, (size_t)2 * out_cap * sizeof(int16_t), cudaHostAllocDefault));
        return CUTESDR_OK;
    }
    // rows = 1 (real) or 2 (complex); in: interleaved when rows == 2
    int run(int rows, int n, double rate, const double* in, double* out, int16_t* out16, double gain)
    {
        if (!real_rs) { set_error("Resample before Init"); return CUTESDR_E_STATE; }
        if (n > max_in) { set_error("Resample: InLength %d exceeds Init(%d)", n, max_in); return CUTESDR_E_ARG; }
        // The reference keeps ONE complex work buffer and ONE time accumulator for all overloads
        // (dsp/fractresampler.h:29-31): the real overloads use its .re plane. Mirrored with a
        // two-row bank whose row 1 is simply left untouched by the real overloads.
        ResamplerBank& rs = *real_rs;
        if (!(rate > 0)) { set_error("Resample: bad rate"); return CUTESDR_E_ARG; }
        CSDR_TRY(ensure_out(rs.max_out(n, rate)));
        const int stride = rs.in_stride();
        for (int i = 0; i < n; i++) {
            if (rows == 2) { h_in[i] = (float)in[2 * i]; h_in[max_in + i] = (float)in[2 * i + 1]; }
            else h_in[i] = (float)in[i];
        }
        CSDR_CK(cudaMemcpyAsync(rs.in_ptr(), h_in, (size_t)n * sizeof(float), cudaMemcpyHostToDevice, st));
        if (rows == 2) CSDR_CK(cudaMemcpyAsync(rs.in_ptr() + stride, h_in + max_in, (size_t)n * sizeof(float), cudaMemcpyHostToDevice, st));
        int m = 0;
        if (out16) CSDR_TRY(rs.run(n, rate, nullptr, 0, 0, nullptr, &m, d_out16, gain, 2));
        else CSDR_TRY(rs.run(n, rate, d_out, out_cap, 0, nullptr, &m));
        if (out16) {
            CSDR_CK(cudaMemcpyAsync(h_out16, d_out16, (size_t)2 * out_cap * sizeof(int16_t), cudaMemcpyDeviceToHost, st));
            CSDR_CK(cudaStreamSynchronize(st));
            if (rows == 2) memcpy(out16, h_out16, (size_t)2 * m * sizeof(int16_t));
            else for (int k = 0; k < m; k++) out16[k] = h_out16[2 * k];
        } else {
            CSDR_CK(cudaMemcpyAsync(h_out, d_out, (size_t)2 * out_cap * sizeof(float), cudaMemcpyDeviceToHost, st));
            CSDR_CK(cudaStreamSynchronize(st));
            for (int k = 0; k < m; k++) {
                if (rows == 2) { out[2 * k] = h_out[k]; out[2 * k + 1] = h_out[out_cap + k]; }
                else out[k] = h_out[k];
            }
        }
        return m;
    }
};

extern "C" {

int cutesdr_resampler_create(cutesdr_resampler** out, int device)
{
    if (!out) { set_error("resampler_create: bad arguments"); return CUTESDR_E_ARG; }
    *out = nullptr;
    std::unique_ptr<cutesdr_resampler> h(new cutesdr_resampler());
    CSDR_TRY(h->open(device));
    *out = h.release();
    return CUTESDR_OK;
}

void cutesdr_resampler_destroy(cutesdr_resampler* h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    delete h;
}

int cutesdr_resampler_init(cutesdr_resampler* h, int max_input_size)
{
    if (!h || max_input_size <= 0) { set_error("resampler_init: bad arguments"); return CUTESDR_E_ARG; }
    std::lock_guard<std::mutex> lk(h->mu);
    CSDR_CK(cudaSetDevice(h->device));
    h->max_in = max_input_size;
    h->real_rs.reset(new ResamplerBank());
    CSDR_TRY(h->real_rs->init(2, max_input_size, h->st, &h->lc));      // Init zeroes buffer and time, :85-116
    if (h->h_in) cudaFreeHost(h->h_in);
    h->h_in = nullptr;
    CSDR_CK(cudaHostAlloc(&h->h_in, (size_t)2 * max_input_size * sizeof(float), cudaHostAllocDefault));
    return CUTESDR_OK;
}

#define RS_ENTER(h)                                                             \
    if (!h || n < 0) { set_error("resampler: bad arguments"); return CUTESDR_E_ARG; } \
    std::lock_guard<std::mutex> lk(h->mu);                                      \
    CSDR_CK(cudaSetDevice(h->device));

int cutesdr_resampler_real(cutesdr_resampler* h, int n, double rate, const double* in, double* out)
{
    RS_ENTER(h);
    return h->run(1, n, rate, in, out, nullptr, 1.0);
}
int cutesdr_resampler_cpx(cutesdr_resampler* h, int n, double rate, const double* in, double* out)
{
    RS_ENTER(h);
    return h->run(2, n, rate, in, out, nullptr, 1.0);
}
int cutesdr_resampler_mono16(cutesdr_resampler* h, int n, double rate, const double* in, int16_t* out, double gain)
{
    RS_ENTER(h);
    return h->run(1, n, rate, in, nullptr, out, gain);
}
int cutesdr_resampler_stereo16(cutesdr_resampler* h, int n, double rate, const double* in, int16_t* out, double gain)
{
    RS_ENTER(h);
    return h->run(2, n, rate, in, nullptr, out, gain);
}

}  // extern "C"

// ================================================================================================
// CNoiseProc
// ================================================================================================
struct cutesdr_noiseproc : HandleBase {
    std::unique_ptr<Blanker> nb;
    static constexpr int kChunk = 1 << 20;
    float2 *d_in = nullptr, *d_out = nullptr, *h_buf = nullptr;
    // CNoiseProc ctor: SetupBlanker(false, 50.0, 2.0, 1000.0)  (dsp/noiseproc.cpp:64)
    bool on = false;
    double thr = 50.0, width = 2.0, rate = 1000.0;
    ~cutesdr_noiseproc()
    {
        if (st) cudaStreamSynchronize(st);
        nb.reset();
        cudaFree(d_in); cudaFree(d_out);
        if (h_buf) cudaFreeHost(h_buf);
        close();
    }
    template <typename T> int process(int n, const T* in, T* out)
    {
        for (int done = 0; done < n;) {
            const int m = std::min(kChunk, n - done);
            for (int i = 0; i < m; i++) h_buf[i] = make_float2((float)in[2 * (done + i)], (float)in[2 * (done + i) + 1]);
            CSDR_CK(cudaMemcpyAsync(d_in, h_buf, (size_t)m * sizeof(float2), cudaMemcpyHostToDevice, st));
            CSDR_TRY(nb->run(d_in, d_out, m));
            CSDR_CK(cudaMemcpyAsync(h_buf, d_out, (size_t)m * sizeof(float2), cudaMemcpyDeviceToHost, st));
            CSDR_CK(cudaStreamSynchronize(st));
            for (int i = 0; i < m; i++) { out[2 * (done + i)] = (T)h_buf[i].x; out[2 * (done + i) + 1] = (T)h_buf[i].y; }
            done += m;
        }
        return CUTESDR_OK;
    }
};

extern "C" {

int cutesdr_noiseproc_create(cutesdr_noiseproc** out, int device)
{
    if (!out) { set_error("noiseproc_create: bad arguments"); return CUTESDR_E_ARG; }
    *out = nullptr;
    std::unique_ptr<cutesdr_noiseproc> h(new cutesdr_noiseproc());
    CSDR_TRY(h->open(device));
    h->nb.reset(new Blanker());
    CSDR_TRY(h->nb->init(cutesdr_noiseproc::kChunk, h->st, &h->lc));
    CSDR_TRY(h->nb->setup(false, 50.0, 2.0, 1000.0));
    CSDR_CK(cudaMalloc(&h->d_in, (size_t)cutesdr_noiseproc::kChunk * sizeof(float2)));
    CSDR_CK(cudaMalloc(&h->d_out, (size_t)cutesdr_noiseproc::kChunk * sizeof(float2)));
    CSDR_CK(cudaHostAlloc(&h->h_buf, (size_t)cutesdr_noiseproc::kChunk * sizeof(float2), cudaHostAllocDefault));
    *out = h.release();
    return CUTESDR_OK;
}

void cutesdr_noiseproc_destroy(cutesdr_noiseproc* h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    delete h;
}

int cutesdr_noiseproc_setup(cutesdr_noiseproc* h, int on, double threshold, double width_us, double sample_rate)
{
    if (!h) { set_error("noiseproc_setup: bad handle"); return CUTESDR_E_ARG; }
    std::lock_guard<std::mutex> lk(h->mu);
    CSDR_CK(cudaSetDevice(h->device));
    return h->nb->setup(on != 0, threshold, width_us, sample_rate);
}

int cutesdr_noiseproc_process(cutesdr_noiseproc* h, int n, const double* in, double* out)
{
    if (!h || n < 0 || (n && (!in || !out))) { set_error("noiseproc_process: bad arguments"); return CUTESDR_E_ARG; }
    std::lock_guard<std::mutex> lk(h->mu);
    CSDR_CK(cudaSetDevice(h->device));
    return h->process(n, in, out);
}

int cutesdr_noiseproc_process_f32(cutesdr_noiseproc* h, int n, const float* in, float* out)
{
    if (!h || n < 0 || (n && (!in || !out))) { set_error("noiseproc_process: bad arguments"); return CUTESDR_E_ARG; }
    std::lock_guard<std::mutex> lk(h->mu);
    CSDR_CK(cudaSetDevice(h->device));
    return h->process(n, in, out);
}

}  // extern "C"

// ================================================================================================
// CDemodulator = a bank of one
// ================================================================================================
struct cutesdr_demodulator {
    cutesdr_bank* bank = nullptr;
    int device = 0;
    double in_rate = 0.0;
    int mode = -1;
    cutesdr_demod_info info{};
    bool have_info = false;
    double freq = 0.0;
    bool have_freq = false;
    std::vector<float> iq, audio;
};

extern "C" {

int cutesdr_demodulator_create(cutesdr_demodulator** out, int device)
{
    if (!out) { set_error("demodulator_create: bad arguments"); return CUTESDR_E_ARG; }
    int n = 0;
    CSDR_TRY(cutesdr_device_count(&n));
    if (device < 0 || device >= n) { set_error("demodulator_create: no CUDA device %d", device); return CUTESDR_E_CUDA; }
    *out = new cutesdr_demodulator();
    (*out)->device = device;
    return CUTESDR_OK;
}

void cutesdr_demodulator_destroy(cutesdr_demodulator* h)
{
    if (!h) return;
    cutesdr_bank_destroy(h->bank);
    delete h;
}

int cutesdr_demodulator_set_input_sample_rate(cutesdr_demodulator* h, double rate)
{
    if (!h || !(rate > 0)) { set_error("set_input_sample_rate: bad arguments"); return CUTESDR_E_ARG; }
    if (h->bank && h->in_rate == rate) return CUTESDR_OK;           // dsp/demodulator.cpp:96
    cutesdr_bank_destroy(h->bank);
    h->bank = nullptr;
    h->in_rate = rate;
    CSDR_TRY(cutesdr_bank_create(&h->bank, 1, rate, h->device));
    if (h->have_freq) CSDR_TRY(cutesdr_bank_set_demod_freq(h->bank, 0, h->freq));
    if (h->have_info) CSDR_TRY(cutesdr_bank_set_demod(h->bank, 0, h->mode, &h->info));
    return CUTESDR_OK;
}

int cutesdr_demodulator_set_demod(cutesdr_demodulator* h, int mode, const cutesdr_demod_info* info)
{
    if (!h || !info) { set_error("set_demod: bad arguments"); return CUTESDR_E_ARG; }
    h->mode = mode;
    h->info = *info;
    h->have_info = true;
    if (!h->bank) return CUTESDR_OK;
    return cutesdr_bank_set_demod(h->bank, 0, mode, info);
}

int cutesdr_demodulator_set_demod_freq(cutesdr_demodulator* h, double freq)
{
    if (!h) { set_error("set_demod_freq: bad handle"); return CUTESDR_E_ARG; }
    h->freq = freq;
    h->have_freq = true;
    if (!h->bank) return CUTESDR_OK;
    return cutesdr_bank_set_demod_freq(h->bank, 0, freq);
}

int cutesdr_demodulator_get_output_rate(cutesdr_demodulator* h, double* rate)
{
    if (!h || !h->bank) { set_error("get_output_rate: SetInputSampleRate not called"); return CUTESDR_E_STATE; }
    return cutesdr_bank_get_output_rate(h->bank, 0, rate);
}

int cutesdr_demodulator_get_smeter(cutesdr_demodulator* h, double* peak, double* ave)
{
    if (!h || !h->bank) { set_error("get_smeter: SetInputSampleRate not called"); return CUTESDR_E_STATE; }
    return cutesdr_bank_get_smeter(h->bank, 0, peak, ave);
}

int cutesdr_demodulator_process_stereo(cutesdr_demodulator* h, int n_in, const double* in, double* out)
{
    if (!h || !h->bank || n_in < 0 || (n_in && (!in || !out))) { set_error("demodulator_process_stereo: bad arguments / state"); return CUTESDR_E_STATE; }
    CSDR_TRY(cutesdr_bank_set_stereo(h->bank, 1));
    int L = 0;
    CSDR_TRY(cutesdr_bank_block_length(h->bank, &L));
    h->iq.resize((size_t)2 * n_in);
    for (int i = 0; i < 2 * n_in; i++) h->iq[i] = (float)in[i];
    const int cap = 2 * (n_in / L + 2) * kMaxBurstSamples;
    h->audio.resize(cap);
    int nout = 0;
    int rc = cutesdr_bank_process(h->bank, n_in, h->iq.data(), h->audio.data(), cap, &nout);
    if (rc < 0) return rc;
    for (int i = 0; i < 2 * nout; i++) out[i] = h->audio[i];
    return nout;
}

int cutesdr_demodulator_process(cutesdr_demodulator* h, int n_in, const double* in, double* out)
{
    if (!h || !h->bank || n_in < 0 || (n_in && (!in || !out))) { set_error("demodulator_process: bad arguments / state"); return CUTESDR_E_STATE; }
    int L = 0;
    CSDR_TRY(cutesdr_bank_block_length(h->bank, &L));
    CSDR_TRY(cutesdr_bank_set_stereo(h->bank, 0));
    h->iq.resize((size_t)2 * n_in);
    for (int i = 0; i < 2 * n_in; i++) h->iq[i] = (float)in[i];
    const int cap = (n_in / L + 2) * kMaxBurstSamples;
    h->audio.resize(cap);
    int nout = 0;
    int rc = cutesdr_bank_process(h->bank, n_in, h->iq.data(), h->audio.data(), cap, &nout);
    if (rc < 0) return rc;
    for (int i = 0; i < nout; i++) out[i] = h->audio[i];
    return nout;
}

}  // extern "C"
