// iir.cu -- CIir (dsp/iir.h:16-40, dsp/iir.cpp:77-201): one biquad in direct form 2, real or complex stream.
// The reference object is embedded in CSdrInterface (interface/sdrinterface.h:178) and in CFmDemod; the FM
// demodulator's copy runs fused inside kernel group 4 (post.cu), this handle is the stand-alone class.
// The recurrence is sequential: one thread per real stream (a complex stream is two independent ones), double
// precision with explicitly rounded multiplies and adds (no FMA contraction), so the output is bit-identical to
// the reference's.
#include "common.cuh"

namespace csdr {

struct IirCoef { double a1, a2, b0, b1, b2; };

// x, y: [n][width] interleaved doubles (width 1 = real, 2 = complex); w: [2][2] delay storage (w1, w2 per stream)
__global__ void k_iir(const double* __restrict__ x, double* __restrict__ y, int n, int width, IirCoef c, double* __restrict__ w)
{
    const int s = threadIdx.x;
    if (s >= width) return;
    double w1 = w[2 * s], w2 = w[2 * s + 1];
    for (int i = 0; i < n; i++) {
        // w0 = in - A1*w1 - A2*w2;  out = B0*w0 + B1*w1 + B2*w2      (dsp/iir.cpp:171-180)
        const double w0 = __dsub_rn(__dsub_rn(x[(size_t)i * width + s], __dmul_rn(c.a1, w1)), __dmul_rn(c.a2, w2));
        y[(size_t)i * width + s] = __dadd_rn(__dadd_rn(__dmul_rn(c.b0, w0), __dmul_rn(c.b1, w1)), __dmul_rn(c.b2, w2));
        w2 = w1;
        w1 = w0;
    }
    w[2 * s] = w1;
    w[2 * s + 1] = w2;
}

}  // namespace csdr

using namespace csdr;

struct cutesdr_iir {
    int device = 0;
    cudaStream_t st = 0;
    std::mutex mu;
    IirCoef c{};
    double* d_w = nullptr;
    double* d_buf = nullptr;
    size_t cap = 0;
    long long launches = 0;
    ~cutesdr_iir()
    {
        if (st) cudaStreamSynchronize(st);
        cudaFree(d_w);
        cudaFree(d_buf);
        if (st) cudaStreamDestroy(st);
    }
};

extern "C" {

int cutesdr_iir_init(cutesdr_iir* h, int kind, double f0, double q, double sample_rate);

int cutesdr_iir_create(cutesdr_iir** out, int device)
{
    if (!out) { set_error("iir_create: bad arguments"); return CUTESDR_E_ARG; }
    *out = nullptr;
    CSDR_CK(cudaSetDevice(device));
    std::unique_ptr<cutesdr_iir> h(new cutesdr_iir());
    h->device = device;
    CSDR_CK(cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking));
    CSDR_CK(cudaMalloc(&h->d_w, 4 * sizeof(double)));
    CSDR_TRY(cutesdr_iir_init(h.get(), CUTESDR_IIR_BR, 25000, 1000.0, 100000));      // CIir::CIir(), dsp/iir.cpp:77-80
    *out = h.release();
    return CUTESDR_OK;
}

void cutesdr_iir_destroy(cutesdr_iir* h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    delete h;
}

int cutesdr_iir_init(cutesdr_iir* h, int kind, double f0, double q, double sample_rate)
{
    if (!h || kind < CUTESDR_IIR_LP || kind > CUTESDR_IIR_BR) { set_error("iir_init: bad arguments"); return CUTESDR_E_ARG; }
    std::lock_guard<std::mutex> lk(h->mu);
    CSDR_CK(cudaSetDevice(h->device));
    // RBJ biquads scaled by 1/a0, dsp/iir.cpp:86-163
    const double w0 = kTwoPi * f0 / sample_rate;
    const double alpha = sin(w0) / (2.0 * q);
    const double A = 1.0 / (1.0 + alpha);
    IirCoef& c = h->c;
    switch (kind) {
    case CUTESDR_IIR_LP: c.b0 = A * ((1.0 - cos(w0)) / 2.0); c.b1 = A * (1.0 - cos(w0)); c.b2 = A * ((1.0 - cos(w0)) / 2.0); break;
    case CUTESDR_IIR_HP: c.b0 = A * ((1.0 + cos(w0)) / 2.0); c.b1 = -A * (1.0 + cos(w0)); c.b2 = A * ((1.0 + cos(w0)) / 2.0); break;
    case CUTESDR_IIR_BP: c.b0 = A * alpha; c.b1 = 0.0; c.b2 = A * -alpha; break;
    default: c.b0 = A * 1.0; c.b1 = A * (-2.0 * cos(w0)); c.b2 = A * 1.0; break;
    }
    c.a1 = A * (-2.0 * cos(w0));
    c.a2 = A * (1.0 - alpha);
    CSDR_CK(cudaMemsetAsync(h->d_w, 0, 4 * sizeof(double), h->st));       // every Init* clears the delay storage
    return CUTESDR_OK;
}

static int iir_run(cutesdr_iir* h, int n, int width, const double* in, double* out)
{
    if (!h || n < 0 || (n > 0 && (!in || !out))) { set_error("iir_process: bad arguments"); return CUTESDR_E_ARG; }
    if (n == 0) return 0;
    std::lock_guard<std::mutex> lk(h->mu);
    CSDR_CK(cudaSetDevice(h->device));
    const size_t bytes = (size_t)n * width * sizeof(double);
    if (2 * bytes > h->cap) {
        CSDR_CK(cudaStreamSynchronize(h->st));
        cudaFree(h->d_buf);
        h->d_buf = nullptr;
        h->cap = 0;
        CSDR_CK(cudaMalloc(&h->d_buf, 2 * bytes));
        h->cap = 2 * bytes;
    }
    double* d_in = h->d_buf;
    double* d_out = h->d_buf + (size_t)n * width;
    CSDR_CK(cudaMemcpyAsync(d_in, in, bytes, cudaMemcpyHostToDevice, h->st));
    k_iir<<<1, 32, 0, h->st>>>(d_in, d_out, n, width, h->c, h->d_w);
    h->launches++;
    CSDR_CK(cudaGetLastError());
    CSDR_CK(cudaMemcpyAsync(out, d_out, bytes, cudaMemcpyDeviceToHost, h->st));
    CSDR_CK(cudaStreamSynchronize(h->st));
    return n;
}

int cutesdr_iir_process_real(cutesdr_iir* h, int n, const double* in, double* out) { return iir_run(h, n, 1, in, out); }
int cutesdr_iir_process_cpx(cutesdr_iir* h, int n, const double* in, double* out) { return iir_run(h, n, 2, in, out); }

}  // extern "C"
