// C-callable harness around the UNMODIFIED reference DSP classes (compiled from
// /root/reference/dsp/*.cpp by oracle/Makefile into oracle/_ref/*.so).
//
// TEST INFRASTRUCTURE ONLY. Used to (1) pin the oracle restatement
// (oracle/cutesdr_oracle.c), (2) generate tests/golden fixtures, (3) serve as
// the CPU reference arm of bench.py. Never linked into libcutesdr_cuda.
//
// Every reference object is placement-new'd into zeroed storage because
// CDemodulator leaves members uninitialised (dsp/demodulator.cpp:47-60) and
// DeleteAllDemods() deletes m_pFmDemod (:80-81).
#define private public   // harness peeks at a few members (m_InBufLimit, FFT average buffer)
#include "dsp/demodulator.h"
#include "dsp/noiseproc.h"
#include "dsp/fractresampler.h"
#include "dsp/fft.h"
#include "dsp/iir.h"
#undef private
#include "gui/testbench.h"
#include "interface/perform.h"

#include <stdlib.h>
#include <string.h>
#include <new>
#include <thread>
#include <vector>
#include <chrono>

thread_local CTestBench* g_pTestBench = NULL;
static thread_local CTestBench t_bench;

// interface/perform.h declares these; every call site in dsp/ is commented out
// but three files include the header.
void InitPerformance() {}
void StartPerformance() {}
void StopPerformance(int) {}
void ReadPerformance() {}
void SamplePerformance() {}
int GetDeltaPerformance() { return 0; }

static void ensure_bench() { if (!g_pTestBench) g_pTestBench = &t_bench; }

template <typename T> static T* znew() {
    ensure_bench();
    void* p = calloc(1, sizeof(T));
    return new (p) T();
}
template <typename T> static void zdelete(T* p) { if (p) { p->~T(); free(p); } }

static tDemodInfo make_info(const int* v) {
    // order: HiCut, HiCutmin, HiCutmax, LowCut, LowCutmin, LowCutmax, Offset,
    //        SquelchValue, AgcSlope, AgcThresh, AgcManualGain, AgcDecay, AgcOn, AgcHangOn
    tDemodInfo d;
    d.HiCut = v[0]; d.HiCutmin = v[1]; d.HiCutmax = v[2];
    d.LowCut = v[3]; d.LowCutmin = v[4]; d.LowCutmax = v[5];
    d.FilterClickResolution = 100;
    d.Offset = v[6]; d.SquelchValue = v[7]; d.AgcSlope = v[8]; d.AgcThresh = v[9];
    d.AgcManualGain = v[10]; d.AgcDecay = v[11]; d.AgcOn = v[12] != 0; d.AgcHangOn = v[13] != 0;
    d.Symetric = false;
    return d;
}

extern "C" {

int ref_max_decstages() { return MAX_DECSTAGES; }
int ref_max_inbufsize() { return MAX_INBUFSIZE; }

// ---------------- tap capture (PROFILE_1..7 of gui/testbench.cpp:71-81) ----------------
void ref_tap_enable(unsigned mask) { ensure_bench(); g_pTestBench->m_CaptureMask = mask; }
void ref_tap_clear() { ensure_bench(); for (int i = 0; i < NUM_PROFILES; i++) g_pTestBench->m_Tap[i].clear(); }
long ref_tap_size(int profile) { ensure_bench(); return (long)g_pTestBench->m_Tap[profile].size(); }
void ref_tap_read(int profile, double* out) {
    ensure_bench();
    std::vector<double>& v = g_pTestBench->m_Tap[profile];
    if (!v.empty()) memcpy(out, v.data(), v.size() * sizeof(double));
}

// ---------------- CDownConvert ----------------
void* ref_downconvert_new() { return znew<CDownConvert>(); }
void ref_downconvert_delete(void* h) { zdelete((CDownConvert*)h); }
void ref_downconvert_set_frequency(void* h, double f) { ((CDownConvert*)h)->SetFrequency(f); }
void ref_downconvert_set_cw_offset(void* h, double f) { ((CDownConvert*)h)->SetCwOffset(f); }
double ref_downconvert_set_data_rate(void* h, double r, double bw) { return ((CDownConvert*)h)->SetDataRate(r, bw); }
int ref_downconvert_process(void* h, int n, double* in, double* out) {
    return ((CDownConvert*)h)->ProcessData(n, (TYPECPX*)in, (TYPECPX*)out);
}
// stage list as tap counts: 3 = CIC3, 11 = fixed 11-tap, N = generic half-band
int ref_downconvert_stages(void* h, int* lens, int maxn) {
    CDownConvert* d = (CDownConvert*)h;
    int n = 0;
    for (int i = 0; i < MAX_DECSTAGES && d->m_pDecimatorPtrs[i]; i++) {
        CDownConvert::CDec2* p = d->m_pDecimatorPtrs[i];
        int len = 0;
        if (dynamic_cast<CDownConvert::CCicN3DecimateBy2*>(p)) len = 3;
        else if (dynamic_cast<CDownConvert::CHalfBand11TapDecimateBy2*>(p)) len = 11;
        else if (CDownConvert::CHalfBandDecimateBy2* hb = dynamic_cast<CDownConvert::CHalfBandDecimateBy2*>(p)) len = hb->m_FirLength;
        if (n < maxn) lens[n] = len;
        n++;
    }
    return n;
}

// ---------------- CFastFIR ----------------
void* ref_fastfir_new() { return znew<CFastFIR>(); }
void ref_fastfir_delete(void* h) { zdelete((CFastFIR*)h); }
void ref_fastfir_setup(void* h, double lo, double hi, double off, double rate) { ((CFastFIR*)h)->SetupParameters(lo, hi, off, rate); }
int ref_fastfir_process(void* h, int n, double* in, double* out) { return ((CFastFIR*)h)->ProcessData(n, (TYPECPX*)in, (TYPECPX*)out); }
void ref_fastfir_coef(void* h, double* out2048cpx) { memcpy(out2048cpx, ((CFastFIR*)h)->m_pFilterCoef, 2048 * sizeof(TYPECPX)); }

// ---------------- CFft ----------------
void* ref_fft_new() { return znew<CFft>(); }
void ref_fft_delete(void* h) { zdelete((CFft*)h); }
void ref_fft_set_params(void* h, int size, int invert, double dbcomp, double fs) { ((CFft*)h)->SetFFTParams(size, invert != 0, dbcomp, fs); }
void ref_fft_set_ave(void* h, int ave) { ((CFft*)h)->SetFFTAve(ave); }
void ref_fft_reset(void* h) { ((CFft*)h)->ResetFFT(); }
int ref_fft_put(void* h, int n, double* in) { return ((CFft*)h)->PutInDisplayFFT(n, (TYPECPX*)in); }
int ref_fft_get_screen(void* h, int maxh, int maxw, double maxdb, double mindb, int start, int stop, int* out) {
    return ((CFft*)h)->GetScreenIntegerFFTData(maxh, maxw, maxdb, mindb, start, stop, out) ? 1 : 0;
}
void ref_fft_fwd(void* h, double* io) { ((CFft*)h)->FwdFFT((TYPECPX*)io); }
void ref_fft_rev(void* h, double* io) { ((CFft*)h)->RevFFT((TYPECPX*)io); }
int ref_fft_size(void* h) { return ((CFft*)h)->m_FFTSize; }
void ref_fft_avebuf(void* h, double* out) { CFft* f = (CFft*)h; memcpy(out, f->m_pFFTAveBuf, f->m_FFTSize * sizeof(double)); }
void ref_fft_consts(void* h, double* kb_kc) { CFft* f = (CFft*)h; kb_kc[0] = f->m_K_B; kb_kc[1] = f->m_K_C; }
void ref_fft_bins(void* h, int* minmax) { CFft* f = (CFft*)h; minmax[0] = f->m_BinMin; minmax[1] = f->m_BinMax; }

// ---------------- CAgc / CSMeter ----------------
void* ref_agc_new() { return znew<CAgc>(); }
void ref_agc_delete(void* h) { zdelete((CAgc*)h); }
void ref_agc_set(void* h, int on, int hang, int thresh, int mgain, int slope, int decay, double rate) {
    ((CAgc*)h)->SetParameters(on != 0, hang != 0, thresh, mgain, slope, decay, rate);
}
void ref_agc_process(void* h, int n, double* in, double* out) { ((CAgc*)h)->ProcessData(n, (TYPECPX*)in, (TYPECPX*)out); }
void ref_agc_sizes(void* h, int* dw) { dw[0] = ((CAgc*)h)->m_DelaySamples; dw[1] = ((CAgc*)h)->m_WindowSamples; }

void* ref_smeter_new() { return znew<CSMeter>(); }
void ref_smeter_delete(void* h) { zdelete((CSMeter*)h); }
void ref_smeter_process(void* h, int n, double* in, double rate) { ((CSMeter*)h)->ProcessData(n, (TYPECPX*)in, rate); }
double ref_smeter_peak(void* h) { return ((CSMeter*)h)->GetPeak(); }
double ref_smeter_ave(void* h) { return ((CSMeter*)h)->GetAve(); }

// ---------------- demod objects ----------------
void* ref_am_new(double rate) { ensure_bench(); void* p = calloc(1, sizeof(CAmDemod)); return new (p) CAmDemod(rate); }
void ref_am_delete(void* h) { zdelete((CAmDemod*)h); }
void ref_am_set_bandwidth(void* h, double bw) { ((CAmDemod*)h)->SetBandwidth(bw); }
int ref_am_process(void* h, int n, double* in, double* out) { return ((CAmDemod*)h)->ProcessData(n, (TYPECPX*)in, (TYPEREAL*)out); }
int ref_am_process_stereo(void* h, int n, double* in, double* out) { return ((CAmDemod*)h)->ProcessData(n, (TYPECPX*)in, (TYPECPX*)out); }

void* ref_sam_new(double rate) { ensure_bench(); void* p = calloc(1, sizeof(CSamDemod)); return new (p) CSamDemod(rate); }
void ref_sam_delete(void* h) { zdelete((CSamDemod*)h); }
int ref_sam_process(void* h, int n, double* in, double* out) { return ((CSamDemod*)h)->ProcessData(n, (TYPECPX*)in, (TYPEREAL*)out); }
int ref_sam_process_stereo(void* h, int n, double* in, double* out) { return ((CSamDemod*)h)->ProcessData(n, (TYPECPX*)in, (TYPECPX*)out); }

void* ref_fm_new(double rate) { ensure_bench(); void* p = calloc(1, sizeof(CFmDemod)); return new (p) CFmDemod(rate); }
void ref_fm_delete(void* h) { zdelete((CFmDemod*)h); }
void ref_fm_set_squelch(void* h, int v) { ((CFmDemod*)h)->SetSquelch(v); }
int ref_fm_process(void* h, int n, double bw, double* in, double* out) { return ((CFmDemod*)h)->ProcessData(n, bw, (TYPECPX*)in, (TYPEREAL*)out); }
int ref_fm_process_stereo(void* h, int n, double bw, double* in, double* out) { return ((CFmDemod*)h)->ProcessData(n, bw, (TYPECPX*)in, (TYPECPX*)out); }

int ref_ssb_process(int n, double* in, double* out) { CSsbDemod d; return d.ProcessData(n, (TYPECPX*)in, (TYPEREAL*)out); }

// ---------------- CFir / CIir ----------------
void* ref_fir_new() { return znew<CFir>(); }
void ref_fir_delete(void* h) { zdelete((CFir*)h); }
int ref_fir_init_lp(void* h, double scale, double astop, double fpass, double fstop, double fs) { return ((CFir*)h)->InitLPFilter(scale, astop, fpass, fstop, fs); }
int ref_fir_init_hp(void* h, double scale, double astop, double fpass, double fstop, double fs) { return ((CFir*)h)->InitHPFilter(scale, astop, fpass, fstop, fs); }
void ref_fir_gen_hb(void* h, double off) { ((CFir*)h)->GenerateHBFilter(off); }
int ref_fir_taps(void* h, double* coef, double* icoef, double* qcoef) {
    CFir* f = (CFir*)h;
    for (int i = 0; i < f->m_NumTaps; i++) { if (coef) coef[i] = f->m_Coef[i]; if (icoef) icoef[i] = f->m_ICoef[i]; if (qcoef) qcoef[i] = f->m_QCoef[i]; }
    return f->m_NumTaps;
}
void ref_fir_process_real(void* h, int n, double* in, double* out) { ((CFir*)h)->ProcessFilter(n, (TYPEREAL*)in, (TYPEREAL*)out); }
void ref_fir_process_cpx(void* h, int n, double* in, double* out) { ((CFir*)h)->ProcessFilter(n, (TYPECPX*)in, (TYPECPX*)out); }

void* ref_iir_new() { return znew<CIir>(); }
void ref_iir_delete(void* h) { zdelete((CIir*)h); }
void ref_iir_init_lp(void* h, double f0, double q, double fs) { ((CIir*)h)->InitLP(f0, q, fs); }
void ref_iir_process_real(void* h, int n, double* in, double* out) { ((CIir*)h)->ProcessFilter(n, (TYPEREAL*)in, (TYPEREAL*)out); }

// ---------------- CFractResampler ----------------
void* ref_resampler_new(int maxin) { CFractResampler* r = znew<CFractResampler>(); r->Init(maxin); return r; }
void ref_resampler_delete(void* h) { zdelete((CFractResampler*)h); }
int ref_resampler_real(void* h, int n, double rate, double* in, double* out) { return ((CFractResampler*)h)->Resample(n, rate, (TYPEREAL*)in, (TYPEREAL*)out); }
int ref_resampler_cpx(void* h, int n, double rate, double* in, double* out) { return ((CFractResampler*)h)->Resample(n, rate, (TYPECPX*)in, (TYPECPX*)out); }
int ref_resampler_mono16(void* h, int n, double rate, double* in, short* out, double gain) { return ((CFractResampler*)h)->Resample(n, rate, (TYPEREAL*)in, (TYPEMONO16*)out, gain); }
int ref_resampler_stereo16(void* h, int n, double rate, double* in, short* out, double gain) { return ((CFractResampler*)h)->Resample(n, rate, (TYPECPX*)in, (TYPESTEREO16*)out, gain); }

// ---------------- CNoiseProc ----------------
void* ref_noiseproc_new() { return znew<CNoiseProc>(); }
void ref_noiseproc_delete(void* h) { zdelete((CNoiseProc*)h); }
void ref_noiseproc_setup(void* h, int on, double thr, double width, double fs) { ((CNoiseProc*)h)->SetupBlanker(on != 0, thr, width, fs); }
// fed in <=4096-sample slices (scratch m_TestBenchDataBuf[4096], dsp/noiseproc.cpp:64)
void ref_noiseproc_process(void* h, long n, double* io) {
    CNoiseProc* p = (CNoiseProc*)h;
    for (long i = 0; i < n; i += 4096) {
        int m = (int)((n - i) < 4096 ? (n - i) : 4096);
        p->ProcessBlanker(m, (TYPECPX*)io + i, (TYPECPX*)io + i);
    }
}

// ---------------- CDemodulator ----------------
void* ref_demod_new() { return znew<CDemodulator>(); }
void ref_demod_delete(void* h) { zdelete((CDemodulator*)h); }
void ref_demod_set_input_rate(void* h, double r) { ((CDemodulator*)h)->SetInputSampleRate(r); }
void ref_demod_set_demod(void* h, int mode, const int* info14) { ((CDemodulator*)h)->SetDemod(mode, make_info(info14)); }
void ref_demod_set_freq(void* h, double f) { ((CDemodulator*)h)->SetDemodFreq(f); }
double ref_demod_output_rate(void* h) { return ((CDemodulator*)h)->GetOutputRate(); }
double ref_demod_smeter_peak(void* h) { return ((CDemodulator*)h)->GetSMeterPeak(); }
double ref_demod_smeter_ave(void* h) { return ((CDemodulator*)h)->GetSMeterAve(); }
int ref_demod_inbuf_limit(void* h) { return ((CDemodulator*)h)->m_InBufLimit; }
// test aid: run the unmodified chain on another DSP block length (valid until the next SetDemod recomputes it)
void ref_demod_set_inbuf_limit(void* h, int limit) { ((CDemodulator*)h)->m_InBufLimit = limit; }

// Feed n complex samples in `packet`-sample calls (the app uses 256,
// interface/netiobase.cpp:593); mono audio appended to out. Returns count.
// `in` is complex64 (float pairs) widened to double here -- the same bits the
// GPU consumes -- or double pairs when in_is_double.
long ref_demod_run(void* h, long n, const void* in, int in_is_double, int packet, double* out, long out_cap, int stereo) {
    CDemodulator* d = (CDemodulator*)h;
    std::vector<TYPECPX> pkt(packet);
    std::vector<TYPECPX> obuf(16384);
    long nout = 0;
    for (long i = 0; i < n; i += packet) {
        int m = (int)((n - i) < packet ? (n - i) : packet);
        if (in_is_double) memcpy(pkt.data(), (const double*)in + 2 * i, m * sizeof(TYPECPX));
        else { const float* f = (const float*)in + 2 * i; for (int k = 0; k < m; k++) { pkt[k].re = f[2 * k]; pkt[k].im = f[2 * k + 1]; } }
        int r;
        if (stereo) r = d->ProcessData(m, pkt.data(), obuf.data());
        else r = d->ProcessData(m, pkt.data(), (TYPEREAL*)obuf.data());
        int w = stereo ? 2 * r : r;
        if (out && nout + w <= out_cap) memcpy(out + nout, obuf.data(), w * sizeof(double));
        nout += w;
    }
    return nout;
}

// ---------------- multi-threaded CPU baseline ----------------
// nch independent CDemodulator chains (+ optional resampler to 48 kHz) over the same
// complex64 stream, channels partitioned over nthreads std::threads. Returns seconds.
// modes[c], freqs[c] per channel; info14 shared per mode via infos[mode*14..].
double ref_bench_chains(int nch, const int* modes, const double* freqs, const int* infos,
                        double in_rate, long n, const float* iq, int nthreads, int resample48k,
                        double* checksum) {
    std::vector<double> sums(nch, 0.0);
    auto work = [&](int t) {
        ensure_bench();
        g_pTestBench->m_CaptureMask = 0;
        std::vector<TYPECPX> pkt(256);
        std::vector<double> obuf(16384), rbuf(32768);
        for (int c = t; c < nch; c += nthreads) {
            CDemodulator* d = znew<CDemodulator>();
            CFractResampler* rs = NULL;
            d->SetInputSampleRate(in_rate);
            d->SetDemod(modes[c], make_info(infos + 14 * modes[c]));
            d->SetDemodFreq(freqs[c]);
            if (resample48k) { rs = znew<CFractResampler>(); rs->Init(8192); }
            double orate = d->GetOutputRate();
            double acc = 0.0;
            for (long i = 0; i < n; i += 256) {
                int m = (int)((n - i) < 256 ? (n - i) : 256);
                const float* f = iq + 2 * i;
                for (int k = 0; k < m; k++) { pkt[k].re = f[2 * k]; pkt[k].im = f[2 * k + 1]; }
                int r = d->ProcessData(m, pkt.data(), (TYPEREAL*)obuf.data());
                if (r > 0 && rs) { r = rs->Resample(r, orate / 48000.0, obuf.data(), rbuf.data()); for (int k = 0; k < r; k++) acc += rbuf[k]; }
                else for (int k = 0; k < r; k++) acc += obuf[k];
            }
            sums[c] = acc;
            zdelete(d);
            if (rs) zdelete(rs);
        }
    };
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; t++) th.emplace_back(work, t);
    for (auto& x : th) x.join();
    auto t1 = std::chrono::steady_clock::now();
    if (checksum) { double s = 0; for (int c = 0; c < nch; c++) s += sums[c]; *checksum = s; }
    return std::chrono::duration<double>(t1 - t0).count();
}


// ---------------- persistent multi-channel chain set ----------------
// N independent CDemodulator chains (+ optional CFractResampler to `audio_rate`) that live across calls, so a
// benchmark can build them once (object construction, the 1025-tap design and CFractResampler::Init's 280 001
// sin/cos are NOT in the timed region) and a parity test can stream 0.2 s of signal through a subset of a big
// bank on all host cores and read every channel's audio back.
struct ChainSet {
    int nch;
    double in_rate, audio_rate;
    int keep;
    std::vector<CDemodulator*> d;
    std::vector<CFractResampler*> rs;
    std::vector<double> orate;
    std::vector<std::vector<double> > out;
    std::vector<double> sums;
};

void* ref_chains_new(int nch, const int* modes, const double* freqs, const int* infos14, double in_rate, double audio_rate, int keep_output) {
    ChainSet* s = new ChainSet();
    s->nch = nch; s->in_rate = in_rate; s->audio_rate = audio_rate; s->keep = keep_output;
    s->d.resize(nch); s->rs.assign(nch, (CFractResampler*)NULL); s->orate.resize(nch); s->out.resize(nch); s->sums.assign(nch, 0.0);
    for (int c = 0; c < nch; c++) {
        CDemodulator* d = znew<CDemodulator>();
        d->SetInputSampleRate(in_rate);
        d->SetDemod(modes[c], make_info(infos14 + 14 * c));
        d->SetDemodFreq(freqs[c]);
        s->d[c] = d;
        s->orate[c] = d->GetOutputRate();
        if (audio_rate > 0) { s->rs[c] = znew<CFractResampler>(); s->rs[c]->Init(8192); }
    }
    return s;
}
void ref_chains_delete(void* h) {
    ChainSet* s = (ChainSet*)h;
    if (!s) return;
    for (int c = 0; c < s->nch; c++) { zdelete(s->d[c]); if (s->rs[c]) zdelete(s->rs[c]); }
    delete s;
}
void ref_chains_set_freq(void* h, int c, double f) { ((ChainSet*)h)->d[c]->SetDemodFreq(f); }
void ref_chains_set_demod(void* h, int c, int mode, const int* info14) { ((ChainSet*)h)->d[c]->SetDemod(mode, make_info(info14)); }
void ref_chains_set_inbuf_limit(void* h, int limit) { ChainSet* s = (ChainSet*)h; for (int c = 0; c < s->nch; c++) s->d[c]->m_InBufLimit = limit; }
double ref_chains_smeter_ave(void* h, int c) { return ((ChainSet*)h)->d[c]->GetSMeterAve(); }
// Feed n complex64 samples to every chain in 256-sample packets (interface/netiobase.cpp:593), channels
// partitioned over nthreads std::threads. Returns the wall-clock seconds of the threaded section.
double ref_chains_run(void* h, long n, const float* iq, int nthreads) {
    ChainSet* s = (ChainSet*)h;
    if (nthreads < 1) nthreads = 1;
    auto work = [&](int t) {
        ensure_bench();
        g_pTestBench->m_CaptureMask = 0;
        std::vector<TYPECPX> pkt(256);
        std::vector<double> obuf(16384), rbuf(32768);
        for (int c = t; c < s->nch; c += nthreads) {
            CDemodulator* d = s->d[c];
            CFractResampler* rs = s->rs[c];
            double acc = 0.0;
            for (long i = 0; i < n; i += 256) {
                int m = (int)((n - i) < 256 ? (n - i) : 256);
                const float* f = iq + 2 * i;
                for (int k = 0; k < m; k++) { pkt[k].re = f[2 * k]; pkt[k].im = f[2 * k + 1]; }
                int r = d->ProcessData(m, pkt.data(), (TYPEREAL*)obuf.data());
                const double* o = obuf.data();
                if (r > 0 && rs) { r = rs->Resample(r, s->orate[c] / s->audio_rate, obuf.data(), rbuf.data()); o = rbuf.data(); }
                for (int k = 0; k < r; k++) acc += o[k];
                if (s->keep && r > 0) s->out[c].insert(s->out[c].end(), o, o + r);
            }
            s->sums[c] += acc;
        }
    };
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; t++) th.emplace_back(work, t);
    for (auto& x : th) x.join();
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}
long ref_chains_out_size(void* h, int c) { return (long)((ChainSet*)h)->out[c].size(); }
void ref_chains_out_read(void* h, int c, double* out, int clear) {
    std::vector<double>& v = ((ChainSet*)h)->out[c];
    if (!v.empty()) memcpy(out, v.data(), v.size() * sizeof(double));
    if (clear) std::vector<double>().swap(v);
}
double ref_chains_checksum(void* h) { ChainSet* s = (ChainSet*)h; double a = 0; for (int c = 0; c < s->nch; c++) a += s->sums[c]; return a; }

// noise blanker over a complex64 stream (shared wideband pre-processing, interface/sdrinterface.cpp:884): in place,
// <= 4096 samples per call. Returns seconds.
double ref_noiseproc_process_f32(void* h, long n, float* io) {
    CNoiseProc* p = (CNoiseProc*)h;
    std::vector<TYPECPX> buf(4096);
    auto t0 = std::chrono::steady_clock::now();
    for (long i = 0; i < n; i += 4096) {
        int m = (int)((n - i) < 4096 ? (n - i) : 4096);
        float* f = io + 2 * i;
        for (int k = 0; k < m; k++) { buf[k].re = f[2 * k]; buf[k].im = f[2 * k + 1]; }
        p->ProcessBlanker(m, buf.data(), buf.data());
        for (int k = 0; k < m; k++) { f[2 * k] = (float)buf[k].re; f[2 * k + 1] = (float)buf[k].im; }
    }
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

} // extern "C"
