// post.cu -- kernel 4: S-meter, AGC and demodulator recursions, one THREAD per channel.
//
// These stages are recurrences in time (AGC averagers with data-dependent rise/fall constants,
// the hang timer, second-order PLLs through atan2/sin/cos, DC trackers, a biquad), so time is
// sequential per channel and channels are the parallel axis: lane = channel, all streams
// time-major [t][stride] so a warp touches one contiguous row per step. Recurrent state is kept
// in double precision -- the averagers have time constants of 1e4 samples and float32 rounding
// in the recursion would sit near -85 dB -- while the burst data itself is float32.
#include "post.cuh"

namespace csdr {

// ------------------------------------------------------------------------------------------
// host: Kaiser FIR designs (dsp/fir.cpp:173-367, 414-432)
// ------------------------------------------------------------------------------------------
static double bessel_i0(double x)
{
    double x2 = x / 2.0, sum = 1.0, ds = 1.0, di = 1.0, t;
    do {
        t = x2 / di;
        t *= t;
        ds *= t;
        sum += ds;
        di += 1.0;
    } while (ds >= 1e-9 * sum);
    return sum;
}

static double kaiser_beta(double astop)
{
    if (astop < 20.96) return 0;
    if (astop >= 50.0) return .1102 * (astop - 8.71);
    return .5842 * pow((astop - 20.96), 0.4) + .07886 * (astop - 20.96);
}

int design_kaiser_lp(double scale, double astop, double fpass, double fstop, double fs, double* coef)
{
    const double nfp = fpass / fs, nfs = fstop / fs, nfc = (nfs + nfp) / 2.0;
    const double beta = kaiser_beta(astop);
    int ntaps = (astop - 8.0) / (2.285 * kTwoPi * (nfs - nfp)) + 1;
    if (ntaps > kFirMax) ntaps = kFirMax;
    if (ntaps < 3) ntaps = 3;
    const double centre = .5 * (double)(ntaps - 1);
    const double izb = bessel_i0(beta);
    for (int n = 0; n < ntaps; n++) {
        double x = (double)n - centre, c;
        if ((double)n == centre) c = 2.0 * nfc;
        else c = sin(kTwoPi * x * nfc) / (kPi * x);
        x = ((double)n - ((double)ntaps - 1.0) / 2.0) / (((double)ntaps - 1.0) / 2.0);
        coef[n] = scale * c * bessel_i0(beta * sqrt(1 - (x * x))) / izb;
    }
    return ntaps;
}

int design_kaiser_hp(double scale, double astop, double fpass, double fstop, double fs, double* coef)
{
    const double nfp = fpass / fs, nfs = fstop / fs, nfc = (nfs + nfp) / 2.0;
    const double beta = kaiser_beta(astop);
    int ntaps = (astop - 8.0) / (2.285 * kTwoPi * (nfp - nfs)) + 1;
    if (ntaps > (kFirMax - 1)) ntaps = kFirMax - 1;
    if (ntaps < 3) ntaps = 3;
    ntaps |= 1;
    const double izb = bessel_i0(beta);
    const double centre = .5 * (double)(ntaps - 1);
    for (int n = 0; n < ntaps; n++) {
        double x = (double)n - (double)(ntaps - 1) / 2.0, c;
        if ((double)n == centre) c = 1.0 - 2.0 * nfc;
        else c = (sin(kPi * x) / (kPi * x) - sin(kTwoPi * x * nfc) / (kPi * x));
        x = ((double)n - ((double)ntaps - 1.0) / 2.0) / (((double)ntaps - 1.0) / 2.0);
        coef[n] = scale * c * bessel_i0(beta * sqrt(1 - (x * x))) / izb;
    }
    return ntaps;
}

// ------------------------------------------------------------------------------------------
// device state layout (struct of arrays, [field][stride])
// ------------------------------------------------------------------------------------------
enum { P_AGC_ON, P_AGC_HANG, P_KNEE, P_GAIN_SLOPE, P_FIXED_GAIN, P_MANUAL_GAIN, P_A_RISE, P_A_FALL, P_D_RISE,
       P_D_FALL, P_HANG_TIME, P_SQ_THRESH, P_NTAPS, P_COUNT };
enum { S_SM_ATT, S_SM_DEC, S_SM_AVE, S_SM_PEAK, S_AGC_PEAK, S_AGC_ATT, S_AGC_DEC, S_Z1, S_PHASE, S_FREQ, S_FM_DC,
       S_SQ_AVE, S_LP_W1, S_LP_W2, S_COUNT };
enum { I_AGC_DPTR, I_AGC_MPOS, I_AGC_HANGT, I_SQUELCHED, I_COUNT };
enum { R_AGC = 1, R_DEMOD = 2, R_FIR = 4, R_SMETER = 8 };

constexpr int kHist = kFirMax - 1;

__global__ void __launch_bounds__(64) k_post(const float2* __restrict__ y, int n, int nch, int stride, PostUniform u,
                                             const double* __restrict__ par, const double* __restrict__ taps,
                                             const int* __restrict__ mode_arr, int* __restrict__ reset_arr,
                                             double* __restrict__ state, int* __restrict__ istate,
                                             float2* __restrict__ agc_delay, double* __restrict__ agc_mag,
                                             double* __restrict__ v, float* __restrict__ audio, int audio_stride,
                                             int audio_off, const int* __restrict__ chan_map, float2* __restrict__ tap3)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nch) return;
#define PAR(f) par[(size_t)(f) * stride + c]
#define ST(f) state[(size_t)(f) * stride + c]
#define IST(f) istate[(size_t)(f) * stride + c]
    const int mode = mode_arr[c];
    const int rf = reset_arr[c];
    if (rf) {
        if (rf & R_SMETER) { ST(S_SM_ATT) = -120.0; ST(S_SM_DEC) = -120.0; ST(S_SM_AVE) = 0.0; ST(S_SM_PEAK) = 0.0; }
        if (rf & R_AGC) {      // dsp/agc.cpp:121-136
            for (int i = 0; i < kAgcBuf; i++) {
                agc_delay[(size_t)i * stride + c] = make_float2(0.f, 0.f);
                agc_mag[(size_t)i * stride + c] = -16.0;
            }
            IST(I_AGC_DPTR) = 0; IST(I_AGC_HANGT) = 0; IST(I_AGC_MPOS) = 0;
            ST(S_AGC_PEAK) = -16.0; ST(S_AGC_DEC) = -5.0; ST(S_AGC_ATT) = -5.0;
        }
        if (rf & R_DEMOD) {
            ST(S_Z1) = 0.0; ST(S_PHASE) = 0.0; ST(S_FREQ) = 0.0; ST(S_FM_DC) = 0.0; ST(S_SQ_AVE) = 0.0;
            ST(S_LP_W1) = 0.0; ST(S_LP_W2) = 0.0;
            IST(I_SQUELCHED) = 1;
        }
        if (rf & R_FIR) for (int i = 0; i < kHist; i++) v[(size_t)i * stride + c] = 0.0;
        reset_arr[c] = 0;
    }

    // ---- parameters
    const bool agc_on = PAR(P_AGC_ON) != 0.0, use_hang = PAR(P_AGC_HANG) != 0.0;
    const double knee = PAR(P_KNEE), gain_slope = PAR(P_GAIN_SLOPE), fixed_gain = PAR(P_FIXED_GAIN);
    const double manual_gain = PAR(P_MANUAL_GAIN);
    const double a_rise = PAR(P_A_RISE), a_fall = PAR(P_A_FALL), d_rise = PAR(P_D_RISE), d_fall = PAR(P_D_FALL);
    const int hang_time = (int)PAR(P_HANG_TIME);
    const int ntaps = (int)PAR(P_NTAPS);

    // ---- state
    double sm_att = ST(S_SM_ATT), sm_dec = ST(S_SM_DEC), sm_ave = ST(S_SM_AVE), sm_peak = ST(S_SM_PEAK);
    double peak = ST(S_AGC_PEAK), att = ST(S_AGC_ATT), dec = ST(S_AGC_DEC);
    int dptr = IST(I_AGC_DPTR), mpos = IST(I_AGC_MPOS), hang_timer = IST(I_AGC_HANGT);
    double z1 = ST(S_Z1), phase = ST(S_PHASE), freq = ST(S_FREQ), fm_dc = ST(S_FM_DC);

    const double pll_alpha = mode == POST_FM ? u.fm_alpha : u.sam_alpha;
    const double pll_beta = mode == POST_FM ? u.fm_beta : u.sam_beta;
    const double pll_lo = mode == POST_FM ? u.fm_lo : u.sam_lo;
    const double pll_hi = mode == POST_FM ? u.fm_hi : u.sam_hi;
    float* aout = audio ? audio + (size_t)chan_map[c] * audio_stride + audio_off : nullptr;

    for (int i = 0; i < n; i++) {
        const float2 xf = y[(size_t)i * stride + c];
        const double xr = xf.x, xi = xf.y;
        // ---- CSMeter::ProcessData, dsp/smeter.cpp:71-92
        if (mode != POST_AGC_ONLY) {
            const double mag = 10.0 * log10((xr * xr + xi * xi) / (32767.0 * 32767.0) + 1e-50);
            sm_att = (1.0 - u.sm_attack) * sm_att + u.sm_attack * mag;
            sm_dec = (1.0 - u.sm_decay) * sm_dec + u.sm_decay * mag;
            if (sm_att > sm_dec) { sm_ave = sm_att; sm_dec = sm_att; }
            else sm_ave = sm_dec;
            if (mag > sm_peak) sm_peak = mag;
        }
        // ---- CAgc::ProcessData, dsp/agc.cpp:174-296
        double outr, outi;
        if (agc_on) {
            const float2 dl = agc_delay[(size_t)dptr * stride + c];
            agc_delay[(size_t)dptr * stride + c] = xf;
            if (++dptr >= u.agc_delay) dptr = 0;
            double mag = fabs(xr);
            const double mim = fabs(xi);
            if (mim > mag) mag = mim;
            mag = log10(mag + 3.2767e-4) - log10(32767.0);
            const double oldest = agc_mag[(size_t)mpos * stride + c];
            agc_mag[(size_t)mpos * stride + c] = mag;
            if (++mpos >= u.agc_window) mpos = 0;
            if (mag > peak) peak = mag;
            else if (oldest == peak) {
                peak = -8.0;
                for (int k = 0; k < u.agc_window; k++) {
                    const double t = agc_mag[(size_t)k * stride + c];
                    if (t > peak) peak = t;
                }
            }
            if (peak > att) att = (1.0 - a_rise) * att + a_rise * peak;
            else att = (1.0 - a_fall) * att + a_fall * peak;
            if (use_hang) {
                if (peak > dec) { dec = (1.0 - d_rise) * dec + d_rise * peak; hang_timer = 0; }
                else if (hang_timer < hang_time) hang_timer++;
                else dec = (1.0 - d_fall) * dec + d_fall * peak;
            } else {
                if (peak > dec) dec = (1.0 - d_rise) * dec + d_rise * peak;
                else dec = (1.0 - d_fall) * dec + d_fall * peak;
            }
            const double m = att > dec ? att : dec;
            const double gain = (m <= knee) ? fixed_gain : 0.7 * pow(10.0, m * (gain_slope - 1.0));
            outr = (double)dl.x * gain;
            outi = (double)dl.y * gain;
        } else {
            outr = manual_gain * xr;
            outi = manual_gain * xi;
        }
        if (tap3) tap3[(size_t)i * stride + c] = make_float2((float)outr, (float)outi);

        // ---- demodulators (first, sample-recursive part)
        if (mode == POST_AM) {
            // dsp/amdemod.cpp:68-78: envelope then DC-removal IIR
            const double mag = sqrt(outr * outr + outi * outi);
            const double z0 = mag + (z1 * 0.99);
            v[(size_t)(kHist + i) * stride + c] = z0 - z1;
            z1 = z0;
        } else if (mode == POST_SAM || mode == POST_FM) {
            // second-order PLL, dsp/samdemod.cpp:81-105 / dsp/fmdemod.cpp:166-187. The SAM mono
            // path mixes with (cos, -sin) and uses +atan2; FM mixes with (cos, sin) and uses -atan2.
            double sn, cs;
            sincos(phase, &sn, &cs);
            if (mode == POST_SAM) sn = -sn;
            const double tr = cs * outr - sn * outi;
            const double ti = cs * outi + sn * outr;
            double err = atan2(ti, tr);
            if (mode == POST_FM) err = -err;
            freq += (pll_beta * err);
            if (freq > pll_hi) freq = pll_hi;
            else if (freq < pll_lo) freq = pll_lo;
            phase += (freq + pll_alpha * err);
            // the reference wraps once per call (fmod after the loop); wrapping every sample is
            // the same angle and keeps sincos in its accurate range
            if (phase > kTwoPi) phase -= kTwoPi;
            else if (phase < -kTwoPi) phase += kTwoPi;
            if (mode == POST_SAM) {
                const double z0 = tr + (z1 * 0.99);
                if (aout) aout[i] = (float)(z0 - z1);
                z1 = z0;
            } else {
                fm_dc = (1.0 - u.fm_dc_alpha) * fm_dc + u.fm_dc_alpha * freq;
                v[(size_t)(kHist + i) * stride + c] = (freq - fm_dc) * u.fm_gain;
            }
        } else if (mode == POST_SSB) {
            if (aout) aout[i] = (float)outr;       // dsp/ssbdemod.cpp:48-53
        }
    }

    // ---- second part: feed-forward FIRs and the squelch decision
    if (mode == POST_AM) {
        // Kaiser low-pass, dsp/amdemod.cpp:80 -> CFir::ProcessFilter, dsp/fir.cpp:72-91
        for (int i = 0; i < n; i++) {
            double acc = 0.0;
            for (int k = 0; k < ntaps; k++) acc += taps[(size_t)k * stride + c] * v[(size_t)(kHist + i - k) * stride + c];
            if (aout) aout[i] = (float)acc;
        }
    } else if (mode == POST_FM) {
        // PerformNoiseSquelch, dsp/fmdemod.cpp:113-152: evaluated once per burst
        double sq_ave = ST(S_SQ_AVE);
        int squelched = IST(I_SQUELCHED);
        const double sq_thresh = PAR(P_SQ_THRESH);
        for (int i = 0; i < n; i++) {
            double acc = 0.0;
            for (int k = 0; k < ntaps; k++) acc += taps[(size_t)k * stride + c] * v[(size_t)(kHist + i - k) * stride + c];
            sq_ave = (1.0 - u.fm_sq_alpha) * sq_ave + u.fm_sq_alpha * fabs(acc);
        }
        if (0 == sq_thresh) squelched = 1;
        else if (squelched) { if (sq_ave < (sq_thresh - 100.0)) squelched = 0; }
        else { if (sq_ave >= (sq_thresh + 100.0)) squelched = 1; }
        if (squelched) {
            if (aout) for (int i = 0; i < n; i++) aout[i] = 0.f;
        } else {
            double w1 = ST(S_LP_W1), w2 = ST(S_LP_W2);
            for (int i = 0; i < n; i++) {       // CIir::ProcessFilter, dsp/iir.cpp:171-180
                const double w0 = v[(size_t)(kHist + i) * stride + c] - u.lp_a1 * w1 - u.lp_a2 * w2;
                if (aout) aout[i] = (float)(u.lp_b0 * w0 + u.lp_b1 * w1 + u.lp_b2 * w2);
                w2 = w1;
                w1 = w0;
            }
            ST(S_LP_W1) = w1; ST(S_LP_W2) = w2;
        }
        ST(S_SQ_AVE) = sq_ave;
        IST(I_SQUELCHED) = squelched;
    }
    if (mode == POST_AM || mode == POST_FM) {
        // keep the last kHist inputs of the FIR for the next burst
        for (int i = 0; i < kHist; i++) v[(size_t)i * stride + c] = v[(size_t)(n + i) * stride + c];
    }

    ST(S_SM_ATT) = sm_att; ST(S_SM_DEC) = sm_dec; ST(S_SM_AVE) = sm_ave; ST(S_SM_PEAK) = sm_peak;
    ST(S_AGC_PEAK) = peak; ST(S_AGC_ATT) = att; ST(S_AGC_DEC) = dec;
    IST(I_AGC_DPTR) = dptr; IST(I_AGC_MPOS) = mpos; IST(I_AGC_HANGT) = hang_timer;
    ST(S_Z1) = z1; ST(S_PHASE) = phase; ST(S_FREQ) = freq; ST(S_FM_DC) = fm_dc;
#undef PAR
#undef ST
#undef IST
}

// ------------------------------------------------------------------------------------------
// PostBank
// ------------------------------------------------------------------------------------------
PostBank::~PostBank()
{
    cudaFree(d_par_); cudaFree(d_taps_); cudaFree(d_mode_); cudaFree(d_reset_); cudaFree(d_state_);
    cudaFree(d_istate_); cudaFree(d_agc_delay_); cudaFree(d_agc_mag_); cudaFree(d_v_);
}

int PostBank::init(int nch, int stride, double rate, int max_samples, cudaStream_t st, LaunchCounter* lc)
{
    nch_ = nch; stride_ = stride; rate_ = rate; st_ = st; lc_ = lc;
    max_n_ = max_samples;
    // CSMeter, dsp/smeter.cpp:66-69
    uni_.sm_attack = (1.0 - exp(-1.0 / (rate * .01)));
    uni_.sm_decay = (1.0 - exp(-1.0 / (rate * .5)));
    // CAgc, dsp/agc.cpp:157-164
    uni_.agc_delay = (int)(rate * .015);
    uni_.agc_window = (int)(rate * .018);
    if (uni_.agc_delay >= kAgcBuf - 1) uni_.agc_delay = kAgcBuf - 1;
    if (uni_.agc_window > kAgcBuf) { set_error("AGC window %d exceeds MAX_DELAY_BUF at %g Hz", uni_.agc_window, rate); return CUTESDR_E_ARG; }
    if (uni_.agc_delay < 1 || uni_.agc_window < 1) { set_error("sample rate %g too low for the AGC", rate); return CUTESDR_E_ARG; }
    const double norm = kTwoPi / rate;
    // CSamDemod ctor, dsp/samdemod.cpp:59-65
    uni_.sam_lo = -1000.0 * norm; uni_.sam_hi = 1000.0 * norm;
    uni_.sam_alpha = 2.0 * .707 * 100.0 * norm;
    uni_.sam_beta = (uni_.sam_alpha * uni_.sam_alpha) / (4.0 * .707 * .707);
    // CFmDemod ctor, dsp/fmdemod.cpp:68-86
    uni_.fm_lo = -6000.0 * norm; uni_.fm_hi = 6000.0 * norm;
    uni_.fm_alpha = 2.0 * .707 * 3000.0 * 2.0 * norm;
    uni_.fm_beta = (uni_.fm_alpha * uni_.fm_alpha) / (4.0 * .707 * .707);
    uni_.fm_gain = 25000.0 / uni_.fm_hi;
    uni_.fm_dc_alpha = (1.0 - exp(-1.0 / (rate * 0.01)));
    uni_.fm_sq_alpha = (1.0 - exp(-1.0 / (rate * .02)));
    {   // CIir::InitLP(3000, 1.0, rate), dsp/iir.cpp:86-101
        const double w0 = kTwoPi * 3000.0 / rate, alpha = sin(w0) / 2.0, A = 1.0 / (1.0 + alpha);
        uni_.lp_b0 = A * ((1.0 - cos(w0)) / 2.0);
        uni_.lp_b1 = A * (1.0 - cos(w0));
        uni_.lp_b2 = A * ((1.0 - cos(w0)) / 2.0);
        uni_.lp_a1 = A * (-2.0 * cos(w0));
        uni_.lp_a2 = A * (1.0 - alpha);
    }
    agc_.assign(nch, AgcHost());
    fm_bw_.assign(nch, 3000.0);
    h_par_.assign((size_t)P_COUNT * stride, 0.0);
    h_taps_.assign((size_t)kFirMax * stride, 0.0);
    h_mode_.assign(stride, POST_NONE);
    h_reset_.assign(stride, R_AGC | R_DEMOD | R_FIR | R_SMETER);
    for (int i = 0; i < nch; i++) {
        h_par_[(size_t)P_AGC_ON * stride + i] = 1.0;
        h_par_[(size_t)P_NTAPS * stride + i] = 1.0;
    }
    CSDR_CK(cudaMalloc(&d_par_, h_par_.size() * sizeof(double)));
    CSDR_CK(cudaMalloc(&d_taps_, h_taps_.size() * sizeof(double)));
    CSDR_CK(cudaMalloc(&d_mode_, stride * sizeof(int)));
    CSDR_CK(cudaMalloc(&d_reset_, stride * sizeof(int)));
    CSDR_CK(cudaMalloc(&d_state_, (size_t)S_COUNT * stride * sizeof(double)));
    CSDR_CK(cudaMalloc(&d_istate_, (size_t)I_COUNT * stride * sizeof(int)));
    CSDR_CK(cudaMalloc(&d_agc_delay_, (size_t)kAgcBuf * stride * sizeof(float2)));
    CSDR_CK(cudaMalloc(&d_agc_mag_, (size_t)kAgcBuf * stride * sizeof(double)));
    CSDR_CK(cudaMalloc(&d_v_, (size_t)(kHist + max_n_) * stride * sizeof(double)));
    CSDR_CK(cudaMemsetAsync(d_state_, 0, (size_t)S_COUNT * stride * sizeof(double), st_));
    CSDR_CK(cudaMemsetAsync(d_istate_, 0, (size_t)I_COUNT * stride * sizeof(int), st_));
    CSDR_CK(cudaMemsetAsync(d_v_, 0, (size_t)(kHist + max_n_) * stride * sizeof(double), st_));
    dirty_ = true;
    return CUTESDR_OK;
}

void PostBank::set_mode(int i, int mode)
{
    int m;
    switch (mode) {
    case CUTESDR_DEMOD_AM: m = POST_AM; break;
    case CUTESDR_DEMOD_SAM: m = POST_SAM; break;
    case CUTESDR_DEMOD_FM: m = POST_FM; break;
    case CUTESDR_DEMOD_USB: case CUTESDR_DEMOD_LSB: case CUTESDR_DEMOD_CWU: case CUTESDR_DEMOD_CWL: m = POST_SSB; break;
    case POST_AGC_ONLY: m = POST_AGC_ONLY; break;
    default: m = POST_NONE; break;
    }
    h_mode_[i] = m;
    h_reset_[i] |= R_DEMOD | R_FIR;
    if (m == POST_FM) {   // CFmDemod ctor designs its squelch high-pass for 3 kHz (dsp/fmdemod.cpp:79,88)
        fm_bw_[i] = -1.0;
    }
    dirty_ = true;
}

void PostBank::set_agc(int i, int on, int hang, int thresh, int manual_gain, int slope, int decay)
{
    // CAgc::SetParameters, dsp/agc.cpp:104-167 (the rate never changes inside a group)
    AgcHost& a = agc_[i];
    if (a.valid && on == a.on && hang == a.hang && thresh == a.thresh && manual_gain == a.mgain && slope == a.slope &&
        decay == a.decay)
        return;
    a.valid = true; a.on = on; a.hang = hang; a.thresh = thresh; a.mgain = manual_gain; a.slope = slope; a.decay = decay;
    auto P = [&](int f) -> double& { return h_par_[(size_t)f * stride_ + i]; };
    P(P_AGC_ON) = on ? 1.0 : 0.0;
    P(P_AGC_HANG) = hang ? 1.0 : 0.0;
    P(P_MANUAL_GAIN) = 32767.0 * pow(10.0, -(100 - (double)manual_gain) / 20.0);
    const double knee = (double)thresh / 20.0;
    const double gain_slope = a.slope / (100.0);
    P(P_KNEE) = knee;
    P(P_GAIN_SLOPE) = gain_slope;
    P(P_FIXED_GAIN) = 0.7 * pow(10.0, knee * (gain_slope - 1.0));
    P(P_A_RISE) = (1.0 - exp(-1.0 / (rate_ * .002)));
    P(P_A_FALL) = (1.0 - exp(-1.0 / (rate_ * .005)));
    P(P_D_RISE) = (1.0 - exp(-1.0 / (rate_ * (double)decay * .001 * .3)));
    P(P_HANG_TIME) = (double)(int)(rate_ * (double)decay * .001);
    if (hang) P(P_D_FALL) = (1.0 - exp(-1.0 / (rate_ * .05)));
    else P(P_D_FALL) = (1.0 - exp(-1.0 / (rate_ * (double)decay * .001)));
    dirty_ = true;
}

void PostBank::set_am_bandwidth(int i, double bw)
{
    double coef[kFirMax];
    int n = design_kaiser_lp(1.0, 50.0, bw, bw * 1.8, rate_, coef);
    for (int k = 0; k < kFirMax; k++) h_taps_[(size_t)k * stride_ + i] = k < n ? coef[k] : 0.0;
    h_par_[(size_t)P_NTAPS * stride_ + i] = n;
    h_reset_[i] |= R_FIR;
    dirty_ = true;
}

void PostBank::set_fm(int i, int squelch_value, double fm_bw)
{
    h_par_[(size_t)P_SQ_THRESH * stride_ + i] = (double)(5000.0 - ((5000.0 * squelch_value) / 99));
    if (fm_bw_[i] != fm_bw) {     // dsp/fmdemod.cpp:160-164 -> InitNoiseSquelch re-designs and clears the HP FIR
        fm_bw_[i] = fm_bw;
        double coef[kFirMax];
        int n = design_kaiser_hp(1.0, 50.0, fm_bw, fm_bw * .6, rate_, coef);
        for (int k = 0; k < kFirMax; k++) h_taps_[(size_t)k * stride_ + i] = k < n ? coef[k] : 0.0;
        h_par_[(size_t)P_NTAPS * stride_ + i] = n;
        h_reset_[i] |= R_FIR;
    }
    dirty_ = true;
}

int PostBank::upload()
{
    if (!dirty_) return CUTESDR_OK;
    CSDR_CK(cudaMemcpyAsync(d_par_, h_par_.data(), h_par_.size() * sizeof(double), cudaMemcpyHostToDevice, st_));
    CSDR_CK(cudaMemcpyAsync(d_taps_, h_taps_.data(), h_taps_.size() * sizeof(double), cudaMemcpyHostToDevice, st_));
    CSDR_CK(cudaMemcpyAsync(d_mode_, h_mode_.data(), stride_ * sizeof(int), cudaMemcpyHostToDevice, st_));
    // Reset flags: the host copy holds every flag raised since the last upload; the kernel clears
    // a device flag after acting on it, and every upload is followed by a run (upload() is only
    // called from run()), so no stale flag can survive.
    CSDR_CK(cudaMemcpyAsync(d_reset_, h_reset_.data(), stride_ * sizeof(int), cudaMemcpyHostToDevice, st_));
    CSDR_CK(cudaStreamSynchronize(st_));      // host vectors are reused below / by the next setter
    std::fill(h_reset_.begin(), h_reset_.end(), 0);
    dirty_ = false;
    return CUTESDR_OK;
}

int PostBank::run(const float2* d_y, int n, float* d_audio, int audio_stride, int audio_off, const int* d_chan_map,
                  float2* d_tap3)
{
    if (n <= 0) return CUTESDR_OK;
    if (n > max_n_) { set_error("PostBank::run: %d samples exceed capacity %d", n, max_n_); return CUTESDR_E_ARG; }
    CSDR_TRY(upload());
    k_post<<<(nch_ + 63) / 64, 64, 0, st_>>>(d_y, n, nch_, stride_, uni_, d_par_, d_taps_, d_mode_, d_reset_, d_state_,
                                             d_istate_, d_agc_delay_, d_agc_mag_, d_v_, d_audio, audio_stride, audio_off,
                                             d_chan_map, d_tap3);
    lc_->n++;
    CSDR_CK(cudaGetLastError());
    return CUTESDR_OK;
}

int PostBank::read_smeter(int i, double* peak, double* ave)
{
    double pk = 0, av = 0;
    CSDR_CK(cudaMemcpyAsync(&pk, d_state_ + (size_t)S_SM_PEAK * stride_ + i, sizeof(double), cudaMemcpyDeviceToHost, st_));
    CSDR_CK(cudaMemcpyAsync(&av, d_state_ + (size_t)S_SM_AVE * stride_ + i, sizeof(double), cudaMemcpyDeviceToHost, st_));
    // GetPeak resets the held peak (dsp/smeter.cpp:99-104)
    CSDR_CK(cudaMemsetAsync(d_state_ + (size_t)S_SM_PEAK * stride_ + i, 0, sizeof(double), st_));
    CSDR_CK(cudaStreamSynchronize(st_));
    if (peak) *peak = pk + 5.0;     // SMETER_CALIBRATION, dsp/smeter.cpp:45
    if (ave) *ave = av + 5.0;
    return CUTESDR_OK;
}

}  // namespace csdr
