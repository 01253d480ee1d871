/* cutesdr_oracle.h -- CPU restatement (double precision, plain C) of the CuteSDR
 * receive DSP chain, used ONLY as the parity checker for libcutesdr_cuda.
 *
 * TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg may load liboracle.so. The product never calls it.
 *
 * Parity pinning: every function below is checked in tests/test_oracle_vs_ref.py
 * against the unmodified reference compiled headless (oracle/_ref/libcutesdr_ref.so)
 * and against the committed fixtures in tests/golden/ (generated from that same
 * build by tests/golden/make_golden.py). The reference itself ships no tests or
 * golden vectors.
 *
 * Citations are file:line relative to the reference tree.
 */
#ifndef CUTESDR_ORACLE_H
#define CUTESDR_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { double re, im; } orc_cpx;

/* demod modes, dsp/demodulator.h:20-28 */
enum { ORC_AM = 0, ORC_SAM = 1, ORC_FM = 2, ORC_USB = 3, ORC_LSB = 4, ORC_CWU = 5, ORC_CWL = 6 };

/* tDemodInfo without the QString, dsp/demodulator.h:35-54 */
typedef struct {
    int HiCut, HiCutmin, HiCutmax, LowCut, LowCutmin, LowCutmax, Offset, SquelchValue;
    int AgcSlope, AgcThresh, AgcManualGain, AgcDecay, AgcOn, AgcHangOn;
} orc_demod_info;

/* ---- decimation ladder: CDownConvert::SetDataRate, dsp/downconvert.cpp:114-173 ---- */
int orc_plan_stages(double in_rate, double max_bw, int* lens, int cap, double* out_rate);

/* ---- CDownConvert, dsp/downconvert.cpp:60-460 ---- */
typedef struct orc_downconvert orc_downconvert;
orc_downconvert* orc_downconvert_create(void);
void orc_downconvert_destroy(orc_downconvert*);
void orc_downconvert_set_frequency(orc_downconvert*, double nco_freq);
void orc_downconvert_set_cw_offset(orc_downconvert*, double off);
double orc_downconvert_set_data_rate(orc_downconvert*, double in_rate, double max_bw);
int orc_downconvert_stages(const orc_downconvert*, int* lens, int cap);
/* mutates `in` (NCO product is written back, :238-239) */
int orc_downconvert_process(orc_downconvert*, int n, orc_cpx* in, orc_cpx* out);

/* ---- CFastFIR, dsp/fastfir.cpp:55-321 ---- */
typedef struct orc_fastfir orc_fastfir;
orc_fastfir* orc_fastfir_create(void);
void orc_fastfir_destroy(orc_fastfir*);
void orc_fastfir_setup(orc_fastfir*, double lo, double hi, double offset, double rate);
int orc_fastfir_process(orc_fastfir*, int n, const orc_cpx* in, orc_cpx* out);
/* time-domain taps (1025, already scaled by 1/2048 like the reference's) */
void orc_fastfir_taps(const orc_fastfir*, orc_cpx* taps1025);

/* ---- CFft display path, dsp/fft.cpp:118-410,510-589 ---- */
typedef struct orc_fft orc_fft;
orc_fft* orc_fft_create(void);
void orc_fft_destroy(orc_fft*);
void orc_fft_set_params(orc_fft*, int size, int invert, double db_comp, double sample_freq);
void orc_fft_set_ave(orc_fft*, int ave);
void orc_fft_reset(orc_fft*);
int orc_fft_put(orc_fft*, int n, const orc_cpx* in);
int orc_fft_get_screen(orc_fft*, int max_h, int max_w, double max_db, double min_db,
                       int start_freq, int stop_freq, int* out);
int orc_fft_size(const orc_fft*);
void orc_fft_avebuf(const orc_fft*, double* out);

/* ---- CSMeter, dsp/smeter.cpp:49-112 ---- */
typedef struct orc_smeter orc_smeter;
orc_smeter* orc_smeter_create(void);
void orc_smeter_destroy(orc_smeter*);
void orc_smeter_process(orc_smeter*, int n, const orc_cpx* in, double rate);
double orc_smeter_peak(orc_smeter*);
double orc_smeter_ave(const orc_smeter*);

/* ---- CAgc (complex path), dsp/agc.cpp:80-296 ---- */
typedef struct orc_agc orc_agc;
orc_agc* orc_agc_create(void);
void orc_agc_destroy(orc_agc*);
void orc_agc_set(orc_agc*, int on, int hang, int thresh, int manual_gain, int slope, int decay, double rate);
void orc_agc_process(orc_agc*, int n, const orc_cpx* in, orc_cpx* out);

/* ---- CFir, dsp/fir.cpp:57-432 ---- */
typedef struct orc_fir orc_fir;
orc_fir* orc_fir_create(void);
void orc_fir_destroy(orc_fir*);
int orc_fir_init_lp(orc_fir*, double scale, double astop, double fpass, double fstop, double fs);
int orc_fir_init_hp(orc_fir*, double scale, double astop, double fpass, double fstop, double fs);
void orc_fir_make_hilbert_pair(orc_fir*, double freq_offset);
int orc_fir_taps(const orc_fir*, double* coef, double* icoef, double* qcoef);
void orc_fir_process_real(orc_fir*, int n, const double* in, double* out);
void orc_fir_process_cpx(orc_fir*, int n, const orc_cpx* in, orc_cpx* out);

/* ---- CIir (low-pass only is live), dsp/iir.cpp:86-101,171-180 ---- */
typedef struct { double a1, a2, b0, b1, b2, w1, w2; } orc_biquad;
void orc_biquad_init_lp(orc_biquad*, double f0, double q, double fs);
void orc_biquad_process(orc_biquad*, int n, const double* in, double* out);

/* ---- demodulators ---- */
typedef struct orc_am orc_am;     /* dsp/amdemod.cpp:50-104 */
orc_am* orc_am_create(double rate);
void orc_am_destroy(orc_am*);
void orc_am_set_bandwidth(orc_am*, double bw);
int orc_am_process(orc_am*, int n, const orc_cpx* in, double* out);

typedef struct orc_sam orc_sam;   /* dsp/samdemod.cpp:54-158 */
orc_sam* orc_sam_create(double rate);
void orc_sam_destroy(orc_sam*);
int orc_sam_process(orc_sam*, int n, const orc_cpx* in, double* out);
int orc_sam_process_stereo(orc_sam*, int n, const orc_cpx* in, orc_cpx* out);

typedef struct orc_fm orc_fm;     /* dsp/fmdemod.cpp:62-236 */
orc_fm* orc_fm_create(double rate);
void orc_fm_destroy(orc_fm*);
void orc_fm_set_squelch(orc_fm*, int value);
int orc_fm_process(orc_fm*, int n, double fm_bw, const orc_cpx* in, double* out);

/* ---- CFractResampler, dsp/fractresampler.cpp:50-352 ---- */
typedef struct orc_resampler orc_resampler;
orc_resampler* orc_resampler_create(int max_input);
void orc_resampler_destroy(orc_resampler*);
int orc_resampler_real(orc_resampler*, int n, double rate, const double* in, double* out);
int orc_resampler_cpx(orc_resampler*, int n, double rate, const orc_cpx* in, orc_cpx* out);
int orc_resampler_mono16(orc_resampler*, int n, double rate, const double* in, short* out, double gain);
int orc_resampler_stereo16(orc_resampler*, int n, double rate, const orc_cpx* in, short* out, double gain);
/* the 280001-entry window-sinc table, dsp/fractresampler.cpp:104-115 */
const double* orc_resampler_table(const orc_resampler*, int* len);

/* ---- CNoiseProc blanker, dsp/noiseproc.cpp:59-176 ---- */
typedef struct orc_blanker orc_blanker;
orc_blanker* orc_blanker_create(void);
void orc_blanker_destroy(orc_blanker*);
void orc_blanker_setup(orc_blanker*, int on, double threshold, double width_us, double fs);
void orc_blanker_process(orc_blanker*, long n, orc_cpx* io);

/* ---- CDemodulator sequencer (mono), dsp/demodulator.cpp:47-215 ---- */
typedef struct orc_demod orc_demod;
orc_demod* orc_demod_create(void);
void orc_demod_destroy(orc_demod*);
void orc_demod_set_input_rate(orc_demod*, double rate);
void orc_demod_set_demod(orc_demod*, int mode, const orc_demod_info* info);
void orc_demod_set_freq(orc_demod*, double f);
double orc_demod_output_rate(const orc_demod*);
int orc_demod_inbuf_limit(const orc_demod*);
/* test aid, not in the reference: override m_InBufLimit until the next set_demod */
void orc_demod_set_inbuf_limit(orc_demod*, int limit);
double orc_demod_smeter_peak(orc_demod*);
double orc_demod_smeter_ave(const orc_demod*);
/* tap capture: profile 1..4 as in gui/testbench.cpp:71-81 (1=post-downconvert cpx,
 * 2=post-FIR cpx, 3=post-AGC cpx, 4=audio real). buf may be NULL to disable. */
void orc_demod_set_tap(orc_demod*, int profile, double* buf, long cap_doubles);
long orc_demod_tap_count(const orc_demod*, int profile);
/* returns number of mono samples written to out (0 or 1024 per completed DSP block) */
int orc_demod_process(orc_demod*, int n, const orc_cpx* in, double* out);
/* convenience: feed complex64 stream in `packet`-sized calls */
long orc_demod_run_c64(orc_demod*, long n, const float* iq, int packet, double* out, long out_cap);

#ifdef __cplusplus
}
#endif
#endif
