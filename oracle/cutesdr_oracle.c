/* cutesdr_oracle.c -- see cutesdr_oracle.h. TEST INFRASTRUCTURE ONLY (parity checker).
 *
 * A from-scratch double-precision restatement of the reference receive chain. The
 * arithmetic follows the reference's evaluation order so that, apart from the
 * FFT (a plain radix-2 here instead of the vendored Ooura radix-4), outputs agree
 * with the compiled reference to the last bit or two; tests pin that.
 */
#include "cutesdr_oracle.h"
#include "halfband_tables.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define TWO_PI (2.0 * 3.14159265358979323846)   /* K_2PI, dsp/datatypes.h:42 */
#define ONE_PI (3.14159265358979323846)

/* ======================================================================= */
/* decimation ladder                                                        */
/* ======================================================================= */

/* dsp/downconvert.cpp:127-166: walk rates fs, fs/2, ... while the rate is above
 * MaxBW/HB51 limit and above 15800 Hz; at each rate pick the cheapest stage whose
 * alias-free limit still covers MaxBW. */
int orc_plan_stages(double in_rate, double max_bw, int* lens, int cap, double* out_rate)
{
    int n = 0;
    double f = in_rate;
    const double hb51_limit = .5 - orc_hb_alias_free[ORC_HB_NUM_KINDS - 1];
    while (f > (max_bw / hb51_limit) && f > (7900.0 * 2.0)) {
        for (int k = 0; k < ORC_HB_NUM_KINDS; k++) {
            if (f >= (max_bw / (.5 - orc_hb_alias_free[k]))) {
                if (n < cap) lens[n] = orc_hb_len[k];
                n++;
                break;
            }
        }
        f /= 2.0;
    }
    if (out_rate) *out_rate = f;
    return n;
}

/* ======================================================================= */
/* down-converter                                                           */
/* ======================================================================= */

#define ORC_MAX_STAGES 24
#define ORC_HB_WORK 32768          /* MAX_HALF_BAND_BUFSIZE, dsp/downconvert.cpp:54 */

typedef struct {
    int len;                       /* 3 = CIC3, 11 = fixed 11-tap, else generic */
    double h[51];                  /* full expanded taps */
    orc_cpx* work;                 /* generic: history(len-1) + block */
    orc_cpx d[10];                 /* HB11 delay (d0..d9) / CIC: d[0]=Xodd d[1]=Xeven */
} orc_stage;

struct orc_downconvert {
    double out_rate, nco_freq, cw_offset, nco_inc, in_rate, max_bw;
    orc_cpx osc1;
    double osc_cos, osc_sin;
    int nstages;
    orc_stage st[ORC_MAX_STAGES];
};

static void stage_free(orc_downconvert* d)
{
    for (int i = 0; i < d->nstages; i++) { free(d->st[i].work); d->st[i].work = NULL; }
    d->nstages = 0;
}

static void stage_init(orc_stage* s, int len)
{
    memset(s, 0, sizeof(*s));
    s->len = len;
    if (len > 3) {
        int kind = 0;
        for (int k = 1; k < ORC_HB_NUM_KINDS; k++) if (orc_hb_len[k] == len) kind = k;
        const double* u = orc_hb_taps + orc_hb_tap_off[kind];
        int c = (len - 1) / 2;
        for (int i = 0; i < len; i++) s->h[i] = 0.0;
        for (int i = 0, j = 0; i < c; i += 2, j++) { s->h[i] = u[j]; s->h[len - 1 - i] = u[j]; }
        s->h[c] = 0.5;
        if (len != 11) s->work = (orc_cpx*)calloc(ORC_HB_WORK, sizeof(orc_cpx));
    }
}

orc_downconvert* orc_downconvert_create(void)
{
    /* dsp/downconvert.cpp:60-73 */
    orc_downconvert* d = (orc_downconvert*)calloc(1, sizeof(*d));
    d->in_rate = 100000.0;
    d->max_bw = 10000.0;
    d->osc1.re = 1.0;
    return d;
}

void orc_downconvert_destroy(orc_downconvert* d) { if (d) { stage_free(d); free(d); } }

void orc_downconvert_set_frequency(orc_downconvert* d, double nco_freq)
{
    /* dsp/downconvert.cpp:98-107 -- stores the sum, so a later SetDataRate re-adds the CW offset */
    double f = nco_freq + d->cw_offset;
    d->nco_freq = f;
    d->nco_inc = TWO_PI * d->nco_freq / d->in_rate;
    d->osc_cos = cos(d->nco_inc);
    d->osc_sin = sin(d->nco_inc);
}

void orc_downconvert_set_cw_offset(orc_downconvert* d, double off) { d->cw_offset = off; }

double orc_downconvert_set_data_rate(orc_downconvert* d, double in_rate, double max_bw)
{
    /* dsp/downconvert.cpp:114-173 */
    if (d->in_rate != in_rate || d->max_bw != max_bw) {
        int lens[ORC_MAX_STAGES];
        d->in_rate = in_rate;
        d->max_bw = max_bw;
        stage_free(d);
        int n = orc_plan_stages(in_rate, max_bw, lens, ORC_MAX_STAGES, &d->out_rate);
        if (n > ORC_MAX_STAGES) n = ORC_MAX_STAGES;
        for (int i = 0; i < n; i++) stage_init(&d->st[i], lens[i]);
        d->nstages = n;
        orc_downconvert_set_frequency(d, d->nco_freq);
    }
    return d->out_rate;
}

int orc_downconvert_stages(const orc_downconvert* d, int* lens, int cap)
{
    for (int i = 0; i < d->nstages && i < cap; i++) lens[i] = d->st[i].len;
    return d->nstages;
}

/* dsp/downconvert.cpp:444-460 */
static int dec_cic3(orc_stage* s, int n, const orc_cpx* in, orc_cpx* out)
{
    int j = 0;
    for (int i = 0; i < n; i += 2, j++) {
        orc_cpx even = in[i], odd = in[i + 1];
        out[j].re = .125 * (odd.re + s->d[1].re + 3.0 * (s->d[0].re + even.re));
        out[j].im = .125 * (odd.im + s->d[1].im + 3.0 * (s->d[0].im + even.im));
        s->d[0] = odd;
        s->d[1] = even;
    }
    return j;
}

/* dsp/downconvert.cpp:348-423. One uniform formula over [d0..d9 | block]; summation
 * order H0,H2,H4,H5,H6,H8,H10 as in the reference's unrolled expressions. */
static int dec_hb11(orc_stage* s, int n, const orc_cpx* in, orc_cpx* out)
{
    const double H0 = s->h[0], H2 = s->h[2], H4 = s->h[4], H5 = s->h[5], H6 = s->h[6], H8 = s->h[8], H10 = s->h[10];
    int nout = n / 2;
    orc_cpx* x = (orc_cpx*)malloc((size_t)(n + 10) * sizeof(orc_cpx));
    memcpy(x, s->d, 10 * sizeof(orc_cpx));
    memcpy(x + 10, in, (size_t)n * sizeof(orc_cpx));
    for (int k = 0; k < nout; k++) {
        const orc_cpx* p = x + 2 * k;   /* p[0] == x[2k-10] */
        out[k].re = H0 * p[0].re + H2 * p[2].re + H4 * p[4].re + H5 * p[5].re + H6 * p[6].re + H8 * p[8].re + H10 * p[10].re;
        out[k].im = H0 * p[0].im + H2 * p[2].im + H4 * p[4].im + H5 * p[5].im + H6 * p[6].im + H8 * p[8].im + H10 * p[10].im;
    }
    memcpy(s->d, x + n, 10 * sizeof(orc_cpx));   /* last 10 inputs */
    free(x);
    return nout;
}

/* dsp/downconvert.cpp:286-320 */
static int dec_hb(orc_stage* s, int n, const orc_cpx* in, orc_cpx* out)
{
    const int N = s->len, c = (N - 1) / 2;
    if (n < N) return n / 2;                         /* :291-292 -- nothing computed */
    orc_cpx* b = s->work;
    memcpy(b + (N - 1), in, (size_t)n * sizeof(orc_cpx));
    int nout = 0;
    for (int i = 0; i < n; i += 2) {
        orc_cpx acc;
        acc.re = b[i].re * s->h[0];
        acc.im = b[i].im * s->h[0];
        for (int j = 2; j < N; j += 2) {
            acc.re += b[i + j].re * s->h[j];
            acc.im += b[i + j].im * s->h[j];
        }
        acc.re += b[i + c].re * s->h[c];
        acc.im += b[i + c].im * s->h[c];
        out[nout++] = acc;
    }
    /* :314-317 reads the caller's buffer AFTER outputs were written; with in==out that is
     * only safe when n >= 2(N-1). Mirrored: the tail comes from `in` as it is now. */
    memcpy(b, in + (n - N + 1), (size_t)(N - 1) * sizeof(orc_cpx));
    return nout;
}

int orc_downconvert_process(orc_downconvert* d, int n, orc_cpx* in, orc_cpx* out)
{
    /* NCO: quadrature oscillator with amplitude servo, dsp/downconvert.cpp:203-240 */
    for (int i = 0; i < n; i++) {
        orc_cpx x = in[i], o;
        o.re = d->osc1.re * d->osc_cos - d->osc1.im * d->osc_sin;
        o.im = d->osc1.im * d->osc_cos + d->osc1.re * d->osc_sin;
        double g = 1.95 - (d->osc1.re * d->osc1.re + d->osc1.im * d->osc1.im);
        d->osc1.re = g * o.re;
        d->osc1.im = g * o.im;
        in[i].re = (x.re * o.re) - (x.im * o.im);
        in[i].im = (x.re * o.im) + (x.im * o.re);
    }
    int m = n;
    for (int s = 0; s < d->nstages; s++) {
        orc_stage* st = &d->st[s];
        if (st->len == 3) m = dec_cic3(st, m, in, in);
        else if (st->len == 11) m = dec_hb11(st, m, in, in);
        else m = dec_hb(st, m, in, in);
    }
    for (int i = 0; i < m; i++) out[i] = in[i];
    return m;
}

/* ======================================================================= */
/* plain radix-2 complex FFT (sign = -1 forward DFT, +1 inverse, unnormalised) */
/* ======================================================================= */
static void fft_radix2(orc_cpx* a, int n, int sign)
{
    for (int i = 1, j = 0; i < n; i++) {
        int bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) { orc_cpx t = a[i]; a[i] = a[j]; a[j] = t; }
    }
    for (int len = 2; len <= n; len <<= 1) {
        double ang = sign * TWO_PI / len;
        int half = len >> 1;
        for (int k = 0; k < half; k++) {
            double wr = cos(ang * k), wi = sin(ang * k);
            for (int i = k; i < n; i += len) {
                orc_cpx u = a[i], v = a[i + half];
                double tr = v.re * wr - v.im * wi, ti = v.re * wi + v.im * wr;
                a[i].re = u.re + tr; a[i].im = u.im + ti;
                a[i + half].re = u.re - tr; a[i + half].im = u.im - ti;
            }
        }
    }
}

/* ======================================================================= */
/* fast FIR (overlap-save)                                                  */
/* ======================================================================= */
#define FF_FFT 2048                /* CONV_FFT_SIZE, dsp/fastfir.cpp:55 */
#define FF_FIR 1025                /* CONV_FIR_SIZE, :56 */

struct orc_fastfir {
    double lo, hi, offset, rate;
    double window[FF_FIR];
    orc_cpx taps[FF_FIR];          /* time domain, scaled 1/2048 */
    orc_cpx H[FF_FFT];             /* DFT of zero-padded taps */
    orc_cpx buf[FF_FFT];
    int pos;
};

orc_fastfir* orc_fastfir_create(void)
{
    orc_fastfir* f = (orc_fastfir*)calloc(1, sizeof(*f));
    /* Blackman-Nuttall, dsp/fastfir.cpp:93-101 */
    for (int i = 0; i < FF_FIR; i++)
        f->window[i] = (0.3635819
            - 0.4891775 * cos((TWO_PI * i) / (FF_FIR - 1))
            + 0.1365995 * cos((2.0 * TWO_PI * i) / (FF_FIR - 1))
            - 0.0106411 * cos((3.0 * TWO_PI * i) / (FF_FIR - 1)));
    f->pos = FF_FIR - 1;
    f->lo = -1.0; f->hi = 1.0; f->offset = 1.0; f->rate = 1.0;   /* :126-129 */
    return f;
}

void orc_fastfir_destroy(orc_fastfir* f) { free(f); }

void orc_fastfir_setup(orc_fastfir* f, double lo, double hi, double offset, double rate)
{
    /* dsp/fastfir.cpp:178-259 */
    if (lo == f->lo && hi == f->hi && offset == f->offset && rate == f->rate) return;
    f->lo = lo; f->hi = hi; f->offset = offset; f->rate = rate;
    lo += offset;
    hi += offset;
    if (lo >= hi || lo >= rate / 2.0 || lo <= -rate / 2.0 || hi >= rate / 2.0 || hi <= -rate / 2.0)
        return;                    /* new params stay stored, old filter stays active (:195-203) */
    double nFL = lo / rate, nFH = hi / rate;
    double nFc = (nFH - nFL) / 2.0;
    double nFs = TWO_PI * (nFH + nFL) / 2.0;
    double centre = 0.5 * (double)(FF_FIR - 1);
    memset(f->H, 0, sizeof(f->H));
    for (int i = 0; i < FF_FIR; i++) {
        double x = (double)i - centre, z;
        if ((double)i == centre) z = 2.0 * nFc;
        else z = sin(TWO_PI * x * nFc) / (ONE_PI * x) * f->window[i];
        f->taps[i].re = z * cos(nFs * x) / (double)FF_FFT;
        f->taps[i].im = z * sin(nFs * x) / (double)FF_FFT;
        f->H[i] = f->taps[i];
    }
    /* The reference transforms with its e^{+j} "forward" kernel and inverts with the
     * conjugate kernel (dsp/fft.cpp:416-426); convolution is the same for either sign
     * pair, so the oracle uses the textbook pair. */
    fft_radix2(f->H, FF_FFT, -1);
}

void orc_fastfir_taps(const orc_fastfir* f, orc_cpx* taps1025) { memcpy(taps1025, f->taps, sizeof(f->taps)); }

int orc_fastfir_process(orc_fastfir* f, int n, const orc_cpx* in, orc_cpx* out)
{
    /* dsp/fastfir.cpp:268-306: buffer = [1024 previous | 1024 new]; emit samples 1024..2047 */
    int outpos = 0;
    static __thread orc_cpx work[FF_FFT];
    for (int i = 0; i < n; i++) {
        f->buf[f->pos++] = in[i];
        if (f->pos >= FF_FFT) {
            memcpy(work, f->buf, sizeof(work));
            fft_radix2(work, FF_FFT, -1);
            for (int k = 0; k < FF_FFT; k++) {
                double sr = work[k].re, si = work[k].im;
                work[k].re = f->H[k].re * sr - f->H[k].im * si;
                work[k].im = f->H[k].re * si + f->H[k].im * sr;
            }
            fft_radix2(work, FF_FFT, +1);
            for (int k = FF_FIR - 1; k < FF_FFT; k++) out[outpos++] = work[k];
            memmove(f->buf, f->buf + (FF_FFT - (FF_FIR - 1)), (FF_FIR - 1) * sizeof(orc_cpx));
            f->pos = FF_FIR - 1;
        }
    }
    return outpos;
}

/* ======================================================================= */
/* display FFT                                                              */
/* ======================================================================= */
struct orc_fft {
    int overload, invert, ave_count, total_count, size, last_size, ave_size;
    int start_freq, stop_freq, bin_min, bin_max, plot_width;
    double k_c, k_b, db_comp, sample_freq;
    double *window, *pwr_ave, *ave, *sum;
    int* translate;
    orc_cpx* work;
};

orc_fft* orc_fft_create(void)
{
    /* dsp/fft.cpp:29-50 */
    orc_fft* f = (orc_fft*)calloc(1, sizeof(*f));
    f->ave_size = 1;
    f->size = 1024;
    f->db_comp = 0.0;
    orc_fft_set_params(f, 2048, 0, 0.0, 1000);
    orc_fft_set_ave(f, 1);
    return f;
}

static void fft_free(orc_fft* f)
{
    free(f->window); free(f->pwr_ave); free(f->ave); free(f->sum); free(f->translate); free(f->work);
    f->window = f->pwr_ave = f->ave = f->sum = NULL; f->translate = NULL; f->work = NULL;
}

void orc_fft_destroy(orc_fft* f) { if (f) { fft_free(f); free(f); } }

void orc_fft_set_ave(orc_fft* f, int ave)
{
    /* dsp/fft.cpp:103-113 */
    if (f->ave_size != ave) f->ave_size = ave > 0 ? ave : 1;
    orc_fft_reset(f);
}

void orc_fft_reset(orc_fft* f)
{
    /* dsp/fft.cpp:248-259 (PwrAve is deliberately not cleared there) */
    for (int i = 0; i < f->size; i++) { f->ave[i] = 0.0; f->sum[i] = 0.0; }
    f->ave_count = 0;
    f->total_count = 0;
}

void orc_fft_set_params(orc_fft* f, int size, int invert, double db_comp, double sample_freq)
{
    /* dsp/fft.cpp:118-243 */
    if (size == 0) return;
    f->bin_min = f->bin_max = f->start_freq = f->stop_freq = f->plot_width = 0;
    f->invert = invert;
    f->sample_freq = sample_freq;
    if (f->db_comp != db_comp) { f->last_size = 0; f->db_comp = db_comp; }
    if (size < 512) f->size = 512;
    else if (size > 65536) f->size = 65536;
    else f->size = size;
    if (f->last_size != f->size) {
        int N = f->size;
        f->last_size = N;
        fft_free(f);
        f->window = (double*)malloc(N * sizeof(double));
        f->pwr_ave = (double*)calloc(N, sizeof(double));
        f->ave = (double*)calloc(N, sizeof(double));
        f->sum = (double*)calloc(N, sizeof(double));
        f->translate = (int*)calloc(N + 65536, sizeof(int));   /* reference sizes this N and overruns when width > N */
        f->work = (orc_cpx*)calloc(N, sizeof(orc_cpx));
        f->k_b = f->db_comp - 20 * log10((double)N * 32767.0 / 2.0);
        f->k_c = pow(10.0, (-220.0 - f->k_b) / 10.0);
        f->k_b = f->k_b / 10.0;
        for (int i = 0; i < N; i++) f->window[i] = 2.0 * (.5 - .5 * cos((TWO_PI * i) / (N - 1)));   /* Hann x2 */
    }
    orc_fft_reset(f);
}

int orc_fft_size(const orc_fft* f) { return f->size; }
void orc_fft_avebuf(const orc_fft* f, double* out) { memcpy(out, f->ave, f->size * sizeof(double)); }

int orc_fft_put(orc_fft* f, int n, const orc_cpx* in)
{
    /* dsp/fft.cpp:267-288 + power section of CpxFFT :560-589.
     * The reference swaps I and Q and runs its e^{+j} kernel; |.|^2 of that equals
     * |DFT|^2 of the un-swapped windowed input, which is what is computed here. */
    const int N = f->size;
    f->overload = 0;
    for (int i = 0; i < n; i++) {
        if (in[i].re > 32000.0) f->overload = 1;
        double w = f->window[i];
        f->work[i].re = w * in[i].re;
        f->work[i].im = w * in[i].im;
    }
    fft_radix2(f->work, N, -1);
    f->total_count++;
    if (f->ave_count < f->ave_size) f->ave_count++;
    for (int k = 0; k < N; k++) {
        int j = (k < N / 2) ? k + N / 2 : k - N / 2;     /* fft-shift: index N/2 is DC */
        double p = f->work[k].re * f->work[k].re + f->work[k].im * f->work[k].im;
        if (f->total_count <= f->ave_size) f->sum[j] = f->sum[j] + p;
        else f->sum[j] = f->sum[j] - f->pwr_ave[j] + p;
        f->pwr_ave[j] = f->sum[j] / (double)f->ave_count;
        f->ave[j] = log10(f->pwr_ave[j] + f->k_c) + f->k_b;
    }
    return f->total_count;
}

int orc_fft_get_screen(orc_fft* f, int max_h, int max_w, double max_db, double min_db,
                       int start_freq, int stop_freq, int* out)
{
    /* dsp/fft.cpp:308-410 */
    int ymax = 10000, xprev = -1;
    double off = max_db / 10.0;
    double gain = -10.0 / (max_db - min_db);
    const int N = f->size;
    if (f->start_freq != start_freq || f->stop_freq != stop_freq || f->plot_width != max_w) {
        int maxbin = N - 1;
        f->start_freq = start_freq; f->stop_freq = stop_freq; f->plot_width = max_w;
        f->bin_min = (int)((double)start_freq * (double)N / f->sample_freq) + N / 2;
        f->bin_max = (int)((double)stop_freq * (double)N / f->sample_freq) + N / 2;
        if (f->bin_min < 0) f->bin_min = 0;
        if (f->bin_min >= maxbin) f->bin_min = maxbin;
        if (f->bin_max < 0) f->bin_max = 0;
        if (f->bin_max >= maxbin) f->bin_max = maxbin;
        if ((f->bin_max - f->bin_min) > f->plot_width) {
            for (int i = f->bin_min; i <= f->bin_max; i++)
                f->translate[i] = ((i - f->bin_min) * f->plot_width) / (f->bin_max - f->bin_min);
        } else {
            for (int i = 0; i < f->plot_width; i++)
                f->translate[i] = f->bin_min + (i * (f->bin_max - f->bin_min)) / f->plot_width;
        }
    }
    if ((f->bin_max - f->bin_min) > f->plot_width) {
        for (int i = f->bin_min; i <= f->bin_max; i++) {
            double v = f->invert ? f->ave[N - i] : f->ave[i];
            int y = (int)((double)max_h * gain * (v - off));
            if (y < 0) y = 0;
            if (y > max_h) y = max_h;
            int x = f->translate[i];
            if (x == xprev) {
                if (y < ymax) { out[x] = y; ymax = y; }       /* smaller y = stronger: peak hold */
            } else { out[x] = y; xprev = x; ymax = y; }
        }
    } else {
        for (int x = 0; x < f->plot_width; x++) {
            int i = f->translate[x];
            double v = f->invert ? f->ave[N - i] : f->ave[i];
            int y = (int)((double)max_h * gain * (v - off));
            if (y < 0) y = 0;
            if (y > max_h) y = max_h;
            out[x] = y;
        }
    }
    return f->overload;
}

/* ======================================================================= */
/* S-meter                                                                  */
/* ======================================================================= */
struct orc_smeter { double peak, rate, a_alpha, d_alpha, a_ave, d_ave, ave; };

orc_smeter* orc_smeter_create(void)
{
    orc_smeter* s = (orc_smeter*)calloc(1, sizeof(*s));
    s->rate = 1.0; s->a_alpha = 1.0; s->d_alpha = 1.0; s->a_ave = -120.0; s->d_ave = -120.0;
    return s;
}
void orc_smeter_destroy(orc_smeter* s) { free(s); }

void orc_smeter_process(orc_smeter* s, int n, const orc_cpx* in, double rate)
{
    /* dsp/smeter.cpp:62-93 */
    if (rate != s->rate) {
        s->rate = rate;
        s->a_alpha = (1.0 - exp(-1.0 / (rate * .01)));
        s->d_alpha = (1.0 - exp(-1.0 / (rate * .5)));
    }
    for (int i = 0; i < n; i++) {
        double mag = 10.0 * log10((in[i].re * in[i].re + in[i].im * in[i].im) / (32767.0 * 32767.0) + 1e-50);
        s->a_ave = (1.0 - s->a_alpha) * s->a_ave + s->a_alpha * mag;
        s->d_ave = (1.0 - s->d_alpha) * s->d_ave + s->d_alpha * mag;
        if (s->a_ave > s->d_ave) { s->ave = s->a_ave; s->d_ave = s->a_ave; }
        else s->ave = s->d_ave;
        if (mag > s->peak) s->peak = mag;
    }
}
double orc_smeter_peak(orc_smeter* s) { double x = s->peak; s->peak = 0; return x + 5.0; }
double orc_smeter_ave(const orc_smeter* s) { return s->ave + 5.0; }

/* ======================================================================= */
/* AGC                                                                      */
/* ======================================================================= */
#define AGC_BUF 2048               /* MAX_DELAY_BUF, dsp/agc.h:15 */
struct orc_agc {
    int on, hang, thresh, manual_gain, decay;
    double slope_factor;           /* the reference stores the int in a double member */
    double rate;
    double manual_agc_gain, decay_ave, attack_ave;
    double a_rise, a_fall, d_rise, d_fall;
    double fixed_gain, knee, gain_slope, peak;
    int delay_ptr, mag_pos, delay_samples, window_samples, hang_time, hang_timer;
    orc_cpx delay[AGC_BUF];
    double mag[AGC_BUF];
};

orc_agc* orc_agc_create(void)
{
    orc_agc* a = (orc_agc*)calloc(1, sizeof(*a));
    a->on = 1; a->rate = 100.0;    /* dsp/agc.cpp:80-89 */
    return a;
}
void orc_agc_destroy(orc_agc* a) { free(a); }

void orc_agc_set(orc_agc* a, int on, int hang, int thresh, int manual_gain, int slope, int decay, double rate)
{
    /* dsp/agc.cpp:104-167 */
    if (on == a->on && hang == a->hang && thresh == a->thresh && manual_gain == a->manual_gain &&
        slope == a->slope_factor && decay == a->decay && rate == a->rate)
        return;
    a->on = on; a->hang = hang; a->thresh = thresh; a->manual_gain = manual_gain;
    a->slope_factor = slope; a->decay = decay;
    if (a->rate != rate) {
        a->rate = rate;
        for (int i = 0; i < AGC_BUF; i++) { a->delay[i].re = a->delay[i].im = 0.0; a->mag[i] = -16.0; }
        a->delay_ptr = 0; a->hang_timer = 0; a->peak = -16.0;
        a->decay_ave = -5.0; a->attack_ave = -5.0; a->mag_pos = 0;
    }
    a->manual_agc_gain = 32767.0 * pow(10.0, -(100 - (double)a->manual_gain) / 20.0);
    a->knee = (double)a->thresh / 20.0;
    a->gain_slope = a->slope_factor / (100.0);
    a->fixed_gain = 0.7 * pow(10.0, a->knee * (a->gain_slope - 1.0));
    a->a_rise = (1.0 - exp(-1.0 / (a->rate * .002)));
    a->a_fall = (1.0 - exp(-1.0 / (a->rate * .005)));
    a->d_rise = (1.0 - exp(-1.0 / (a->rate * (double)a->decay * .001 * .3)));
    a->hang_time = (int)(a->rate * (double)a->decay * .001);
    if (a->hang) a->d_fall = (1.0 - exp(-1.0 / (a->rate * .05)));
    else a->d_fall = (1.0 - exp(-1.0 / (a->rate * (double)a->decay * .001)));
    a->delay_samples = (int)(a->rate * .015);
    a->window_samples = (int)(a->rate * .018);
    if (a->delay_samples >= AGC_BUF - 1) a->delay_samples = AGC_BUF - 1;
}

void orc_agc_process(orc_agc* a, int n, const orc_cpx* in, orc_cpx* out)
{
    /* dsp/agc.cpp:174-296 */
    if (!a->on) {
        for (int i = 0; i < n; i++) {
            orc_cpx x = in[i];
            out[i].re = a->manual_agc_gain * x.re;
            out[i].im = a->manual_agc_gain * x.im;
        }
        return;
    }
    for (int i = 0; i < n; i++) {
        orc_cpx x = in[i];
        orc_cpx delayed = a->delay[a->delay_ptr];
        a->delay[a->delay_ptr++] = x;
        if (a->delay_ptr >= a->delay_samples) a->delay_ptr = 0;

        double mag = fabs(x.re), mim = fabs(x.im);
        if (mim > mag) mag = mim;
        mag = log10(mag + 3.2767e-4) - log10(32767.0);

        double oldest = a->mag[a->mag_pos];
        a->mag[a->mag_pos++] = mag;
        if (a->mag_pos >= a->window_samples) a->mag_pos = 0;
        if (mag > a->peak) a->peak = mag;
        else if (oldest == a->peak) {
            a->peak = -8.0;
            for (int k = 0; k < a->window_samples; k++) if (a->mag[k] > a->peak) a->peak = a->mag[k];
        }

        if (a->peak > a->attack_ave) a->attack_ave = (1.0 - a->a_rise) * a->attack_ave + a->a_rise * a->peak;
        else a->attack_ave = (1.0 - a->a_fall) * a->attack_ave + a->a_fall * a->peak;
        if (a->hang) {
            if (a->peak > a->decay_ave) {
                a->decay_ave = (1.0 - a->d_rise) * a->decay_ave + a->d_rise * a->peak;
                a->hang_timer = 0;
            } else if (a->hang_timer < a->hang_time) a->hang_timer++;
            else a->decay_ave = (1.0 - a->d_fall) * a->decay_ave + a->d_fall * a->peak;
        } else {
            if (a->peak > a->decay_ave) a->decay_ave = (1.0 - a->d_rise) * a->decay_ave + a->d_rise * a->peak;
            else a->decay_ave = (1.0 - a->d_fall) * a->decay_ave + a->d_fall * a->peak;
        }
        double m = a->attack_ave > a->decay_ave ? a->attack_ave : a->decay_ave;
        double gain = (m <= a->knee) ? a->fixed_gain : 0.7 * pow(10.0, m * (a->gain_slope - 1.0));
        out[i].re = delayed.re * gain;
        out[i].im = delayed.im * gain;
    }
}

/* ======================================================================= */
/* small FIR (Kaiser designs) and biquad                                    */
/* ======================================================================= */
#define FIR_MAX 75
struct orc_fir {
    double fs;
    int ntaps, state;
    double coef[FIR_MAX], icoef[FIR_MAX], qcoef[FIR_MAX];
    double rz[FIR_MAX];
    orc_cpx cz[FIR_MAX];
};

orc_fir* orc_fir_create(void)
{
    orc_fir* f = (orc_fir*)calloc(1, sizeof(*f));
    f->ntaps = 1;
    return f;
}
void orc_fir_destroy(orc_fir* f) { free(f); }

/* modified Bessel I0 by series, dsp/fir.cpp:414-432 */
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

static void fir_finish(orc_fir* f)
{
    for (int n = 0; n < f->ntaps; n++) { f->icoef[n] = f->coef[n]; f->qcoef[n] = f->coef[n]; }
    memset(f->rz, 0, sizeof(f->rz));
    memset(f->cz, 0, sizeof(f->cz));
    f->state = 0;
}

int orc_fir_init_lp(orc_fir* f, double scale, double astop, double fpass, double fstop, double fs)
{
    /* dsp/fir.cpp:173-261 */
    f->fs = fs;
    double nfp = fpass / fs, nfs = fstop / fs, nfc = (nfs + nfp) / 2.0;
    double beta = kaiser_beta(astop);
    f->ntaps = (astop - 8.0) / (2.285 * TWO_PI * (nfs - nfp)) + 1;
    if (f->ntaps > FIR_MAX) f->ntaps = FIR_MAX;
    if (f->ntaps < 3) f->ntaps = 3;
    double centre = .5 * (double)(f->ntaps - 1);
    double izb = bessel_i0(beta);
    for (int n = 0; n < f->ntaps; n++) {
        double x = (double)n - centre, c;
        if ((double)n == centre) c = 2.0 * nfc;
        else c = sin(TWO_PI * x * nfc) / (ONE_PI * x);
        x = ((double)n - ((double)f->ntaps - 1.0) / 2.0) / (((double)f->ntaps - 1.0) / 2.0);
        f->coef[n] = scale * c * bessel_i0(beta * sqrt(1 - (x * x))) / izb;
    }
    fir_finish(f);
    return f->ntaps;
}

int orc_fir_init_hp(orc_fir* f, double scale, double astop, double fpass, double fstop, double fs)
{
    /* dsp/fir.cpp:278-367 */
    f->fs = fs;
    double nfp = fpass / fs, nfs = fstop / fs, nfc = (nfs + nfp) / 2.0;
    double beta = kaiser_beta(astop);
    f->ntaps = (astop - 8.0) / (2.285 * TWO_PI * (nfp - nfs)) + 1;
    if (f->ntaps > (FIR_MAX - 1)) f->ntaps = FIR_MAX - 1;
    if (f->ntaps < 3) f->ntaps = 3;
    f->ntaps |= 1;
    double izb = bessel_i0(beta);
    double centre = .5 * (double)(f->ntaps - 1);
    for (int n = 0; n < f->ntaps; n++) {
        double x = (double)n - (double)(f->ntaps - 1) / 2.0, c;
        if ((double)n == centre) c = 1.0 - 2.0 * nfc;
        else c = (sin(ONE_PI * x) / (ONE_PI * x) - sin(TWO_PI * x * nfc) / (ONE_PI * x));
        x = ((double)n - ((double)f->ntaps - 1.0) / 2.0) / (((double)f->ntaps - 1.0) / 2.0);
        f->coef[n] = scale * c * bessel_i0(beta * sqrt(1 - (x * x))) / izb;
    }
    fir_finish(f);
    return f->ntaps;
}

void orc_fir_make_hilbert_pair(orc_fir* f, double freq_offset)
{
    /* dsp/fir.cpp:374-386 */
    for (int n = 0; n < f->ntaps; n++) {
        f->icoef[n] = 2.0 * f->coef[n] * cos((TWO_PI * freq_offset / f->fs) * ((double)n - ((double)(f->ntaps - 1) / 2.0)));
        f->qcoef[n] = 2.0 * f->coef[n] * sin((TWO_PI * freq_offset / f->fs) * ((double)n - ((double)(f->ntaps - 1) / 2.0)));
    }
}

int orc_fir_taps(const orc_fir* f, double* coef, double* icoef, double* qcoef)
{
    for (int i = 0; i < f->ntaps; i++) {
        if (coef) coef[i] = f->coef[i];
        if (icoef) icoef[i] = f->icoef[i];
        if (qcoef) qcoef[i] = f->qcoef[i];
    }
    return f->ntaps;
}

/* dsp/fir.cpp:72-91: circular delay line; newest sample meets coef[0]. The MAC order
 * (delay-line slot 0..N-1, i.e. coefficient index rotated by state) is kept so sums
 * round the same way. */
void orc_fir_process_real(orc_fir* f, int n, const double* in, double* out)
{
    const int N = f->ntaps;
    for (int i = 0; i < n; i++) {
        f->rz[f->state] = in[i];
        double acc = 0.0;
        for (int j = 0; j < N; j++) {
            int k = N - f->state + j;
            if (k >= N) k -= N;
            double p = f->coef[k] * f->rz[j];
            acc = (j == 0) ? p : acc + p;
        }
        if (--f->state < 0) f->state += N;
        out[i] = acc;
    }
}

void orc_fir_process_cpx(orc_fir* f, int n, const orc_cpx* in, orc_cpx* out)
{
    /* dsp/fir.cpp:101-127: I and Q get separate REAL coefficient sets */
    const int N = f->ntaps;
    for (int i = 0; i < n; i++) {
        f->cz[f->state] = in[i];
        orc_cpx acc = {0.0, 0.0};
        for (int j = 0; j < N; j++) {
            int k = N - f->state + j;
            if (k >= N) k -= N;
            double pr = f->icoef[k] * f->cz[j].re, pi = f->qcoef[k] * f->cz[j].im;
            if (j == 0) { acc.re = pr; acc.im = pi; } else { acc.re += pr; acc.im += pi; }
        }
        if (--f->state < 0) f->state += N;
        out[i] = acc;
    }
}

void orc_biquad_init_lp(orc_biquad* b, double f0, double q, double fs)
{
    /* dsp/iir.cpp:86-101 */
    double w0 = TWO_PI * f0 / fs;
    double alpha = sin(w0) / (2.0 * q);
    double A = 1.0 / (1.0 + alpha);
    b->b0 = A * ((1.0 - cos(w0)) / 2.0);
    b->b1 = A * (1.0 - cos(w0));
    b->b2 = A * ((1.0 - cos(w0)) / 2.0);
    b->a1 = A * (-2.0 * cos(w0));
    b->a2 = A * (1.0 - alpha);
    b->w1 = b->w2 = 0.0;
}

void orc_biquad_process(orc_biquad* b, int n, const double* in, double* out)
{
    /* dsp/iir.cpp:171-180, direct form II */
    for (int i = 0; i < n; i++) {
        double w0 = in[i] - b->a1 * b->w1 - b->a2 * b->w2;
        out[i] = b->b0 * w0 + b->b1 * b->w1 + b->b2 * b->w2;
        b->w2 = b->w1;
        b->w1 = w0;
    }
}

/* ======================================================================= */
/* AM / SAM / FM                                                            */
/* ======================================================================= */
struct orc_am { double rate, z1; orc_fir fir; };

orc_am* orc_am_create(double rate)
{
    orc_am* a = (orc_am*)calloc(1, sizeof(*a));
    a->rate = rate;
    a->fir.ntaps = 1;
    orc_fir_init_lp(&a->fir, 1.0, 50.0, 10000, 10000 * 1.8, rate);   /* dsp/amdemod.cpp:50-54 */
    return a;
}
void orc_am_destroy(orc_am* a) { free(a); }
void orc_am_set_bandwidth(orc_am* a, double bw) { orc_fir_init_lp(&a->fir, 1.0, 50.0, bw, bw * 1.8, a->rate); }

int orc_am_process(orc_am* a, int n, const orc_cpx* in, double* out)
{
    /* dsp/amdemod.cpp:66-82 */
    for (int i = 0; i < n; i++) {
        double mag = sqrt(in[i].re * in[i].re + in[i].im * in[i].im);
        double z0 = mag + (a->z1 * 0.99);
        out[i] = (z0 - a->z1);
        a->z1 = z0;
    }
    orc_fir_process_real(&a->fir, n, out, out);
    return n;
}

struct orc_sam { double rate, z1, y1, phase, freq, lo_lim, hi_lim, alpha, beta; orc_fir fir; };

orc_sam* orc_sam_create(double rate)
{
    /* dsp/samdemod.cpp:54-73 */
    orc_sam* s = (orc_sam*)calloc(1, sizeof(*s));
    s->rate = rate;
    double norm = TWO_PI / rate;
    s->lo_lim = -1000.0 * norm;
    s->hi_lim = 1000.0 * norm;
    s->alpha = 2.0 * .707 * 100.0 * norm;
    s->beta = (s->alpha * s->alpha) / (4.0 * .707 * .707);
    s->fir.ntaps = 1;
    orc_fir_init_lp(&s->fir, 1.0, 40.0, 4500, 5500, rate);
    orc_fir_make_hilbert_pair(&s->fir, 5000.0);
    return s;
}
void orc_sam_destroy(orc_sam* s) { free(s); }

int orc_sam_process(orc_sam* s, int n, const orc_cpx* in, double* out)
{
    /* dsp/samdemod.cpp:78-110 */
    for (int i = 0; i < n; i++) {
        double sn = -sin(s->phase), cs = cos(s->phase);
        double tr = cs * in[i].re - sn * in[i].im;
        double ti = cs * in[i].im + sn * in[i].re;
        double err = atan2(ti, tr);
        s->freq += (s->beta * err);
        if (s->freq > s->hi_lim) s->freq = s->hi_lim;
        else if (s->freq < s->lo_lim) s->freq = s->lo_lim;
        s->phase += (s->freq + s->alpha * err);
        double z0 = tr + (s->z1 * 0.99);
        out[i] = (z0 - s->z1);
        s->z1 = z0;
    }
    s->phase = fmod(s->phase, TWO_PI);
    return n;
}

int orc_sam_process_stereo(orc_sam* s, int n, const orc_cpx* in, orc_cpx* out)
{
    /* dsp/samdemod.cpp:115-158 -- note the opposite NCO sign convention to the mono path */
    for (int i = 0; i < n; i++) {
        double sn = sin(s->phase), cs = cos(s->phase);
        double tr = cs * in[i].re - sn * in[i].im;
        double ti = cs * in[i].im + sn * in[i].re;
        double err = -atan2(ti, tr);
        s->freq += (s->beta * err);
        if (s->freq > s->hi_lim) s->freq = s->hi_lim;
        else if (s->freq < s->lo_lim) s->freq = s->lo_lim;
        s->phase += (s->freq + s->alpha * err);
        double z0 = tr + (s->z1 * 0.99);
        double y0 = ti + (s->y1 * 0.99);
        out[i].re = (z0 - s->z1);
        out[i].im = (y0 - s->y1);
        s->y1 = y0;
        s->z1 = z0;
    }
    s->phase = fmod(s->phase, TWO_PI);
    orc_fir_process_cpx(&s->fir, n, out, out);
    for (int i = 0; i < n; i++) {
        orc_cpx t = out[i];
        out[i].im = t.re - t.im;
        out[i].re = t.re + t.im;
    }
    return n;
}

#define FM_SQBUF 16384             /* MAX_SQBUF_SIZE, dsp/fmdemod.h:18 */
struct orc_fm {
    int squelched;
    double rate, hp_freq, out_gain, err_dc, dc_alpha, phase, freq, lo_lim, hi_lim, alpha, beta;
    double sq_thresh, sq_ave, sq_alpha;
    orc_fir hp;
    orc_biquad lp;
    double sqbuf[FM_SQBUF];
};

orc_fm* orc_fm_create(double rate)
{
    /* dsp/fmdemod.cpp:62-89 */
    orc_fm* f = (orc_fm*)calloc(1, sizeof(*f));
    f->rate = rate;
    double norm = TWO_PI / rate;
    f->lo_lim = -6000.0 * norm;
    f->hi_lim = 6000.0 * norm;
    f->alpha = 2.0 * .707 * 3000.0 * 2.0 * norm;
    f->beta = (f->alpha * f->alpha) / (4.0 * .707 * .707);
    f->out_gain = 25000.0 / f->hi_lim;
    f->dc_alpha = (1.0 - exp(-1.0 / (rate * 0.01)));
    f->hp_freq = 3000.0;
    f->squelched = 1;
    f->sq_alpha = (1.0 - exp(-1.0 / (rate * .02)));
    orc_biquad_init_lp(&f->lp, 3000.0, 1.0, rate);
    f->hp.ntaps = 1;
    orc_fir_init_hp(&f->hp, 1.0, 50.0, f->hp_freq, f->hp_freq * .6, rate);
    return f;
}
void orc_fm_destroy(orc_fm* f) { free(f); }
void orc_fm_set_squelch(orc_fm* f, int value) { f->sq_thresh = (double)(5000.0 - ((5000.0 * value) / 99)); }

int orc_fm_process(orc_fm* f, int n, double fm_bw, const orc_cpx* in, double* out)
{
    /* dsp/fmdemod.cpp:157-192 then PerformNoiseSquelch :113-152 */
    if (f->hp_freq != fm_bw) {
        f->hp_freq = fm_bw;
        orc_fir_init_hp(&f->hp, 1.0, 50.0, f->hp_freq, f->hp_freq * .6, f->rate);
    }
    for (int i = 0; i < n; i++) {
        double sn = sin(f->phase), cs = cos(f->phase);
        double tr = cs * in[i].re - sn * in[i].im;
        double ti = cs * in[i].im + sn * in[i].re;
        double err = -atan2(ti, tr);
        f->freq += (f->beta * err);
        if (f->freq > f->hi_lim) f->freq = f->hi_lim;
        else if (f->freq < f->lo_lim) f->freq = f->lo_lim;
        f->phase += (f->freq + f->alpha * err);
        f->err_dc = (1.0 - f->dc_alpha) * f->err_dc + f->dc_alpha * f->freq;
        out[i] = (f->freq - f->err_dc) * f->out_gain;
    }
    f->phase = fmod(f->phase, TWO_PI);
    if (n > FM_SQBUF) return n;
    orc_fir_process_real(&f->hp, n, out, f->sqbuf);
    for (int i = 0; i < n; i++) {
        double mag = fabs(f->sqbuf[i]);
        f->sq_ave = (1.0 - f->sq_alpha) * f->sq_ave + f->sq_alpha * mag;
    }
    if (0 == f->sq_thresh) f->squelched = 1;
    else if (f->squelched) { if (f->sq_ave < (f->sq_thresh - 100.0)) f->squelched = 0; }
    else { if (f->sq_ave >= (f->sq_thresh + 100.0)) f->squelched = 1; }
    if (f->squelched) for (int i = 0; i < n; i++) out[i] = 0.0;
    else orc_biquad_process(&f->lp, n, out, out);
    return n;
}

/* ======================================================================= */
/* fractional resampler                                                     */
/* ======================================================================= */
#define RS_PTS 10000
#define RS_PERIODS 28
#define RS_LEN (RS_PERIODS * RS_PTS + 1)
struct orc_resampler { double t; double* sinc; orc_cpx* buf; int cap; };

orc_resampler* orc_resampler_create(int max_input)
{
    /* dsp/fractresampler.cpp:85-116 */
    orc_resampler* r = (orc_resampler*)calloc(1, sizeof(*r));
    r->cap = max_input + RS_PERIODS;
    r->buf = (orc_cpx*)calloc(r->cap, sizeof(orc_cpx));
    r->sinc = (double*)malloc(RS_LEN * sizeof(double));
    for (int i = 0; i < RS_LEN; i++) {
        double w = (0.35875
            - 0.48829 * cos((TWO_PI * i) / (RS_LEN - 1))
            + 0.14128 * cos((2.0 * TWO_PI * i) / (RS_LEN - 1))
            - 0.01168 * cos((3.0 * TWO_PI * i) / (RS_LEN - 1)));
        double fi = ONE_PI * (double)(i - RS_LEN / 2) / (double)RS_PTS;
        r->sinc[i] = (i != RS_LEN / 2) ? w * sin(fi) / fi : 1.0;
    }
    return r;
}
void orc_resampler_destroy(orc_resampler* r) { if (r) { free(r->sinc); free(r->buf); free(r); } }
const double* orc_resampler_table(const orc_resampler* r, int* len) { if (len) *len = RS_LEN; return r->sinc; }

/* shared core, dsp/fractresampler.cpp:144-184: taps i=1..28 at table index trunc((j-t)*10000) */
static int resample_core(orc_resampler* r, int n, double rate, int cpx, orc_cpx* acc_out)
{
    int it = (int)r->t, nout = 0;
    while (it < n) {
        orc_cpx acc = {0.0, 0.0};
        for (int i = 1; i <= RS_PERIODS; i++) {
            int j = it + i;
            int s = (int)(((double)j - r->t) * (double)RS_PTS);
            acc.re += (r->buf[j].re * r->sinc[s]);
            if (cpx) acc.im += (r->buf[j].im * r->sinc[s]);
        }
        acc_out[nout++] = acc;
        r->t += rate;
        it = (int)r->t;
    }
    r->t -= (double)n;
    for (int i = 0; i < RS_PERIODS; i++) {
        if (cpx) r->buf[i] = r->buf[n + i];
        else r->buf[i].re = r->buf[n + i].re;
    }
    return nout;
}

static orc_cpx* rs_tmp(int n, double rate) { return (orc_cpx*)malloc(((size_t)(n / rate) + 64) * sizeof(orc_cpx)); }

int orc_resampler_real(orc_resampler* r, int n, double rate, const double* in, double* out)
{
    for (int i = 0; i < n; i++) r->buf[RS_PERIODS + i].re = in[i];
    orc_cpx* t = rs_tmp(n, rate);
    int m = resample_core(r, n, rate, 0, t);
    for (int i = 0; i < m; i++) out[i] = t[i].re;
    free(t);
    return m;
}

int orc_resampler_cpx(orc_resampler* r, int n, double rate, const orc_cpx* in, orc_cpx* out)
{
    for (int i = 0; i < n; i++) r->buf[RS_PERIODS + i] = in[i];
    return resample_core(r, n, rate, 1, out);
}

static short clip16(double v)
{
    if (v > 32767.0) v = 32767.0;
    if (v < -32767.0) v = -32767.0;
    return (short)v;
}

int orc_resampler_mono16(orc_resampler* r, int n, double rate, const double* in, short* out, double gain)
{
    /* dsp/fractresampler.cpp:306-352 */
    for (int i = 0; i < n; i++) r->buf[RS_PERIODS + i].re = in[i];
    orc_cpx* t = rs_tmp(n, rate);
    int m = resample_core(r, n, rate, 0, t);
    for (int i = 0; i < m; i++) out[i] = clip16(t[i].re * gain);
    free(t);
    return m;
}

int orc_resampler_stereo16(orc_resampler* r, int n, double rate, const orc_cpx* in, short* out, double gain)
{
    /* dsp/fractresampler.cpp:194-249 */
    for (int i = 0; i < n; i++) r->buf[RS_PERIODS + i] = in[i];
    orc_cpx* t = rs_tmp(n, rate);
    int m = resample_core(r, n, rate, 1, t);
    for (int i = 0; i < m; i++) { out[2 * i] = clip16(t[i].re * gain); out[2 * i + 1] = clip16(t[i].im * gain); }
    free(t);
    return m;
}

/* ======================================================================= */
/* noise blanker                                                            */
/* ======================================================================= */
#define NB_MAX_WIDTH 4096
#define NB_MAX_DELAY 4096
#define NB_MAX_AVE 2097152         /* enlarged like the _big reference build; 32768 upstream */
struct orc_blanker {
    int on;
    double threshold, width, fs, ratio, sum;
    int dptr, mptr, blank, delay_samples, mag_samples, width_samples;
    orc_cpx* delay;
    double* mag;
};

orc_blanker* orc_blanker_create(void)
{
    orc_blanker* b = (orc_blanker*)calloc(1, sizeof(*b));
    b->delay = (orc_cpx*)calloc(NB_MAX_DELAY, sizeof(orc_cpx));
    b->mag = (double*)calloc(NB_MAX_AVE, sizeof(double));
    orc_blanker_setup(b, 0, 50.0, 2.0, 1000.0);
    return b;
}
void orc_blanker_destroy(orc_blanker* b) { if (b) { free(b->delay); free(b->mag); free(b); } }

void orc_blanker_setup(orc_blanker* b, int on, double threshold, double width_us, double fs)
{
    /* dsp/noiseproc.cpp:77-119. The reference's change test compares SampleRate with itself
     * (:81), so a rate-only change is ignored; mirrored. */
    if (threshold == b->threshold && width_us == b->width && b->on == on) return;
    b->on = on; b->threshold = threshold; b->width = width_us; b->fs = fs;
    b->width_samples = width_us * 1e-6 * fs;
    if (b->width_samples < 1) b->width_samples = 1;
    else if (b->width_samples > NB_MAX_WIDTH) b->width_samples = NB_MAX_WIDTH;
    b->mag_samples = .005 * fs;
    b->ratio = .005 * (b->threshold) * (double)b->mag_samples;
    b->delay_samples = b->width_samples / 2;
    b->dptr = b->mptr = b->blank = 0;
    b->sum = 0.0;
    memset(b->delay, 0, NB_MAX_DELAY * sizeof(orc_cpx));
    memset(b->mag, 0, NB_MAX_AVE * sizeof(double));
}

void orc_blanker_process(orc_blanker* b, long n, orc_cpx* io)
{
    /* dsp/noiseproc.cpp:121-176 */
    if (!b->on) return;
    for (long i = 0; i < n; i++) {
        orc_cpx x = io[i];
        double mre = fabs(x.re), mim = fabs(x.im);
        double mag = (mre > mim) ? mre : mim;
        b->sum -= b->mag[b->mptr];
        b->sum += mag;
        b->mag[b->mptr++] = mag;
        if (b->mptr > b->mag_samples) b->mptr = 0;
        orc_cpx oldest = b->delay[b->dptr];
        b->delay[b->dptr++] = x;
        if (b->dptr > b->delay_samples) b->dptr = 0;
        if (mag * b->ratio > b->sum) b->blank = b->width_samples;
        if (b->blank) { b->blank--; io[i].re = 0.0; io[i].im = 0.0; }
        else io[i] = oldest;
    }
}

/* ======================================================================= */
/* demodulator sequencer                                                    */
/* ======================================================================= */
struct orc_demod {
    orc_downconvert* dc;
    orc_fastfir* fir;
    orc_agc* agc;
    orc_smeter* sm;
    orc_demod_info info;
    double in_rate, out_rate, max_out_bw, cw_offset;
    int mode, pos, limit;
    orc_cpx *inbuf, *tmp;
    long cap;
    orc_am* am; orc_sam* sam; orc_fm* fm;
    double* tap[5]; long tap_cap[5], tap_n[5];
};

orc_demod* orc_demod_create(void)
{
    /* dsp/demodulator.cpp:47-60 (uninitialised members behave as zero: the harness
     * zero-fills the reference object the same way) */
    orc_demod* d = (orc_demod*)calloc(1, sizeof(*d));
    d->dc = orc_downconvert_create();
    d->fir = orc_fastfir_create();
    d->agc = orc_agc_create();
    d->sm = orc_smeter_create();
    d->max_out_bw = 48000.0;
    d->out_rate = 48000.0;
    d->cap = 4000000;
    d->inbuf = (orc_cpx*)malloc(d->cap * sizeof(orc_cpx));
    d->tmp = (orc_cpx*)malloc(d->cap * sizeof(orc_cpx));
    d->limit = 1000;
    d->mode = -1;
    orc_demod_set_freq(d, 0.0);
    return d;
}

static void demod_drop(orc_demod* d)
{
    orc_am_destroy(d->am); orc_sam_destroy(d->sam); orc_fm_destroy(d->fm);
    d->am = NULL; d->sam = NULL; d->fm = NULL;
}

void orc_demod_destroy(orc_demod* d)
{
    if (!d) return;
    demod_drop(d);
    orc_downconvert_destroy(d->dc); orc_fastfir_destroy(d->fir); orc_agc_destroy(d->agc); orc_smeter_destroy(d->sm);
    free(d->inbuf); free(d->tmp); free(d);
}

void orc_demod_set_input_rate(orc_demod* d, double rate)
{
    /* dsp/demodulator.cpp:94-101 */
    if (d->in_rate != rate) {
        d->in_rate = rate;
        d->out_rate = orc_downconvert_set_data_rate(d->dc, d->in_rate, d->max_out_bw);
    }
}

void orc_demod_set_freq(orc_demod* d, double f)
{
    /* dsp/demodulator.h:68-69 */
    orc_downconvert_set_cw_offset(d->dc, d->cw_offset);
    orc_downconvert_set_frequency(d->dc, f);
}

void orc_demod_set_demod(orc_demod* d, int mode, const orc_demod_info* info)
{
    /* dsp/demodulator.cpp:107-157 */
    d->info = *info;
    if (d->mode != mode) {
        demod_drop(d);
        d->mode = mode;
        if (mode == ORC_LSB || mode == ORC_CWL) d->max_out_bw = -d->info.LowCutmin;
        else d->max_out_bw = d->info.HiCutmax;
        d->out_rate = orc_downconvert_set_data_rate(d->dc, d->in_rate, d->max_out_bw);
        if (mode == ORC_AM) d->am = orc_am_create(d->out_rate);
        else if (mode == ORC_SAM) d->sam = orc_sam_create(d->out_rate);
        else if (mode == ORC_FM) d->fm = orc_fm_create(d->out_rate);
    }
    d->cw_offset = d->info.Offset;
    orc_downconvert_set_cw_offset(d->dc, d->cw_offset);
    orc_fastfir_setup(d->fir, d->info.LowCut, d->info.HiCut, d->cw_offset, d->out_rate);
    d->limit = (d->out_rate / 100.0) * d->in_rate / d->out_rate;
    d->limit &= 0xFFFFFF00;
    orc_agc_set(d->agc, d->info.AgcOn, d->info.AgcHangOn, d->info.AgcThresh, d->info.AgcManualGain,
                d->info.AgcSlope, d->info.AgcDecay, d->out_rate);
    if (d->fm) orc_fm_set_squelch(d->fm, d->info.SquelchValue);
    if (d->am) orc_am_set_bandwidth(d->am, (d->info.HiCut - d->info.LowCut) / 2.0);
}

double orc_demod_output_rate(const orc_demod* d) { return d->out_rate; }
int orc_demod_inbuf_limit(const orc_demod* d) { return d->limit; }
/* Test aid (not in the reference): replace m_InBufLimit until the next SetDemod. The parity tests for input rates
 * whose 10 ms block is not a multiple of 2^stages run the reference algorithm on the block length the CUDA bank
 * picks (the same 10 ms rounded down to such a multiple). */
void orc_demod_set_inbuf_limit(orc_demod* d, int limit) { d->limit = limit; }
double orc_demod_smeter_peak(orc_demod* d) { return orc_smeter_peak(d->sm); }
double orc_demod_smeter_ave(const orc_demod* d) { return orc_smeter_ave(d->sm); }

void orc_demod_set_tap(orc_demod* d, int profile, double* buf, long cap_doubles)
{
    if (profile < 1 || profile > 4) return;
    d->tap[profile] = buf; d->tap_cap[profile] = cap_doubles; d->tap_n[profile] = 0;
}
long orc_demod_tap_count(const orc_demod* d, int profile) { return (profile >= 1 && profile <= 4) ? d->tap_n[profile] : 0; }

static void tap_put(orc_demod* d, int profile, const double* p, long n)
{
    if (!d->tap[profile] || n <= 0) return;
    if (d->tap_n[profile] + n <= d->tap_cap[profile]) memcpy(d->tap[profile] + d->tap_n[profile], p, n * sizeof(double));
    d->tap_n[profile] += n;
}

int orc_demod_process(orc_demod* d, int n_in, const orc_cpx* in, double* out)
{
    /* dsp/demodulator.cpp:163-215 */
    int ret = 0;
    for (int i = 0; i < n_in; i++) {
        d->inbuf[d->pos++] = in[i];
        if (d->pos >= d->limit) {
            int n = orc_downconvert_process(d->dc, d->pos, d->inbuf, d->inbuf);
            tap_put(d, 1, (double*)d->inbuf, 2L * n);
            n = orc_fastfir_process(d->fir, n, d->inbuf, d->tmp);
            tap_put(d, 2, (double*)d->tmp, 2L * n);
            orc_smeter_process(d->sm, n, d->tmp, d->out_rate);
            orc_agc_process(d->agc, n, d->tmp, d->tmp);
            tap_put(d, 3, (double*)d->tmp, 2L * n);
            switch (d->mode) {
            case ORC_AM: n = orc_am_process(d->am, n, d->tmp, out); break;
            case ORC_SAM: n = orc_sam_process(d->sam, n, d->tmp, out); break;
            case ORC_FM: n = orc_fm_process(d->fm, n, d->info.HiCut, d->tmp, out); break;
            case ORC_USB: case ORC_LSB: case ORC_CWU: case ORC_CWL:
                for (int k = 0; k < n; k++) out[k] = d->tmp[k].re;   /* dsp/ssbdemod.cpp:48-53 */
                break;
            }
            tap_put(d, 4, out, n);
            d->pos = 0;
            ret += n;
        }
    }
    return ret;
}

long orc_demod_run_c64(orc_demod* d, long n, const float* iq, int packet, double* out, long out_cap)
{
    orc_cpx* pkt = (orc_cpx*)malloc((size_t)packet * sizeof(orc_cpx));
    double* ob = (double*)malloc(16384 * sizeof(double));
    long nout = 0;
    for (long i = 0; i < n; i += packet) {
        int m = (int)((n - i) < packet ? (n - i) : packet);
        for (int k = 0; k < m; k++) { pkt[k].re = iq[2 * (i + k)]; pkt[k].im = iq[2 * (i + k) + 1]; }
        int r = orc_demod_process(d, m, pkt, ob);
        if (out && nout + r <= out_cap) memcpy(out + nout, ob, r * sizeof(double));
        nout += r;
    }
    free(pkt); free(ob);
    return nout;
}
