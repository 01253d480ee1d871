// Host-side drop-in check: this file is written against the REFERENCE's class API only
// (CNoiseProc, CFft, CDemodulator, CFractResampler with the reference's method names and argument
// order) and mirrors the one call site that drives the chain in the reference,
// CSdrInterface::ProcessIQData (interface/sdrinterface.cpp:878-922) followed by
// CSoundOut::PutOutQueue's resampler call (interface/soundout.cpp:262). Compiled against
// cutesdr_b200/compat (the shims over libcutesdr_cuda) it is the GPU receiver; the same source
// compiles against the reference's own dsp/ headers.
//
//   compat_receiver <iq.c64> <fs> <mode> <lo> <hi> <freq> <fftsize> <out.bin>
// out.bin: int32 n_audio, double audio[n_audio], int32 n48, int16 audio48[n48], int32 w, int32 screen[w], int32 overload
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "dsp/demodulator.h"
#include "dsp/noiseproc.h"
#include "dsp/fft.h"
#include "dsp/fractresampler.h"

int main(int argc, char** argv)
{
    if (argc < 9) { fprintf(stderr, "usage\n"); return 2; }
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror("iq"); return 2; }
    fseek(f, 0, SEEK_END);
    long bytes = ftell(f);
    fseek(f, 0, SEEK_SET);
    long n = bytes / 8;
    std::vector<float> raw(2 * n);
    if (fread(raw.data(), 8, n, f) != (size_t)n) return 2;
    fclose(f);
    const double fs = atof(argv[2]);
    const int mode = atoi(argv[3]);
    const int lo = atoi(argv[4]), hi = atoi(argv[5]);
    const double freq = atof(argv[6]);
    const int fftsize = atoi(argv[7]);

    // static storage = zero-initialised before construction: the reference's CDemodulator leaves
    // m_pFmDemod (and others) uninitialised and deletes it in SetDemod (dsp/demodulator.cpp:47-60,80-81)
    static CNoiseProc m_NoiseProc;
    static CFft m_Fft;
    static CDemodulator m_Demodulator;
    static CFractResampler m_OutResampler;

    m_NoiseProc.SetupBlanker(false, 50.0, 2.0, fs);
    m_Fft.SetFFTParams(fftsize, false, 0.0, fs);
    m_Fft.SetFFTAve(2);
    m_Demodulator.SetInputSampleRate(fs);
    tDemodInfo info;
    memset(&info, 0, sizeof(info));
    info.HiCut = hi; info.LowCut = lo;
    info.HiCutmin = 500; info.HiCutmax = (mode == DEMOD_FM) ? 15000 : ((mode >= DEMOD_USB) ? 20000 : 10000);
    info.LowCutmax = -500; info.LowCutmin = (mode == DEMOD_FM) ? -15000 : ((mode >= DEMOD_USB) ? -20000 : -10000);
    info.AgcSlope = 0; info.AgcThresh = -100; info.AgcManualGain = 30; info.AgcDecay = 200;
    info.AgcOn = true; info.AgcHangOn = false;
    m_Demodulator.SetDemod(mode, info);
    m_Demodulator.SetDemodFreq(freq);
    m_OutResampler.Init(8192);
    const double rate = m_Demodulator.GetOutputRate();

    std::vector<double> audio;
    std::vector<short> audio48;
    std::vector<TYPECPX> pkt(256), fftbuf(fftsize);
    std::vector<double> snd(8192);
    std::vector<short> snd16(16384);
    int fftpos = 0;
    for (long i = 0; i < n; i += 256) {
        int m = (int)((n - i) < 256 ? (n - i) : 256);
        for (int k = 0; k < m; k++) { pkt[k].re = raw[2 * (i + k)]; pkt[k].im = raw[2 * (i + k) + 1]; }
        m_NoiseProc.ProcessBlanker(m, pkt.data(), pkt.data());
        for (int k = 0; k < m; k++) {
            fftbuf[fftpos++] = pkt[k];
            if (fftpos >= fftsize) { fftpos = 0; m_Fft.PutInDisplayFFT(fftsize, fftbuf.data()); }
        }
        int got = m_Demodulator.ProcessData(m, pkt.data(), snd.data());
        if (got > 0) {
            audio.insert(audio.end(), snd.begin(), snd.begin() + got);
            int r = m_OutResampler.Resample(got, rate / 48000.0, snd.data(), (TYPEMONO16*)snd16.data(), 0.5);
            audio48.insert(audio48.end(), snd16.begin(), snd16.begin() + r);
        }
    }
    const int W = 800;
    std::vector<qint32> screen(W);
    bool ov = m_Fft.GetScreenIntegerFFTData(255, W, 0.0, -140.0, (qint32)(-fs / 2), (qint32)(fs / 2), screen.data());

    FILE* o = fopen(argv[8], "wb");
    if (!o) { perror("out"); return 2; }
    int na = (int)audio.size(), n48 = (int)audio48.size(), w = W, iov = ov ? 1 : 0;
    fwrite(&na, 4, 1, o); fwrite(audio.data(), 8, na, o);
    fwrite(&n48, 4, 1, o); fwrite(audio48.data(), 2, n48, o);
    fwrite(&w, 4, 1, o); fwrite(screen.data(), 4, w, o);
    fwrite(&iov, 4, 1, o);
    fclose(o);
    printf("audio %d samples @ %.1f Hz, %d @ 48k, overload %d\n", na, rate, n48, iov);
    return 0;
}
