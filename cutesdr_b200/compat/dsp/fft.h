// Drop-in for dsp/fft.h:24-85.
#ifndef CUTESDR_B200_COMPAT_FFT_H
#define CUTESDR_B200_COMPAT_FFT_H
#include "dsp/datatypes.h"
#include "dsp/cutesdr_shim.h"
#define MAX_FFT_SIZE 65536
#define MIN_FFT_SIZE 512
class CFft {
public:
    CFft() : m_h(0) { cutesdr_shim_check(cutesdr_fft_create(&m_h, CUTESDR_DEVICE), "CFft()"); }
    virtual ~CFft() { cutesdr_fft_destroy(m_h); }
    void SetFFTParams(qint32 size, bool invert, double dBCompensation, double SampleFreq)
    {
        cutesdr_shim_check(cutesdr_fft_set_params(m_h, size, invert ? 1 : 0, dBCompensation, SampleFreq), "SetFFTParams");
    }
    void SetFFTAve(qint32 ave) { cutesdr_shim_check(cutesdr_fft_set_ave(m_h, ave), "SetFFTAve"); }
    void ResetFFT() { cutesdr_shim_check(cutesdr_fft_reset(m_h), "ResetFFT"); }
    bool GetScreenIntegerFFTData(qint32 MaxHeight, qint32 MaxWidth, double MaxdB, double MindB, qint32 StartFreq,
                                 qint32 StopFreq, qint32* OutBuf)
    {
        int ov = 0;
        cutesdr_shim_check(cutesdr_fft_get_screen(m_h, MaxHeight, MaxWidth, MaxdB, MindB, StartFreq, StopFreq, OutBuf, &ov), "GetScreenIntegerFFTData");
        return ov != 0;
    }
    qint32 PutInDisplayFFT(qint32 n, TYPECPX* InBuf)
    {
        int total = 0;
        cutesdr_shim_check(cutesdr_fft_put(m_h, n, (const double*)InBuf, &total), "PutInDisplayFFT");
        return total;
    }
    void FwdFFT(TYPECPX* pInOutBuf) { cutesdr_shim_check(cutesdr_fft_fwd(m_h, (double*)pInOutBuf), "FwdFFT"); }
    void RevFFT(TYPECPX* pInOutBuf) { cutesdr_shim_check(cutesdr_fft_rev(m_h, (double*)pInOutBuf), "RevFFT"); }
private:
    CFft(const CFft&);
    CFft& operator=(const CFft&);
    cutesdr_fft* m_h;
};
#endif
