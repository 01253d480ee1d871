// Drop-in for dsp/fractresampler.h:16-33.
#ifndef CUTESDR_B200_COMPAT_FRACTRESAMPLER_H
#define CUTESDR_B200_COMPAT_FRACTRESAMPLER_H
#include "dsp/datatypes.h"
#include "dsp/cutesdr_shim.h"
class CFractResampler {
public:
    CFractResampler() : m_h(0) { cutesdr_shim_check(cutesdr_resampler_create(&m_h, CUTESDR_DEVICE), "CFractResampler()"); }
    virtual ~CFractResampler() { cutesdr_resampler_destroy(m_h); }
    void Init(int MaxInputSize) { cutesdr_shim_check(cutesdr_resampler_init(m_h, MaxInputSize), "CFractResampler::Init"); }
    int Resample(int InLength, TYPEREAL Rate, TYPEREAL* pInBuf, TYPEREAL* pOutBuf)
    {
        int n = cutesdr_shim_check(cutesdr_resampler_real(m_h, InLength, Rate, pInBuf, pOutBuf), "Resample");
        return n < 0 ? 0 : n;
    }
    int Resample(int InLength, TYPEREAL Rate, TYPECPX* pInBuf, TYPECPX* pOutBuf)
    {
        int n = cutesdr_shim_check(cutesdr_resampler_cpx(m_h, InLength, Rate, (const double*)pInBuf, (double*)pOutBuf), "Resample");
        return n < 0 ? 0 : n;
    }
    int Resample(int InLength, TYPEREAL Rate, TYPEREAL* pInBuf, TYPEMONO16* pOutBuf, TYPEREAL gain)
    {
        int n = cutesdr_shim_check(cutesdr_resampler_mono16(m_h, InLength, Rate, pInBuf, (int16_t*)pOutBuf, gain), "Resample");
        return n < 0 ? 0 : n;
    }
    int Resample(int InLength, TYPEREAL Rate, TYPECPX* pInBuf, TYPESTEREO16* pOutBuf, TYPEREAL gain)
    {
        int n = cutesdr_shim_check(cutesdr_resampler_stereo16(m_h, InLength, Rate, (const double*)pInBuf, (int16_t*)pOutBuf, gain), "Resample");
        return n < 0 ? 0 : n;
    }
private:
    CFractResampler(const CFractResampler&);
    CFractResampler& operator=(const CFractResampler&);
    cutesdr_resampler* m_h;
};
#endif
