// Drop-in for dsp/noiseproc.h:23-53.
#ifndef CUTESDR_B200_COMPAT_NOISEPROC_H
#define CUTESDR_B200_COMPAT_NOISEPROC_H
#include "dsp/datatypes.h"
#include "dsp/cutesdr_shim.h"
typedef struct _snproc { bool NBOn; int NBThreshold; int NBWidth; } tNoiseProcdInfo;
class CNoiseProc {
public:
    CNoiseProc() : m_h(0) { cutesdr_shim_check(cutesdr_noiseproc_create(&m_h, CUTESDR_DEVICE), "CNoiseProc()"); }
    virtual ~CNoiseProc() { cutesdr_noiseproc_destroy(m_h); }
    void SetupBlanker(bool On, TYPEREAL Threshold, TYPEREAL Width, TYPEREAL SampleRate)
    {
        cutesdr_shim_check(cutesdr_noiseproc_setup(m_h, On, Threshold, Width, SampleRate), "SetupBlanker");
    }
    void ProcessBlanker(int InLength, TYPECPX* pInData, TYPECPX* pOutData)
    {
        cutesdr_shim_check(cutesdr_noiseproc_process(m_h, InLength, (const double*)pInData, (double*)pOutData), "ProcessBlanker");
    }
private:
    CNoiseProc(const CNoiseProc&);
    CNoiseProc& operator=(const CNoiseProc&);
    cutesdr_noiseproc* m_h;
};
#endif
