// Drop-in for dsp/fastfir.h:19-44.
#ifndef CUTESDR_B200_COMPAT_FASTFIR_H
#define CUTESDR_B200_COMPAT_FASTFIR_H
#include "dsp/datatypes.h"
#include "dsp/cutesdr_shim.h"
class CFastFIR {
public:
    CFastFIR() : m_h(0) { cutesdr_shim_check(cutesdr_fastfir_create(&m_h, CUTESDR_DEVICE), "CFastFIR()"); }
    virtual ~CFastFIR() { cutesdr_fastfir_destroy(m_h); }
    void SetupParameters(TYPEREAL FLoCut, TYPEREAL FHiCut, TYPEREAL Offset, TYPEREAL SampleRate)
    {
        cutesdr_shim_check(cutesdr_fastfir_setup(m_h, FLoCut, FHiCut, Offset, SampleRate), "SetupParameters");
    }
    int ProcessData(int InLength, TYPECPX* InBuf, TYPECPX* OutBuf)
    {
        int n = cutesdr_shim_check(cutesdr_fastfir_process(m_h, InLength, (const double*)InBuf, (double*)OutBuf), "CFastFIR::ProcessData");
        return n < 0 ? 0 : n;
    }
private:
    CFastFIR(const CFastFIR&);
    CFastFIR& operator=(const CFastFIR&);
    cutesdr_fastfir* m_h;
};
#endif
