// Drop-in for dsp/downconvert.h:23-120 -- same class name and public methods.
#ifndef CUTESDR_B200_COMPAT_DOWNCONVERT_H
#define CUTESDR_B200_COMPAT_DOWNCONVERT_H
#include "dsp/datatypes.h"
#include "dsp/cutesdr_shim.h"
class CDownConvert {
public:
    CDownConvert() : m_h(0) { cutesdr_shim_check(cutesdr_downconvert_create(&m_h, CUTESDR_DEVICE), "CDownConvert()"); }
    virtual ~CDownConvert() { cutesdr_downconvert_destroy(m_h); }
    void SetFrequency(TYPEREAL NcoFreq) { cutesdr_shim_check(cutesdr_downconvert_set_frequency(m_h, NcoFreq), "SetFrequency"); }
    void SetCwOffset(TYPEREAL offset) { cutesdr_shim_check(cutesdr_downconvert_set_cw_offset(m_h, offset), "SetCwOffset"); }
    int ProcessData(int InLength, TYPECPX* pInData, TYPECPX* pOutData)
    {
        int n = cutesdr_shim_check(cutesdr_downconvert_process(m_h, InLength, (const double*)pInData, (double*)pOutData), "CDownConvert::ProcessData");
        return n < 0 ? 0 : n;
    }
    TYPEREAL SetDataRate(TYPEREAL InRate, TYPEREAL MaxBW)
    {
        double r = 0.0;
        cutesdr_shim_check(cutesdr_downconvert_set_data_rate(m_h, InRate, MaxBW, &r), "SetDataRate");
        return r;
    }
private:
    CDownConvert(const CDownConvert&);
    CDownConvert& operator=(const CDownConvert&);
    cutesdr_downconvert* m_h;
};
#endif
