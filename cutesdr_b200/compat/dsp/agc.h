// Drop-in for dsp/agc.h:18-63 (the real-data overload has no caller in the reference and is omitted).
#ifndef CUTESDR_B200_COMPAT_AGC_H
#define CUTESDR_B200_COMPAT_AGC_H
#include "dsp/datatypes.h"
#include "dsp/cutesdr_shim.h"
class CAgc {
public:
    CAgc() : m_h(0) { cutesdr_shim_check(cutesdr_agc_create(&m_h, CUTESDR_DEVICE), "CAgc()"); }
    virtual ~CAgc() { cutesdr_agc_destroy(m_h); }
    void SetParameters(bool AgcOn, bool UseHang, int Threshold, int ManualGain, int Slope, int Decay, TYPEREAL SampleRate)
    {
        cutesdr_shim_check(cutesdr_agc_set_parameters(m_h, AgcOn, UseHang, Threshold, ManualGain, Slope, Decay, SampleRate), "CAgc::SetParameters");
    }
    void ProcessData(int Length, TYPECPX* pInData, TYPECPX* pOutData)
    {
        cutesdr_shim_check(cutesdr_agc_process(m_h, Length, (const double*)pInData, (double*)pOutData), "CAgc::ProcessData");
    }
private:
    CAgc(const CAgc&);
    CAgc& operator=(const CAgc&);
    cutesdr_agc* m_h;
};
#endif
