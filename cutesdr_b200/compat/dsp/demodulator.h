// Drop-in for dsp/demodulator.h:56-100: one receiver = a bank of one. interface/sdrinterface.cpp keeps
// calling SetInputSampleRate / SetDemod / SetDemodFreq / ProcessData exactly as before.
#ifndef CUTESDR_B200_COMPAT_DEMODULATOR_H
#define CUTESDR_B200_COMPAT_DEMODULATOR_H
#include "dsp/datatypes.h"
#include "dsp/cutesdr_shim.h"
#include "dsp/iir.h"      // the reference pulls CIir in here (demodulator.h -> fmdemod.h -> iir.h); interface/sdrinterface.h:178 relies on it
#define DEMOD_AM 0
#define DEMOD_SAM 1
#define DEMOD_FM 2
#define DEMOD_USB 3
#define DEMOD_LSB 4
#define DEMOD_CWU 5
#define DEMOD_CWL 6
#define NUM_DEMODS 7
#define MAX_INBUFSIZE 250000
#define MAX_MAGBUFSIZE 32000
#ifdef QT_VERSION
#include <QString>
#endif
typedef struct _sdmd {
    int HiCut; int HiCutmin; int HiCutmax;
    int LowCut; int LowCutmin; int LowCutmax;
    int FilterClickResolution;
    int Offset; int SquelchValue;
    int AgcSlope; int AgcThresh; int AgcManualGain; int AgcDecay;
    bool AgcOn; bool AgcHangOn; bool Symetric;
#ifdef QT_VERSION
    QString txt;
#endif
} tDemodInfo;
class CDemodulator {
public:
    CDemodulator() : m_h(0) { cutesdr_shim_check(cutesdr_demodulator_create(&m_h, CUTESDR_DEVICE), "CDemodulator()"); }
    virtual ~CDemodulator() { cutesdr_demodulator_destroy(m_h); }
    void SetInputSampleRate(TYPEREAL InputRate) { cutesdr_shim_check(cutesdr_demodulator_set_input_sample_rate(m_h, InputRate), "SetInputSampleRate"); }
    double GetOutputRate() { double r = 0; cutesdr_shim_check(cutesdr_demodulator_get_output_rate(m_h, &r), "GetOutputRate"); return r; }
    double GetSMeterPeak() { double p = 0; cutesdr_shim_check(cutesdr_demodulator_get_smeter(m_h, &p, 0), "GetSMeterPeak"); return p; }
    double GetSMeterAve() { double a = 0; cutesdr_shim_check(cutesdr_demodulator_get_smeter(m_h, 0, &a), "GetSMeterAve"); return a; }   // does not reset the held peak
    void SetDemod(int Mode, tDemodInfo CurrentDemodInfo)
    {
        cutesdr_demod_info d;
        d.HiCut = CurrentDemodInfo.HiCut; d.HiCutmin = CurrentDemodInfo.HiCutmin; d.HiCutmax = CurrentDemodInfo.HiCutmax;
        d.LowCut = CurrentDemodInfo.LowCut; d.LowCutmin = CurrentDemodInfo.LowCutmin; d.LowCutmax = CurrentDemodInfo.LowCutmax;
        d.Offset = CurrentDemodInfo.Offset; d.SquelchValue = CurrentDemodInfo.SquelchValue;
        d.AgcSlope = CurrentDemodInfo.AgcSlope; d.AgcThresh = CurrentDemodInfo.AgcThresh;
        d.AgcManualGain = CurrentDemodInfo.AgcManualGain; d.AgcDecay = CurrentDemodInfo.AgcDecay;
        d.AgcOn = CurrentDemodInfo.AgcOn ? 1 : 0; d.AgcHangOn = CurrentDemodInfo.AgcHangOn ? 1 : 0;
        cutesdr_shim_check(cutesdr_demodulator_set_demod(m_h, Mode, &d), "SetDemod");
    }
    void SetDemodFreq(TYPEREAL Freq) { cutesdr_shim_check(cutesdr_demodulator_set_demod_freq(m_h, Freq), "SetDemodFreq"); }
    int ProcessData(int InLength, TYPECPX* pInData, TYPEREAL* pOutData)
    {
        int n = cutesdr_shim_check(cutesdr_demodulator_process(m_h, InLength, (const double*)pInData, pOutData), "CDemodulator::ProcessData");
        return n < 0 ? 0 : n;
    }
    int ProcessData(int InLength, TYPECPX* pInData, TYPECPX* pOutData)
    {
        int n = cutesdr_shim_check(cutesdr_demodulator_process_stereo(m_h, InLength, (const double*)pInData, (double*)pOutData), "CDemodulator::ProcessData(stereo)");
        return n < 0 ? 0 : n;
    }
private:
    CDemodulator(const CDemodulator&);
    CDemodulator& operator=(const CDemodulator&);
    cutesdr_demodulator* m_h;
};
#endif
