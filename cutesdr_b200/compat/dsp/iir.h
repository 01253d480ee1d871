// Drop-in for dsp/iir.h:16-40 (CIir): same class name and methods, forwarding to libcutesdr_cuda.
// interface/sdrinterface.h:178 embeds one by value (it arrived there through dsp/demodulator.h -> dsp/fmdemod.h ->
// dsp/iir.h in the reference, so the compat demodulator.h includes this header as well).
#ifndef CUTESDR_B200_COMPAT_IIR_H
#define CUTESDR_B200_COMPAT_IIR_H
#include "dsp/datatypes.h"
#include "dsp/cutesdr_shim.h"
class CIir {
public:
    CIir() : m_h(0) { cutesdr_shim_check(cutesdr_iir_create(&m_h, CUTESDR_DEVICE), "CIir()"); }
    ~CIir() { cutesdr_iir_destroy(m_h); }
    void InitLP(TYPEREAL F0Freq, TYPEREAL FilterQ, TYPEREAL SampleRate) { cutesdr_shim_check(cutesdr_iir_init(m_h, CUTESDR_IIR_LP, F0Freq, FilterQ, SampleRate), "CIir::InitLP"); }
    void InitHP(TYPEREAL F0Freq, TYPEREAL FilterQ, TYPEREAL SampleRate) { cutesdr_shim_check(cutesdr_iir_init(m_h, CUTESDR_IIR_HP, F0Freq, FilterQ, SampleRate), "CIir::InitHP"); }
    void InitBP(TYPEREAL F0Freq, TYPEREAL FilterQ, TYPEREAL SampleRate) { cutesdr_shim_check(cutesdr_iir_init(m_h, CUTESDR_IIR_BP, F0Freq, FilterQ, SampleRate), "CIir::InitBP"); }
    void InitBR(TYPEREAL F0Freq, TYPEREAL FilterQ, TYPEREAL SampleRate) { cutesdr_shim_check(cutesdr_iir_init(m_h, CUTESDR_IIR_BR, F0Freq, FilterQ, SampleRate), "CIir::InitBR"); }
    void ProcessFilter(int InLength, TYPEREAL* InBuf, TYPEREAL* OutBuf) { cutesdr_shim_check(cutesdr_iir_process_real(m_h, InLength, InBuf, OutBuf), "CIir::ProcessFilter"); }
    void ProcessFilter(int InLength, TYPECPX* InBuf, TYPECPX* OutBuf) { cutesdr_shim_check(cutesdr_iir_process_cpx(m_h, InLength, (const double*)InBuf, (double*)OutBuf), "CIir::ProcessFilter"); }
private:
    CIir(const CIir&);
    CIir& operator=(const CIir&);
    cutesdr_iir* m_h;
};
#endif
