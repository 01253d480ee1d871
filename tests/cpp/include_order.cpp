// Compile-only check of the drop-in headers: the dsp/ includes of interface/sdrinterface.h:13-16 and
// interface/soundout.h:16 in the reference's order, and the by-value members CSdrInterface / CSoundOut declare
// (interface/sdrinterface.h:173-178, interface/soundout.h:53), CIir included -- it must arrive through
// dsp/demodulator.h exactly as it does in the reference (demodulator.h -> fmdemod.h -> iir.h).
#include "dsp/fft.h"
#include "dsp/demodulator.h"
#include "dsp/noiseproc.h"
#include "dsp/fractresampler.h"

struct EmbedsLikeCSdrInterface {
    CFft m_Fft;
    CDemodulator m_Demodulator;
    CNoiseProc m_NoiseProc;
    CIir m_Iir;
    CFractResampler m_OutResampler;
};

int include_order_check(TYPEREAL* r, TYPECPX* c)
{
    EmbedsLikeCSdrInterface* s = 0;
    s->m_Iir.InitLP(3000.0, 1.0, 48000.0);
    s->m_Iir.InitHP(300.0, 1.0, 48000.0);
    s->m_Iir.InitBP(1000.0, 2.0, 48000.0);
    s->m_Iir.InitBR(25000.0, 1000.0, 100000.0);
    s->m_Iir.ProcessFilter(16, r, r);
    s->m_Iir.ProcessFilter(16, c, c);
    return 0;
}
