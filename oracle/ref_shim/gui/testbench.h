// Stand-in for the reference's gui/testbench.h: the DSP code calls
// g_pTestBench->DisplayData(...) unconditionally at its named tap points
// (PROFILE_1..7). Here the taps are captured into buffers the harness reads.
// TEST INFRASTRUCTURE ONLY.
#ifndef CUTESDR_B200_TESTBENCH_SHIM_H
#define CUTESDR_B200_TESTBENCH_SHIM_H
#include "dsp/datatypes.h"
#include <vector>

#define PROFILE_OFF 0
#define PROFILE_1 1
#define PROFILE_2 2
#define PROFILE_3 3
#define PROFILE_4 4
#define PROFILE_5 5
#define PROFILE_6 6
#define PROFILE_7 7
#define NUM_PROFILES 8

class CTestBench {
public:
    CTestBench() : m_CaptureMask(0) {}
    void CreateGeneratorSamples(int, TYPECPX*, double) {}
    void CreateGeneratorSamples(int, TYPEREAL*, double) {}
    void DisplayData(int n, TYPEREAL* p, double, int profile) {
        if (!(m_CaptureMask & (1u << profile)) || n <= 0) return;
        std::vector<double>& v = m_Tap[profile];
        v.insert(v.end(), p, p + n);
    }
    void DisplayData(int n, TYPECPX* p, double, int profile) {
        if (!(m_CaptureMask & (1u << profile)) || n <= 0) return;
        std::vector<double>& v = m_Tap[profile];
        v.insert(v.end(), (double*)p, (double*)p + 2 * n);
    }
    void DisplayData(int, TYPEMONO16*, double, int) {}
    void DisplayData(int, TYPESTEREO16*, double, int) {}
    void SendDebugTxt(QString) {}
    unsigned m_CaptureMask;
    std::vector<double> m_Tap[NUM_PROFILES];
};
extern thread_local CTestBench* g_pTestBench;
#endif
