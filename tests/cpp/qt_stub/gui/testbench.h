// Stand-in for gui/testbench.h (a QDialog in the reference): only what interface/sdrinterface.cpp touches.
// TEST INFRASTRUCTURE ONLY, see ../qt_stub.h.
#ifndef CUTESDR_B200_TESTBENCH_STUB_H
#define CUTESDR_B200_TESTBENCH_STUB_H
#include "dsp/datatypes.h"
#define PROFILE_OFF 0
#define PROFILE_1 1
#define PROFILE_2 2
#define PROFILE_3 3
#define PROFILE_4 4
#define PROFILE_5 5
#define PROFILE_6 6
#define PROFILE_7 7
class CTestBench {
public:
    void CreateGeneratorSamples(int, TYPECPX*, double) {}
    void CreateGeneratorSamples(int, TYPEREAL*, double) {}
    void DisplayData(int, TYPEREAL*, double, int) {}
    void DisplayData(int, TYPECPX*, double, int) {}
    void DisplayData(int, TYPEMONO16*, double, int) {}
    void DisplayData(int, TYPESTEREO16*, double, int) {}
    void SendDebugTxt(QString) {}
};
extern CTestBench* g_pTestBench;
#endif
