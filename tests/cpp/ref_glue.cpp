// Link glue for building compat_receiver.cpp against the REFERENCE's own dsp/*.cpp: the global
// test-bench pointer the DSP code dereferences and the (commented-out everywhere) perf hooks.
#include "gui/testbench.h"
#include "interface/perform.h"
static thread_local CTestBench t_bench;
thread_local CTestBench* g_pTestBench = &t_bench;
void InitPerformance() {}
void StartPerformance() {}
void StopPerformance(int) {}
void ReadPerformance() {}
void SamplePerformance() {}
int GetDeltaPerformance() { return 0; }
