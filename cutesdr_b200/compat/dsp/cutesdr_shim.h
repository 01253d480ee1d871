// Shared helpers of the drop-in class shims: every method forwards to libcutesdr_cuda's C ABI
// (include/cutesdr_cuda.h). Like the reference's classes they report no errors to the caller;
// a failed call logs cutesdr_last_error() to stderr and, for ProcessData, returns 0 samples.
#ifndef CUTESDR_B200_COMPAT_SHIM_H
#define CUTESDR_B200_COMPAT_SHIM_H
#include <stdio.h>
#include "cutesdr_cuda.h"
#ifndef CUTESDR_DEVICE
#define CUTESDR_DEVICE 0
#endif
static inline int cutesdr_shim_check(int rc, const char* what)
{
    if (rc < 0) fprintf(stderr, "libcutesdr_cuda: %s failed (%d): %s\n", what, rc, cutesdr_last_error());
    return rc;
}
#endif
