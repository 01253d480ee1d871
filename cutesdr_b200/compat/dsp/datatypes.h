// Drop-in replacement for the reference's dsp/datatypes.h (types only; no Qt needed).
#ifndef CUTESDR_B200_COMPAT_DATATYPES_H
#define CUTESDR_B200_COMPAT_DATATYPES_H
#include <math.h>
#include <stdint.h>
#ifndef QT_VERSION
typedef int8_t qint8;
typedef int16_t qint16;
typedef int32_t qint32;
typedef int64_t qint64;
#endif
typedef float tSReal;
typedef double tDReal;
typedef struct _sCplx { tSReal re; tSReal im; } tSComplex;
typedef struct _dCplx { tDReal re; tDReal im; } tDComplex;
typedef struct _isCplx { qint16 re; qint16 im; } tStereo16;
#define TYPEREAL tDReal
#define TYPECPX tDComplex
#define TYPESTEREO16 tStereo16
#define TYPEMONO16 qint16
#define K_2PI (2.0 * 3.14159265358979323846)
#define K_PI (3.14159265358979323846)
#endif
