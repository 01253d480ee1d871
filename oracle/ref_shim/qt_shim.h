// Minimal stand-in for the handful of Qt names the reference's dsp/*.cpp touch
// (QtGui/QApplication, QMutex, QDebug, QDir, QFile, QString, qintN, TRUE/FALSE).
// TEST INFRASTRUCTURE ONLY: lets oracle/Makefile compile the *unmodified*
// reference DSP sources headless (no Qt in this image). Nothing here is shipped.
#ifndef CUTESDR_B200_QT_SHIM_H
#define CUTESDR_B200_QT_SHIM_H
#include <stdint.h>
#include <string>

typedef int8_t qint8;
typedef uint8_t quint8;
typedef int16_t qint16;
typedef uint16_t quint16;
typedef int32_t qint32;
typedef uint32_t quint32;
typedef int64_t qint64;
typedef uint64_t quint64;

#ifndef TRUE
#define TRUE true
#endif
#ifndef FALSE
#define FALSE false
#endif

class QMutex {
public:
    void lock() {}
    void unlock() {}
};

class QString {
public:
    QString() {}
    QString(const char* s) : m_s(s ? s : "") {}
    std::string m_s;
};

// qDebug() << anything  -> swallowed
struct QDebugSink {
    template <typename T> QDebugSink& operator<<(const T&) { return *this; }
};
inline QDebugSink qDebug() { return QDebugSink(); }

struct QIODevice { enum OpenModeFlag { WriteOnly = 2 }; };
class QDir { public: static bool setCurrent(const char*) { return false; } };
class QFile {
public:
    void setFileName(const char*) {}
    bool open(int) { return false; }
    void write(const char*) {}
};
#endif
