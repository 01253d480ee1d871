// Headless stand-in for the Qt names that the reference's interface/sdrinterface.{h,cpp}, netiobase.h, soundout.h and
// ad6620.h mention. TEST INFRASTRUCTURE ONLY (tests/cpp/Makefile `interface_syntax`): it lets g++ -fsyntax-only push
// the UNMODIFIED interface/sdrinterface.cpp through the compat dsp/*.h headers, which is the drop-in claim of
// the file that embeds CFft, CDemodulator, CNoiseProc, CIir and (via CSoundOut) CFractResampler by value
// compiles against libcutesdr_cuda's class headers without an edit. Nothing here is linked or shipped.
#ifndef CUTESDR_B200_QT_STUB_H
#define CUTESDR_B200_QT_STUB_H
#include <stdint.h>
#include <string>
#include <vector>
typedef int8_t qint8; typedef uint8_t quint8; typedef int16_t qint16; typedef uint16_t quint16;
typedef int32_t qint32; typedef uint32_t quint32; typedef int64_t qint64; typedef uint64_t quint64;
typedef double qreal;
#ifndef TRUE
#define TRUE true
#endif
#ifndef FALSE
#define FALSE false
#endif
#define Q_OBJECT
#define signals public
#define slots
#define emit
#define SIGNAL(x) #x
#define SLOT(x) #x
class QString {
public:
    QString() {}
    QString(const char* s) : m_s(s ? s : "") {}
    QString& operator+=(const QString& o) { m_s += o.m_s; return *this; }
    QString operator+(const QString& o) const { QString r(*this); r += o; return r; }
    bool operator==(const QString& o) const { return m_s == o.m_s; }
    bool operator!=(const QString& o) const { return m_s != o.m_s; }
    static QString number(double, char = 'g', int = 6) { return QString(); }
    static QString number(int, int = 10) { return QString(); }
    QString arg(double) const { return *this; }
    QString arg(const QString&) const { return *this; }
    int toInt() const { return 0; }
    int length() const { return (int)m_s.size(); }
    std::string m_s;
};
inline bool operator==(const char* a, const QString& b) { return b == QString(a); }
inline bool operator!=(const char* a, const QString& b) { return b != QString(a); }
struct QDebugSink { template <typename T> QDebugSink& operator<<(const T&) { return *this; } };
inline QDebugSink qDebug() { return QDebugSink(); }
class QMutex { public: void lock() {} void unlock() {} bool tryLock() { return true; } };
class QWaitCondition { public: bool wait(QMutex*, unsigned long = 0) { return true; } void wakeAll() {} void wakeOne() {} };
class QObject {
public:
    QObject(QObject* = 0) {}
    virtual ~QObject() {}
    static bool connect(const void*, const char*, const void*, const char*, int = 0) { return true; }
    static bool disconnect(const void*, const char*, const void*, const char*) { return true; }
    void moveToThread(void*) {}
    void deleteLater() {}
    QObject* parent() const { return 0; }
};
namespace Qt { enum ConnectionType { AutoConnection, DirectConnection, QueuedConnection, BlockingQueuedConnection }; }
class QThread : public QObject {
public:
    enum Priority { IdlePriority, LowestPriority, LowPriority, NormalPriority, HighPriority, HighestPriority, TimeCriticalPriority, InheritPriority };
    QThread(QObject* p = 0) : QObject(p) {}
    void start(Priority = InheritPriority) {}
    void setPriority(Priority) {}
    void quit() {}
    void exit(int = 0) {}
    bool wait(unsigned long = ~0ul) { return true; }
    bool isRunning() const { return false; }
    int exec() { return 0; }
    static void msleep(unsigned long) {}
    static void usleep(unsigned long) {}
    static QThread* currentThread() { return 0; }
protected:
    virtual void run() {}
};
class QHostAddress {
public:
    enum SpecialAddress { Any, LocalHost, Broadcast };
    QHostAddress() {}
    QHostAddress(SpecialAddress) {}
    QHostAddress(quint32) {}
    QHostAddress(const QString&) {}
    quint32 toIPv4Address() const { return 0; }
    QString toString() const { return QString(); }
    bool operator==(const QHostAddress&) const { return true; }
    bool operator!=(const QHostAddress&) const { return false; }
};
class QIODevice : public QObject {
public:
    enum OpenModeFlag { NotOpen = 0, ReadOnly = 1, WriteOnly = 2, ReadWrite = 3, Append = 4, Truncate = 8, Text = 16, Unbuffered = 32 };
    QIODevice(QObject* p = 0) : QObject(p) {}
    virtual qint64 read(char*, qint64) { return 0; }
    virtual qint64 write(const char*, qint64) { return 0; }
    qint64 write(const char*) { return 0; }
    virtual qint64 bytesAvailable() const { return 0; }
    virtual bool open(int) { return true; }
    virtual void close() {}
    bool isOpen() const { return true; }
};
class QAbstractSocket : public QIODevice {
public:
    enum SocketState { UnconnectedState, HostLookupState, ConnectingState, ConnectedState, BoundState, ListeningState, ClosingState };
    enum SocketError { ConnectionRefusedError, RemoteHostClosedError, HostNotFoundError, SocketAccessError };
    enum SocketOption { LowDelayOption, KeepAliveOption };
    QAbstractSocket(QObject* p = 0) : QIODevice(p) {}
    void connectToHost(const QHostAddress&, quint16, int = ReadWrite) {}
    void connectToHost(const QString&, quint16, int = ReadWrite) {}
    void disconnectFromHost() {}
    void abort() {}
    bool waitForConnected(int = 30000) { return true; }
    bool waitForDisconnected(int = 30000) { return true; }
    bool waitForReadyRead(int = 30000) { return true; }
    bool flush() { return true; }
    SocketState state() const { return UnconnectedState; }
    QString errorString() const { return QString(); }
    void setSocketOption(SocketOption, int) {}
    QHostAddress peerAddress() const { return QHostAddress(); }
    QHostAddress localAddress() const { return QHostAddress(); }
    quint16 localPort() const { return 0; }
    quint16 peerPort() const { return 0; }
    bool bind(const QHostAddress&, quint16 = 0, int = 0) { return true; }
    bool bind(quint16 = 0, int = 0) { return true; }
    bool isValid() const { return true; }
};
class QTcpSocket : public QAbstractSocket { public: QTcpSocket(QObject* p = 0) : QAbstractSocket(p) {} };
class QUdpSocket : public QAbstractSocket {
public:
    enum BindFlag { DefaultForPlatform = 0, ShareAddress = 1, DontShareAddress = 2, ReuseAddressHint = 4 };
    QUdpSocket(QObject* p = 0) : QAbstractSocket(p) {}
    using QAbstractSocket::bind;
    bool hasPendingDatagrams() const { return false; }
    qint64 pendingDatagramSize() const { return 0; }
    qint64 readDatagram(char*, qint64, QHostAddress* = 0, quint16* = 0) { return 0; }
    qint64 writeDatagram(const char*, qint64, const QHostAddress&, quint16) { return 0; }
    void setReadBufferSize(qint64) {}
};
template <typename T> class QList : public std::vector<T> {
public:
    void append(const T& v) { this->push_back(v); }
    int count() const { return (int)this->size(); }
    bool isEmpty() const { return this->empty(); }
    const T& at(int i) const { return (*this)[i]; }
};
class QAudioFormat {
public:
    enum Endian { BigEndian, LittleEndian };
    enum SampleType { Unknown, SignedInt, UnSignedInt, Float };
    void setCodec(const QString&) {}
    void setFrequency(int) {}
    void setSampleRate(int) {}
    void setSampleSize(int) {}
    void setSampleType(SampleType) {}
    void setByteOrder(Endian) {}
    void setChannels(int) {}
    void setChannelCount(int) {}
    int channels() const { return 2; }
    int channelCount() const { return 2; }
    int frequency() const { return 48000; }
    int sampleRate() const { return 48000; }
    int sampleSize() const { return 16; }
};
namespace QAudio {
enum Mode { AudioInput, AudioOutput };
enum Error { NoError, OpenError, IOError, UnderrunError, FatalError };
enum State { ActiveState, SuspendedState, StoppedState, IdleState };
}
class QAudioDeviceInfo {
public:
    static QList<QAudioDeviceInfo> availableDevices(QAudio::Mode) { return QList<QAudioDeviceInfo>(); }
    static QAudioDeviceInfo defaultOutputDevice() { return QAudioDeviceInfo(); }
    bool isFormatSupported(const QAudioFormat&) const { return true; }
    QAudioFormat nearestFormat(const QAudioFormat& f) const { return f; }
    QString deviceName() const { return QString(); }
};
class QAudioOutput : public QObject {
public:
    QAudioOutput(const QAudioDeviceInfo&, const QAudioFormat&, QObject* = 0) {}
    QAudioOutput(const QAudioFormat&, QObject* = 0) {}
    QIODevice* start() { return 0; }
    void start(QIODevice*) {}
    void stop() {}
    void reset() {}
    void suspend() {}
    void resume() {}
    int bytesFree() const { return 0; }
    int periodSize() const { return 0; }
    int bufferSize() const { return 0; }
    void setBufferSize(int) {}
    QAudio::Error error() const { return QAudio::NoError; }
    QAudio::State state() const { return QAudio::ActiveState; }
};
class QDir { public: static bool setCurrent(const char*) { return false; } };
class QFile : public QIODevice {
public:
    void setFileName(const char*) {}
    using QIODevice::write;
};
#endif
