"""ctypes binding for oracle/liboracle.so -- the double-precision CPU restatement of
the reference receive chain (oracle/cutesdr_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, `__graft_entry__.smoke()` and
bench.py's cpu_baseline leg as the checker. `cutesdr_b200` never imports it.

Class and method names mirror oracle/ref_binding.py so a test can run the same
scenario against either the compiled reference or the restatement.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)

INFO_FIELDS = ("HiCut", "HiCutmin", "HiCutmax", "LowCut", "LowCutmin", "LowCutmax", "Offset",
               "SquelchValue", "AgcSlope", "AgcThresh", "AgcManualGain", "AgcDecay", "AgcOn", "AgcHangOn")


class DemodInfo(C.Structure):
    _fields_ = [(k, C.c_int) for k in INFO_FIELDS]


def make_info(d):
    s = DemodInfo()
    for k in INFO_FIELDS:
        setattr(s, k, int(d[k]))
    return s


def lib_path():
    return os.path.join(_HERE, "liboracle.so")


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "oracle"])


_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(lib_path()):
        build()
    L = C.CDLL(lib_path())
    vp = C.c_void_p

    def sig(name, res, *args):
        f = getattr(L, name)
        f.restype = res
        f.argtypes = list(args)

    sig("orc_plan_stages", C.c_int, C.c_double, C.c_double, _ip, C.c_int, _dp)
    sig("orc_downconvert_create", vp)
    sig("orc_downconvert_destroy", None, vp)
    sig("orc_downconvert_set_frequency", None, vp, C.c_double)
    sig("orc_downconvert_set_cw_offset", None, vp, C.c_double)
    sig("orc_downconvert_set_data_rate", C.c_double, vp, C.c_double, C.c_double)
    sig("orc_downconvert_stages", C.c_int, vp, _ip, C.c_int)
    sig("orc_downconvert_process", C.c_int, vp, C.c_int, _dp, _dp)
    sig("orc_fastfir_create", vp)
    sig("orc_fastfir_destroy", None, vp)
    sig("orc_fastfir_setup", None, vp, C.c_double, C.c_double, C.c_double, C.c_double)
    sig("orc_fastfir_process", C.c_int, vp, C.c_int, _dp, _dp)
    sig("orc_fastfir_taps", None, vp, _dp)
    sig("orc_fft_create", vp)
    sig("orc_fft_destroy", None, vp)
    sig("orc_fft_set_params", None, vp, C.c_int, C.c_int, C.c_double, C.c_double)
    sig("orc_fft_set_ave", None, vp, C.c_int)
    sig("orc_fft_reset", None, vp)
    sig("orc_fft_put", C.c_int, vp, C.c_int, _dp)
    sig("orc_fft_get_screen", C.c_int, vp, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, _ip)
    sig("orc_fft_size", C.c_int, vp)
    sig("orc_fft_avebuf", None, vp, _dp)
    sig("orc_smeter_create", vp)
    sig("orc_smeter_destroy", None, vp)
    sig("orc_smeter_process", None, vp, C.c_int, _dp, C.c_double)
    sig("orc_smeter_peak", C.c_double, vp)
    sig("orc_smeter_ave", C.c_double, vp)
    sig("orc_agc_create", vp)
    sig("orc_agc_destroy", None, vp)
    sig("orc_agc_set", None, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double)
    sig("orc_agc_process", None, vp, C.c_int, _dp, _dp)
    sig("orc_fir_create", vp)
    sig("orc_fir_destroy", None, vp)
    sig("orc_fir_init_lp", C.c_int, vp, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double)
    sig("orc_fir_init_hp", C.c_int, vp, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double)
    sig("orc_fir_make_hilbert_pair", None, vp, C.c_double)
    sig("orc_fir_taps", C.c_int, vp, _dp, _dp, _dp)
    sig("orc_fir_process_real", None, vp, C.c_int, _dp, _dp)
    sig("orc_fir_process_cpx", None, vp, C.c_int, _dp, _dp)
    sig("orc_biquad_init_lp", None, _dp, C.c_double, C.c_double, C.c_double)
    sig("orc_biquad_process", None, _dp, C.c_int, _dp, _dp)
    sig("orc_am_create", vp, C.c_double)
    sig("orc_am_destroy", None, vp)
    sig("orc_am_set_bandwidth", None, vp, C.c_double)
    sig("orc_am_process", C.c_int, vp, C.c_int, _dp, _dp)
    sig("orc_sam_create", vp, C.c_double)
    sig("orc_sam_destroy", None, vp)
    sig("orc_sam_process", C.c_int, vp, C.c_int, _dp, _dp)
    sig("orc_sam_process_stereo", C.c_int, vp, C.c_int, _dp, _dp)
    sig("orc_fm_create", vp, C.c_double)
    sig("orc_fm_destroy", None, vp)
    sig("orc_fm_set_squelch", None, vp, C.c_int)
    sig("orc_fm_process", C.c_int, vp, C.c_int, C.c_double, _dp, _dp)
    sig("orc_resampler_create", vp, C.c_int)
    sig("orc_resampler_destroy", None, vp)
    sig("orc_resampler_real", C.c_int, vp, C.c_int, C.c_double, _dp, _dp)
    sig("orc_resampler_cpx", C.c_int, vp, C.c_int, C.c_double, _dp, _dp)
    sig("orc_resampler_mono16", C.c_int, vp, C.c_int, C.c_double, _dp, C.POINTER(C.c_short), C.c_double)
    sig("orc_resampler_stereo16", C.c_int, vp, C.c_int, C.c_double, _dp, C.POINTER(C.c_short), C.c_double)
    sig("orc_resampler_table", _dp, vp, _ip)
    sig("orc_blanker_create", vp)
    sig("orc_blanker_destroy", None, vp)
    sig("orc_blanker_setup", None, vp, C.c_int, C.c_double, C.c_double, C.c_double)
    sig("orc_blanker_process", None, vp, C.c_long, _dp)
    sig("orc_demod_create", vp)
    sig("orc_demod_destroy", None, vp)
    sig("orc_demod_set_input_rate", None, vp, C.c_double)
    sig("orc_demod_set_demod", None, vp, C.c_int, C.POINTER(DemodInfo))
    sig("orc_demod_set_freq", None, vp, C.c_double)
    sig("orc_demod_output_rate", C.c_double, vp)
    sig("orc_demod_inbuf_limit", C.c_int, vp)
    sig("orc_demod_set_inbuf_limit", None, vp, C.c_int)
    sig("orc_demod_smeter_peak", C.c_double, vp)
    sig("orc_demod_smeter_ave", C.c_double, vp)
    sig("orc_demod_set_tap", None, vp, C.c_int, _dp, C.c_long)
    sig("orc_demod_tap_count", C.c_long, vp, C.c_int)
    sig("orc_demod_process", C.c_int, vp, C.c_int, _dp, _dp)
    sig("orc_demod_run_c64", C.c_long, vp, C.c_long, C.POINTER(C.c_float), C.c_int, _dp, C.c_long)
    _lib = L
    return L


def _d(a):
    return a.ctypes.data_as(_dp)


def _cpx_in(x):
    x = np.asarray(x)
    out = np.empty(2 * x.size, dtype=np.float64)
    out[0::2] = x.real
    out[1::2] = x.imag
    return out


def _cpx_out(buf, n):
    return buf[0:2 * n:2] + 1j * buf[1:2 * n:2]


def plan_stages(in_rate, max_bw):
    L = load()
    a = np.zeros(32, dtype=np.int32)
    r = C.c_double(0)
    n = L.orc_plan_stages(float(in_rate), float(max_bw), a.ctypes.data_as(_ip), 32, C.byref(r))
    return [int(v) for v in a[:n]], r.value


class _Obj:
    _new = _delete = None

    def __init__(self, *args):
        self.L = load()
        self.h = getattr(self.L, self._new)(*args)

    def __del__(self):
        try:
            if getattr(self, "h", None):
                getattr(self.L, self._delete)(self.h)
                self.h = None
        except Exception:
            pass


class DownConvert(_Obj):
    _new, _delete = "orc_downconvert_create", "orc_downconvert_destroy"

    def SetFrequency(self, f):
        self.L.orc_downconvert_set_frequency(self.h, float(f))

    def SetCwOffset(self, f):
        self.L.orc_downconvert_set_cw_offset(self.h, float(f))

    def SetDataRate(self, rate, bw):
        return self.L.orc_downconvert_set_data_rate(self.h, float(rate), float(bw))

    def stages(self):
        a = np.zeros(32, dtype=np.int32)
        n = self.L.orc_downconvert_stages(self.h, a.ctypes.data_as(_ip), 32)
        return [int(v) for v in a[:n]]

    def ProcessData(self, x):
        buf = _cpx_in(x)
        out = np.empty_like(buf)
        n = self.L.orc_downconvert_process(self.h, len(x), _d(buf), _d(out))
        return _cpx_out(out, n)


class FastFIR(_Obj):
    _new, _delete = "orc_fastfir_create", "orc_fastfir_destroy"

    def SetupParameters(self, lo, hi, off, rate):
        self.L.orc_fastfir_setup(self.h, float(lo), float(hi), float(off), float(rate))

    def ProcessData(self, x):
        buf = _cpx_in(x)
        out = np.empty(2 * (len(x) + 2048), dtype=np.float64)
        n = self.L.orc_fastfir_process(self.h, len(x), _d(buf), _d(out))
        return _cpx_out(out, n)

    def taps(self):
        out = np.empty(2 * 1025, dtype=np.float64)
        self.L.orc_fastfir_taps(self.h, _d(out))
        return _cpx_out(out, 1025)


class Fft(_Obj):
    _new, _delete = "orc_fft_create", "orc_fft_destroy"

    def SetFFTParams(self, size, invert, dbcomp, fs):
        self.L.orc_fft_set_params(self.h, int(size), int(bool(invert)), float(dbcomp), float(fs))

    def SetFFTAve(self, ave):
        self.L.orc_fft_set_ave(self.h, int(ave))

    def ResetFFT(self):
        self.L.orc_fft_reset(self.h)

    def PutInDisplayFFT(self, x):
        buf = _cpx_in(x)
        return self.L.orc_fft_put(self.h, len(x), _d(buf))

    def GetScreenIntegerFFTData(self, maxh, maxw, maxdb, mindb, start, stop):
        out = np.zeros(max(maxw, 1), dtype=np.int32)
        ov = self.L.orc_fft_get_screen(self.h, maxh, maxw, float(maxdb), float(mindb), int(start), int(stop),
                                       out.ctypes.data_as(_ip))
        return bool(ov), out

    def size(self):
        return self.L.orc_fft_size(self.h)

    def avebuf(self):
        out = np.empty(self.size(), dtype=np.float64)
        self.L.orc_fft_avebuf(self.h, _d(out))
        return out


class SMeter(_Obj):
    _new, _delete = "orc_smeter_create", "orc_smeter_destroy"

    def ProcessData(self, x, rate):
        buf = _cpx_in(x)
        self.L.orc_smeter_process(self.h, len(x), _d(buf), float(rate))

    def GetPeak(self):
        return self.L.orc_smeter_peak(self.h)

    def GetAve(self):
        return self.L.orc_smeter_ave(self.h)


class Agc(_Obj):
    _new, _delete = "orc_agc_create", "orc_agc_destroy"

    def SetParameters(self, on, hang, thresh, mgain, slope, decay, rate):
        self.L.orc_agc_set(self.h, int(on), int(hang), int(thresh), int(mgain), int(slope), int(decay), float(rate))

    def ProcessData(self, x):
        buf = _cpx_in(x)
        out = np.empty_like(buf)
        self.L.orc_agc_process(self.h, len(x), _d(buf), _d(out))
        return _cpx_out(out, len(x))


class Fir(_Obj):
    _new, _delete = "orc_fir_create", "orc_fir_destroy"

    def InitLPFilter(self, scale, astop, fpass, fstop, fs):
        return self.L.orc_fir_init_lp(self.h, scale, astop, fpass, fstop, fs)

    def InitHPFilter(self, scale, astop, fpass, fstop, fs):
        return self.L.orc_fir_init_hp(self.h, scale, astop, fpass, fstop, fs)

    def GenerateHBFilter(self, off):
        self.L.orc_fir_make_hilbert_pair(self.h, float(off))

    def taps(self):
        c, i, q = (np.zeros(75) for _ in range(3))
        n = self.L.orc_fir_taps(self.h, _d(c), _d(i), _d(q))
        return c[:n], i[:n], q[:n]

    def ProcessFilter(self, x):
        x = np.asarray(x)
        if np.iscomplexobj(x):
            buf = _cpx_in(x)
            out = np.empty_like(buf)
            self.L.orc_fir_process_cpx(self.h, len(x), _d(buf), _d(out))
            return _cpx_out(out, len(x))
        buf = np.ascontiguousarray(x, dtype=np.float64).copy()
        out = np.empty_like(buf)
        self.L.orc_fir_process_real(self.h, len(x), _d(buf), _d(out))
        return out


class Iir:
    def __init__(self):
        self.L = load()
        self.s = np.zeros(7, dtype=np.float64)

    def InitLP(self, f0, q, fs):
        self.L.orc_biquad_init_lp(_d(self.s), f0, q, fs)

    def ProcessFilter(self, x):
        buf = np.ascontiguousarray(x, dtype=np.float64).copy()
        out = np.empty_like(buf)
        self.L.orc_biquad_process(_d(self.s), len(x), _d(buf), _d(out))
        return out


class AmDemod(_Obj):
    _new, _delete = "orc_am_create", "orc_am_destroy"

    def SetBandwidth(self, bw):
        self.L.orc_am_set_bandwidth(self.h, float(bw))

    def ProcessData(self, x):
        buf = _cpx_in(x)
        out = np.empty(len(x), dtype=np.float64)
        n = self.L.orc_am_process(self.h, len(x), _d(buf), _d(out))
        return out[:n]


class SamDemod(_Obj):
    _new, _delete = "orc_sam_create", "orc_sam_destroy"

    def ProcessData(self, x, stereo=False):
        buf = _cpx_in(x)
        if stereo:
            out = np.empty(2 * len(x), dtype=np.float64)
            n = self.L.orc_sam_process_stereo(self.h, len(x), _d(buf), _d(out))
            return _cpx_out(out, n)
        out = np.empty(len(x), dtype=np.float64)
        n = self.L.orc_sam_process(self.h, len(x), _d(buf), _d(out))
        return out[:n]


class FmDemod(_Obj):
    _new, _delete = "orc_fm_create", "orc_fm_destroy"

    def SetSquelch(self, v):
        self.L.orc_fm_set_squelch(self.h, int(v))

    def ProcessData(self, x, fmbw):
        buf = _cpx_in(x)
        out = np.empty(len(x), dtype=np.float64)
        n = self.L.orc_fm_process(self.h, len(x), float(fmbw), _d(buf), _d(out))
        return out[:n]


def ssb(x):
    return np.asarray(x).real.astype(np.float64).copy()   # dsp/ssbdemod.cpp:48-53


class FractResampler(_Obj):
    _new, _delete = "orc_resampler_create", "orc_resampler_destroy"

    def __init__(self, maxin=8192):
        super().__init__(int(maxin))

    def table(self):
        n = C.c_int(0)
        p = self.L.orc_resampler_table(self.h, C.byref(n))
        return np.ctypeslib.as_array(p, shape=(n.value,)).copy()

    def Resample(self, x, rate, gain=None):
        x = np.asarray(x)
        cap = int(len(x) / rate) + 64
        if np.iscomplexobj(x):
            buf = _cpx_in(x)
            if gain is None:
                out = np.empty(2 * cap, dtype=np.float64)
                n = self.L.orc_resampler_cpx(self.h, len(x), float(rate), _d(buf), _d(out))
                return _cpx_out(out, n)
            out = np.empty(2 * cap, dtype=np.int16)
            n = self.L.orc_resampler_stereo16(self.h, len(x), float(rate), _d(buf),
                                              out.ctypes.data_as(C.POINTER(C.c_short)), float(gain))
            return out[:2 * n].reshape(n, 2)
        buf = np.ascontiguousarray(x, dtype=np.float64).copy()
        if gain is None:
            out = np.empty(cap, dtype=np.float64)
            n = self.L.orc_resampler_real(self.h, len(x), float(rate), _d(buf), _d(out))
            return out[:n]
        out = np.empty(cap, dtype=np.int16)
        n = self.L.orc_resampler_mono16(self.h, len(x), float(rate), _d(buf),
                                        out.ctypes.data_as(C.POINTER(C.c_short)), float(gain))
        return out[:n]


class NoiseProc(_Obj):
    _new, _delete = "orc_blanker_create", "orc_blanker_destroy"

    def SetupBlanker(self, on, thr, width, fs):
        self.L.orc_blanker_setup(self.h, int(on), float(thr), float(width), float(fs))

    def ProcessBlanker(self, x):
        buf = _cpx_in(x)
        self.L.orc_blanker_process(self.h, len(x), _d(buf))
        return _cpx_out(buf, len(x))


class Demodulator(_Obj):
    _new, _delete = "orc_demod_create", "orc_demod_destroy"

    def SetInputSampleRate(self, r):
        self.L.orc_demod_set_input_rate(self.h, float(r))

    def SetDemod(self, mode, info):
        s = make_info(info)
        self.L.orc_demod_set_demod(self.h, int(mode), C.byref(s))

    def SetDemodFreq(self, f):
        self.L.orc_demod_set_freq(self.h, float(f))

    def GetOutputRate(self):
        return self.L.orc_demod_output_rate(self.h)

    def GetSMeterPeak(self):
        return self.L.orc_demod_smeter_peak(self.h)

    def GetSMeterAve(self):
        return self.L.orc_demod_smeter_ave(self.h)

    def inbuf_limit(self):
        return self.L.orc_demod_inbuf_limit(self.h)

    def set_inbuf_limit(self, limit):
        """test aid: override m_InBufLimit until the next SetDemod"""
        self.L.orc_demod_set_inbuf_limit(self.h, int(limit))

    def run(self, iq, packet=256, taps=()):
        iq = np.ascontiguousarray(iq, dtype=np.complex64)
        cap = len(iq) // 4 + 4096
        out = np.empty(cap, dtype=np.float64)
        bufs = {}
        for p in taps:
            bufs[p] = np.empty(2 * cap if p < 4 else cap, dtype=np.float64)
            self.L.orc_demod_set_tap(self.h, p, _d(bufs[p]), len(bufs[p]))
        n = self.L.orc_demod_run_c64(self.h, len(iq), iq.ctypes.data_as(C.POINTER(C.c_float)), packet, _d(out), cap)
        assert n <= cap
        tapd = {}
        for p in taps:
            cnt = self.L.orc_demod_tap_count(self.h, p)
            assert cnt <= len(bufs[p])
            tapd[p] = bufs[p][:cnt].copy()
            self.L.orc_demod_set_tap(self.h, p, None, 0)
        return (out[:n].copy(), tapd) if taps else out[:n].copy()
