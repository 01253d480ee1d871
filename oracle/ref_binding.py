"""ctypes binding for oracle/_ref/libcutesdr_ref*.so -- the UNMODIFIED reference
`dsp/*.cpp` compiled headless by oracle/Makefile (see oracle/ref_harness.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, the golden-fixture generator,
`__graft_entry__.smoke()` and bench.py's CPU baseline leg. The product
(`cutesdr_b200`) never imports this module.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)

# tDemodInfo field order used by the harness (dsp/demodulator.h:35-54 minus QString)
INFO_FIELDS = ("HiCut", "HiCutmin", "HiCutmax", "LowCut", "LowCutmin", "LowCutmax", "Offset",
               "SquelchValue", "AgcSlope", "AgcThresh", "AgcManualGain", "AgcDecay", "AgcOn", "AgcHangOn")


def ref_lib_path(big=False):
    return os.path.join(_HERE, "_ref", "libcutesdr_ref_big.so" if big else "libcutesdr_ref.so")


def ref_available(big=False):
    return os.path.exists(ref_lib_path(big))


_libs = {}


def load(big=False):
    if big in _libs:
        return _libs[big]
    L = C.CDLL(ref_lib_path(big))
    vp = C.c_void_p

    def sig(name, res, *args):
        f = getattr(L, name)
        f.restype = res
        f.argtypes = list(args)

    sig("ref_max_decstages", C.c_int)
    sig("ref_max_inbufsize", C.c_int)
    sig("ref_tap_enable", None, C.c_uint)
    sig("ref_tap_clear", None)
    sig("ref_tap_size", C.c_long, C.c_int)
    sig("ref_tap_read", None, C.c_int, _dp)
    sig("ref_downconvert_new", vp)
    sig("ref_downconvert_delete", None, vp)
    sig("ref_downconvert_set_frequency", None, vp, C.c_double)
    sig("ref_downconvert_set_cw_offset", None, vp, C.c_double)
    sig("ref_downconvert_set_data_rate", C.c_double, vp, C.c_double, C.c_double)
    sig("ref_downconvert_process", C.c_int, vp, C.c_int, _dp, _dp)
    sig("ref_downconvert_stages", C.c_int, vp, _ip, C.c_int)
    sig("ref_fastfir_new", vp)
    sig("ref_fastfir_delete", None, vp)
    sig("ref_fastfir_setup", None, vp, C.c_double, C.c_double, C.c_double, C.c_double)
    sig("ref_fastfir_process", C.c_int, vp, C.c_int, _dp, _dp)
    sig("ref_fastfir_coef", None, vp, _dp)
    sig("ref_fft_new", vp)
    sig("ref_fft_delete", None, vp)
    sig("ref_fft_set_params", None, vp, C.c_int, C.c_int, C.c_double, C.c_double)
    sig("ref_fft_set_ave", None, vp, C.c_int)
    sig("ref_fft_reset", None, vp)
    sig("ref_fft_put", C.c_int, vp, C.c_int, _dp)
    sig("ref_fft_get_screen", C.c_int, vp, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, _ip)
    sig("ref_fft_fwd", None, vp, _dp)
    sig("ref_fft_rev", None, vp, _dp)
    sig("ref_fft_size", C.c_int, vp)
    sig("ref_fft_avebuf", None, vp, _dp)
    sig("ref_fft_consts", None, vp, _dp)
    sig("ref_fft_bins", None, vp, _ip)
    sig("ref_agc_new", vp)
    sig("ref_agc_delete", None, vp)
    sig("ref_agc_set", None, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double)
    sig("ref_agc_process", None, vp, C.c_int, _dp, _dp)
    sig("ref_agc_sizes", None, vp, _ip)
    sig("ref_smeter_new", vp)
    sig("ref_smeter_delete", None, vp)
    sig("ref_smeter_process", None, vp, C.c_int, _dp, C.c_double)
    sig("ref_smeter_peak", C.c_double, vp)
    sig("ref_smeter_ave", C.c_double, vp)
    sig("ref_am_new", vp, C.c_double)
    sig("ref_am_delete", None, vp)
    sig("ref_am_set_bandwidth", None, vp, C.c_double)
    sig("ref_am_process", C.c_int, vp, C.c_int, _dp, _dp)
    sig("ref_am_process_stereo", C.c_int, vp, C.c_int, _dp, _dp)
    sig("ref_sam_new", vp, C.c_double)
    sig("ref_sam_delete", None, vp)
    sig("ref_sam_process", C.c_int, vp, C.c_int, _dp, _dp)
    sig("ref_sam_process_stereo", C.c_int, vp, C.c_int, _dp, _dp)
    sig("ref_fm_new", vp, C.c_double)
    sig("ref_fm_delete", None, vp)
    sig("ref_fm_set_squelch", None, vp, C.c_int)
    sig("ref_fm_process", C.c_int, vp, C.c_int, C.c_double, _dp, _dp)
    sig("ref_fm_process_stereo", C.c_int, vp, C.c_int, C.c_double, _dp, _dp)
    sig("ref_ssb_process", C.c_int, C.c_int, _dp, _dp)
    sig("ref_fir_new", vp)
    sig("ref_fir_delete", None, vp)
    sig("ref_fir_init_lp", C.c_int, vp, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double)
    sig("ref_fir_init_hp", C.c_int, vp, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double)
    sig("ref_fir_gen_hb", None, vp, C.c_double)
    sig("ref_fir_taps", C.c_int, vp, _dp, _dp, _dp)
    sig("ref_fir_process_real", None, vp, C.c_int, _dp, _dp)
    sig("ref_fir_process_cpx", None, vp, C.c_int, _dp, _dp)
    sig("ref_iir_new", vp)
    sig("ref_iir_delete", None, vp)
    sig("ref_iir_init_lp", None, vp, C.c_double, C.c_double, C.c_double)
    sig("ref_iir_process_real", None, vp, C.c_int, _dp, _dp)
    sig("ref_resampler_new", vp, C.c_int)
    sig("ref_resampler_delete", None, vp)
    sig("ref_resampler_real", C.c_int, vp, C.c_int, C.c_double, _dp, _dp)
    sig("ref_resampler_cpx", C.c_int, vp, C.c_int, C.c_double, _dp, _dp)
    sig("ref_resampler_mono16", C.c_int, vp, C.c_int, C.c_double, _dp, C.POINTER(C.c_short), C.c_double)
    sig("ref_resampler_stereo16", C.c_int, vp, C.c_int, C.c_double, _dp, C.POINTER(C.c_short), C.c_double)
    sig("ref_noiseproc_new", vp)
    sig("ref_noiseproc_delete", None, vp)
    sig("ref_noiseproc_setup", None, vp, C.c_int, C.c_double, C.c_double, C.c_double)
    sig("ref_noiseproc_process", None, vp, C.c_long, _dp)
    sig("ref_demod_new", vp)
    sig("ref_demod_delete", None, vp)
    sig("ref_demod_set_input_rate", None, vp, C.c_double)
    sig("ref_demod_set_demod", None, vp, C.c_int, _ip)
    sig("ref_demod_set_freq", None, vp, C.c_double)
    sig("ref_demod_output_rate", C.c_double, vp)
    sig("ref_demod_smeter_peak", C.c_double, vp)
    sig("ref_demod_smeter_ave", C.c_double, vp)
    sig("ref_demod_inbuf_limit", C.c_int, vp)
    sig("ref_demod_set_inbuf_limit", None, vp, C.c_int)
    sig("ref_chains_set_inbuf_limit", None, vp, C.c_int)
    sig("ref_demod_run", C.c_long, vp, C.c_long, C.c_void_p, C.c_int, C.c_int, _dp, C.c_long, C.c_int)
    sig("ref_bench_chains", C.c_double, C.c_int, _ip, _dp, _ip, C.c_double, C.c_long, C.POINTER(C.c_float),
        C.c_int, C.c_int, _dp)
    fp = C.POINTER(C.c_float)
    sig("ref_chains_new", vp, C.c_int, _ip, _dp, _ip, C.c_double, C.c_double, C.c_int)
    sig("ref_chains_delete", None, vp)
    sig("ref_chains_set_freq", None, vp, C.c_int, C.c_double)
    sig("ref_chains_set_demod", None, vp, C.c_int, C.c_int, _ip)
    sig("ref_chains_smeter_ave", C.c_double, vp, C.c_int)
    sig("ref_chains_run", C.c_double, vp, C.c_long, fp, C.c_int)
    sig("ref_chains_out_size", C.c_long, vp, C.c_int)
    sig("ref_chains_out_read", None, vp, C.c_int, _dp, C.c_int)
    sig("ref_chains_checksum", C.c_double, vp)
    sig("ref_noiseproc_process_f32", C.c_double, vp, C.c_long, fp)
    _libs[big] = L
    return L


def _d(a):
    return a.ctypes.data_as(_dp)


def _cpx_in(x):
    """complex array -> contiguous interleaved float64 copy (reference mutates inputs)."""
    x = np.asarray(x)
    out = np.empty(2 * x.size, dtype=np.float64)
    out[0::2] = x.real
    out[1::2] = x.imag
    return out


def _cpx_out(buf, n):
    return buf[0:2 * n:2] + 1j * buf[1:2 * n:2]


def info_array(info):
    """dict (INFO_FIELDS keys) -> int32[14]"""
    return np.array([int(info[k]) for k in INFO_FIELDS], dtype=np.int32)


class _Obj:
    _new = _delete = None

    def __init__(self, *args, big=False):
        self.L = load(big)
        self.h = getattr(self.L, self._new)(*args)

    def __del__(self):
        try:
            if getattr(self, "h", None):
                getattr(self.L, self._delete)(self.h)
                self.h = None
        except Exception:
            pass


class RefDownConvert(_Obj):
    _new, _delete = "ref_downconvert_new", "ref_downconvert_delete"

    def SetFrequency(self, f):
        self.L.ref_downconvert_set_frequency(self.h, float(f))

    def SetCwOffset(self, f):
        self.L.ref_downconvert_set_cw_offset(self.h, float(f))

    def SetDataRate(self, rate, bw):
        return self.L.ref_downconvert_set_data_rate(self.h, float(rate), float(bw))

    def stages(self):
        a = np.zeros(32, dtype=np.int32)
        n = self.L.ref_downconvert_stages(self.h, a.ctypes.data_as(_ip), 32)
        return [int(v) for v in a[:n]]

    def ProcessData(self, x):
        buf = _cpx_in(x)
        out = np.empty_like(buf)
        n = self.L.ref_downconvert_process(self.h, len(x), _d(buf), _d(out))
        return _cpx_out(out, n)


class RefFastFIR(_Obj):
    _new, _delete = "ref_fastfir_new", "ref_fastfir_delete"

    def SetupParameters(self, lo, hi, off, rate):
        self.L.ref_fastfir_setup(self.h, float(lo), float(hi), float(off), float(rate))

    def ProcessData(self, x):
        buf = _cpx_in(x)
        out = np.empty(2 * (len(x) + 2048), dtype=np.float64)
        n = self.L.ref_fastfir_process(self.h, len(x), _d(buf), _d(out))
        return _cpx_out(out, n)

    def coef(self):
        out = np.empty(4096, dtype=np.float64)
        self.L.ref_fastfir_coef(self.h, _d(out))
        return _cpx_out(out, 2048)


class RefFft(_Obj):
    _new, _delete = "ref_fft_new", "ref_fft_delete"

    def SetFFTParams(self, size, invert, dbcomp, fs):
        self.L.ref_fft_set_params(self.h, int(size), int(bool(invert)), float(dbcomp), float(fs))

    def SetFFTAve(self, ave):
        self.L.ref_fft_set_ave(self.h, int(ave))

    def ResetFFT(self):
        self.L.ref_fft_reset(self.h)

    def PutInDisplayFFT(self, x):
        buf = _cpx_in(x)
        return self.L.ref_fft_put(self.h, len(x), _d(buf))

    def GetScreenIntegerFFTData(self, maxh, maxw, maxdb, mindb, start, stop):
        out = np.zeros(max(maxw, 1), dtype=np.int32)
        ov = self.L.ref_fft_get_screen(self.h, maxh, maxw, float(maxdb), float(mindb), int(start), int(stop),
                                       out.ctypes.data_as(_ip))
        return bool(ov), out

    def FwdFFT(self, x):
        buf = _cpx_in(x)
        self.L.ref_fft_fwd(self.h, _d(buf))
        return _cpx_out(buf, len(x))

    def RevFFT(self, x):
        buf = _cpx_in(x)
        self.L.ref_fft_rev(self.h, _d(buf))
        return _cpx_out(buf, len(x))

    def size(self):
        return self.L.ref_fft_size(self.h)

    def avebuf(self):
        out = np.empty(self.size(), dtype=np.float64)
        self.L.ref_fft_avebuf(self.h, _d(out))
        return out

    def consts(self):
        out = np.empty(2, dtype=np.float64)
        self.L.ref_fft_consts(self.h, _d(out))
        return float(out[0]), float(out[1])

    def bins(self):
        out = np.zeros(2, dtype=np.int32)
        self.L.ref_fft_bins(self.h, out.ctypes.data_as(_ip))
        return int(out[0]), int(out[1])


class RefAgc(_Obj):
    _new, _delete = "ref_agc_new", "ref_agc_delete"

    def SetParameters(self, on, hang, thresh, mgain, slope, decay, rate):
        self.L.ref_agc_set(self.h, int(on), int(hang), int(thresh), int(mgain), int(slope), int(decay), float(rate))

    def ProcessData(self, x):
        buf = _cpx_in(x)
        out = np.empty_like(buf)
        self.L.ref_agc_process(self.h, len(x), _d(buf), _d(out))
        return _cpx_out(out, len(x))

    def sizes(self):
        a = np.zeros(2, dtype=np.int32)
        self.L.ref_agc_sizes(self.h, a.ctypes.data_as(_ip))
        return int(a[0]), int(a[1])


class RefSMeter(_Obj):
    _new, _delete = "ref_smeter_new", "ref_smeter_delete"

    def ProcessData(self, x, rate):
        buf = _cpx_in(x)
        self.L.ref_smeter_process(self.h, len(x), _d(buf), float(rate))

    def GetPeak(self):
        return self.L.ref_smeter_peak(self.h)

    def GetAve(self):
        return self.L.ref_smeter_ave(self.h)


class RefAmDemod(_Obj):
    _new, _delete = "ref_am_new", "ref_am_delete"

    def SetBandwidth(self, bw):
        self.L.ref_am_set_bandwidth(self.h, float(bw))

    def ProcessData(self, x, stereo=False):
        buf = _cpx_in(x)
        if stereo:
            out = np.empty(2 * len(x), dtype=np.float64)
            n = self.L.ref_am_process_stereo(self.h, len(x), _d(buf), _d(out))
            return _cpx_out(out, n)
        out = np.empty(len(x), dtype=np.float64)
        n = self.L.ref_am_process(self.h, len(x), _d(buf), _d(out))
        return out[:n]


class RefSamDemod(_Obj):
    _new, _delete = "ref_sam_new", "ref_sam_delete"

    def ProcessData(self, x, stereo=False):
        buf = _cpx_in(x)
        if stereo:
            out = np.empty(2 * len(x), dtype=np.float64)
            n = self.L.ref_sam_process_stereo(self.h, len(x), _d(buf), _d(out))
            return _cpx_out(out, n)
        out = np.empty(len(x), dtype=np.float64)
        n = self.L.ref_sam_process(self.h, len(x), _d(buf), _d(out))
        return out[:n]


class RefFmDemod(_Obj):
    _new, _delete = "ref_fm_new", "ref_fm_delete"

    def SetSquelch(self, v):
        self.L.ref_fm_set_squelch(self.h, int(v))

    def ProcessData(self, x, fmbw, stereo=False):
        buf = _cpx_in(x)
        if stereo:
            out = np.empty(2 * len(x), dtype=np.float64)
            n = self.L.ref_fm_process_stereo(self.h, len(x), float(fmbw), _d(buf), _d(out))
            return _cpx_out(out, n)
        out = np.empty(len(x), dtype=np.float64)
        n = self.L.ref_fm_process(self.h, len(x), float(fmbw), _d(buf), _d(out))
        return out[:n]


def ref_ssb(x, big=False):
    L = load(big)
    buf = _cpx_in(x)
    out = np.empty(len(x), dtype=np.float64)
    n = L.ref_ssb_process(len(x), _d(buf), _d(out))
    return out[:n]


class RefFir(_Obj):
    _new, _delete = "ref_fir_new", "ref_fir_delete"

    def InitLPFilter(self, scale, astop, fpass, fstop, fs):
        return self.L.ref_fir_init_lp(self.h, scale, astop, fpass, fstop, fs)

    def InitHPFilter(self, scale, astop, fpass, fstop, fs):
        return self.L.ref_fir_init_hp(self.h, scale, astop, fpass, fstop, fs)

    def GenerateHBFilter(self, off):
        self.L.ref_fir_gen_hb(self.h, float(off))

    def taps(self):
        c, i, q = (np.zeros(75) for _ in range(3))
        n = self.L.ref_fir_taps(self.h, _d(c), _d(i), _d(q))
        return c[:n], i[:n], q[:n]

    def ProcessFilter(self, x):
        x = np.asarray(x)
        if np.iscomplexobj(x):
            buf = _cpx_in(x)
            out = np.empty_like(buf)
            self.L.ref_fir_process_cpx(self.h, len(x), _d(buf), _d(out))
            return _cpx_out(out, len(x))
        buf = np.ascontiguousarray(x, dtype=np.float64).copy()
        out = np.empty_like(buf)
        self.L.ref_fir_process_real(self.h, len(x), _d(buf), _d(out))
        return out


class RefIir(_Obj):
    _new, _delete = "ref_iir_new", "ref_iir_delete"

    def InitLP(self, f0, q, fs):
        self.L.ref_iir_init_lp(self.h, f0, q, fs)

    def ProcessFilter(self, x):
        buf = np.ascontiguousarray(x, dtype=np.float64).copy()
        out = np.empty_like(buf)
        self.L.ref_iir_process_real(self.h, len(x), _d(buf), _d(out))
        return out


class RefFractResampler(_Obj):
    _new, _delete = "ref_resampler_new", "ref_resampler_delete"

    def __init__(self, maxin=8192, big=False):
        super().__init__(int(maxin), big=big)

    def Resample(self, x, rate, gain=None):
        x = np.asarray(x)
        cap = int(len(x) / rate) + 64
        if np.iscomplexobj(x):
            buf = _cpx_in(x)
            if gain is None:
                out = np.empty(2 * cap, dtype=np.float64)
                n = self.L.ref_resampler_cpx(self.h, len(x), float(rate), _d(buf), _d(out))
                return _cpx_out(out, n)
            out = np.empty(2 * cap, dtype=np.int16)
            n = self.L.ref_resampler_stereo16(self.h, len(x), float(rate), _d(buf),
                                              out.ctypes.data_as(C.POINTER(C.c_short)), float(gain))
            return out[:2 * n].reshape(n, 2)
        buf = np.ascontiguousarray(x, dtype=np.float64).copy()
        if gain is None:
            out = np.empty(cap, dtype=np.float64)
            n = self.L.ref_resampler_real(self.h, len(x), float(rate), _d(buf), _d(out))
            return out[:n]
        out = np.empty(cap, dtype=np.int16)
        n = self.L.ref_resampler_mono16(self.h, len(x), float(rate), _d(buf),
                                        out.ctypes.data_as(C.POINTER(C.c_short)), float(gain))
        return out[:n]


class RefNoiseProc(_Obj):
    _new, _delete = "ref_noiseproc_new", "ref_noiseproc_delete"

    def SetupBlanker(self, on, thr, width, fs):
        self.L.ref_noiseproc_setup(self.h, int(on), float(thr), float(width), float(fs))

    def ProcessBlanker(self, x):
        buf = _cpx_in(x)
        self.L.ref_noiseproc_process(self.h, len(x), _d(buf))
        return _cpx_out(buf, len(x))


class RefDemodulator(_Obj):
    _new, _delete = "ref_demod_new", "ref_demod_delete"

    def SetInputSampleRate(self, r):
        self.L.ref_demod_set_input_rate(self.h, float(r))

    def SetDemod(self, mode, info):
        a = info_array(info)
        self.L.ref_demod_set_demod(self.h, int(mode), a.ctypes.data_as(_ip))

    def SetDemodFreq(self, f):
        self.L.ref_demod_set_freq(self.h, float(f))

    def GetOutputRate(self):
        return self.L.ref_demod_output_rate(self.h)

    def GetSMeterPeak(self):
        return self.L.ref_demod_smeter_peak(self.h)

    def GetSMeterAve(self):
        return self.L.ref_demod_smeter_ave(self.h)

    def inbuf_limit(self):
        return self.L.ref_demod_inbuf_limit(self.h)

    def set_inbuf_limit(self, limit):
        """test aid: override m_InBufLimit until the next SetDemod"""
        self.L.ref_demod_set_inbuf_limit(self.h, int(limit))

    def run(self, iq, packet=256, stereo=False, taps=()):
        """Feed complex64 (or complex128) samples in `packet`-sized calls; returns audio and
        optionally the PROFILE_n tap streams (dict profile -> array)."""
        iq = np.asarray(iq)
        if iq.dtype == np.complex64:
            raw, is_double = np.ascontiguousarray(iq), 0
        else:
            raw, is_double = np.ascontiguousarray(iq.astype(np.complex128)), 1
        mask = 0
        for p in taps:
            mask |= 1 << p
        self.L.ref_tap_clear()
        self.L.ref_tap_enable(mask)
        cap = 2 * (len(iq) // 4 + 4096)
        out = np.empty(cap, dtype=np.float64)
        n = self.L.ref_demod_run(self.h, len(iq), raw.ctypes.data, is_double, packet, _d(out), cap, int(stereo))
        assert n <= cap
        audio = _cpx_out(out, n // 2) if stereo else out[:n].copy()
        tapd = {}
        for p in taps:
            sz = self.L.ref_tap_size(p)
            b = np.empty(sz, dtype=np.float64)
            if sz:
                self.L.ref_tap_read(p, _d(b))
            tapd[p] = b
        self.L.ref_tap_enable(0)
        self.L.ref_tap_clear()
        return (audio, tapd) if taps else audio


def bench_chains(modes, freqs, infos_by_mode, in_rate, iq_c64, nthreads, resample48k=False, big=False):
    """Multi-threaded CPU baseline: one CDemodulator per channel. Returns (seconds, checksum)."""
    L = load(big)
    modes = np.ascontiguousarray(modes, dtype=np.int32)
    freqs = np.ascontiguousarray(freqs, dtype=np.float64)
    infos = np.zeros(7 * 14, dtype=np.int32)
    for m, info in infos_by_mode.items():
        infos[14 * m:14 * m + 14] = info_array(info)
    iq = np.ascontiguousarray(iq_c64, dtype=np.complex64)
    cs = C.c_double(0.0)
    t = L.ref_bench_chains(len(modes), modes.ctypes.data_as(_ip), _d(freqs), infos.ctypes.data_as(_ip),
                           float(in_rate), len(iq), iq.ctypes.data_as(C.POINTER(C.c_float)), int(nthreads),
                           int(resample48k), C.byref(cs))
    return t, cs.value


class RefChainSet:
    """N persistent CDemodulator (+ CFractResampler) chains on one wideband stream, run over std::threads. Objects are
    built here (outside any timed region) and keep their state across run() calls, like a receiver that is running."""

    def __init__(self, modes, freqs, infos, in_rate, audio_rate=0.0, keep_output=True, big=False):
        self.L = load(big)
        self.n = len(modes)
        m = np.ascontiguousarray(modes, dtype=np.int32)
        f = np.ascontiguousarray(freqs, dtype=np.float64)
        inf = np.concatenate([info_array(i) for i in infos]).astype(np.int32)
        self.h = self.L.ref_chains_new(self.n, m.ctypes.data_as(_ip), _d(f), inf.ctypes.data_as(_ip), float(in_rate),
                                       float(audio_rate), int(bool(keep_output)))

    def __del__(self):
        try:
            if self.h:
                self.L.ref_chains_delete(self.h)
                self.h = None
        except Exception:
            pass

    def run(self, iq_c64, nthreads):
        """Feeds the complex64 samples to every chain; returns the seconds the threaded section took."""
        iq = np.ascontiguousarray(iq_c64, dtype=np.complex64)
        return self.L.ref_chains_run(self.h, len(iq), iq.ctypes.data_as(C.POINTER(C.c_float)), int(nthreads))

    def output(self, c, clear=True):
        n = self.L.ref_chains_out_size(self.h, int(c))
        out = np.empty(n, dtype=np.float64)
        self.L.ref_chains_out_read(self.h, int(c), _d(out), int(clear))
        return out

    def SetDemodFreq(self, c, f):
        self.L.ref_chains_set_freq(self.h, int(c), float(f))

    def SetDemod(self, c, mode, info):
        a = info_array(info)
        self.L.ref_chains_set_demod(self.h, int(c), int(mode), a.ctypes.data_as(_ip))

    def GetSMeterAve(self, c):
        return self.L.ref_chains_smeter_ave(self.h, int(c))

    def set_inbuf_limit(self, limit):
        """test aid: every chain's m_InBufLimit (until a SetDemod recomputes it)"""
        self.L.ref_chains_set_inbuf_limit(self.h, int(limit))

    def checksum(self):
        return self.L.ref_chains_checksum(self.h)


def blank_stream_f32(nb, iq_c64):
    """CNoiseProc::ProcessBlanker over a complex64 stream, in place (<= 4096 samples per call). Returns seconds."""
    assert iq_c64.dtype == np.complex64 and iq_c64.flags.c_contiguous
    return nb.L.ref_noiseproc_process_f32(nb.h, len(iq_c64), iq_c64.ctypes.data_as(C.POINTER(C.c_float)))
