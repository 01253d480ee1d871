"""Host-side mirror of the reference's operator API over the C ABI (include/cutesdr_cuda.h).

Class and method names follow the reference headers (dsp/downconvert.h, dsp/fastfir.h, dsp/fft.h,
dsp/demodulator.h, dsp/agc.h, dsp/fractresampler.h, dsp/noiseproc.h); `ReceiverBank` is the batched
form the north star adds (N x CDemodulator on one wideband stream). Arrays are numpy; complex
streams are complex128 for the TYPECPX entry points and complex64 for the bank's fast path.
"""
import ctypes as C

import numpy as np

from .lib import DemodInfo, check, load_library
from .modes import INFO_FIELDS


def _info_struct(info):
    s = DemodInfo()
    for k in INFO_FIELDS:
        setattr(s, k, int(info[k]))
    return s


def _cpx_to_f64(x):
    x = np.asarray(x)
    out = np.empty(2 * x.size, dtype=np.float64)
    out[0::2] = x.real
    out[1::2] = x.imag
    return out


def _f64_to_cpx(buf, n):
    return buf[0:2 * n:2] + 1j * buf[1:2 * n:2]


def _dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class _Handle:
    _create = _destroy = None

    def __init__(self, device=0):
        self.L = load_library()
        h = C.c_void_p()
        check(getattr(self.L, self._create)(C.byref(h), int(device)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            getattr(self.L, self._destroy)(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ReceiverBank:
    """N virtual receivers (N x CDemodulator, dsp/demodulator.h:56-100) on one wideband stream."""

    def __init__(self, n_channels, in_rate, device=0):
        self.L = load_library()
        h = C.c_void_p()
        check(self.L.cutesdr_bank_create(C.byref(h), int(n_channels), float(in_rate), int(device)))
        self.h = h
        self.n_channels = int(n_channels)
        self.in_rate = float(in_rate)

    def close(self):
        if getattr(self, "h", None):
            self.L.cutesdr_bank_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- CDemodulator setters, per channel
    def SetDemod(self, ch, mode, info):
        s = _info_struct(info)
        check(self.L.cutesdr_bank_set_demod(self.h, int(ch), int(mode), C.byref(s)))

    def SetDemodFreq(self, ch, freq):
        check(self.L.cutesdr_bank_set_demod_freq(self.h, int(ch), float(freq)))

    def GetOutputRate(self, ch):
        r = C.c_double()
        check(self.L.cutesdr_bank_get_output_rate(self.h, int(ch), C.byref(r)))
        return r.value

    def GetSMeter(self, ch):
        """(peak, ave); reading the peak resets it (CSMeter::GetPeak, dsp/smeter.cpp:99-104)"""
        p, a = C.c_double(), C.c_double()
        check(self.L.cutesdr_bank_get_smeter(self.h, int(ch), C.byref(p), C.byref(a)))
        return p.value, a.value

    def GetSMeterAve(self, ch):
        a = C.c_double()
        check(self.L.cutesdr_bank_get_smeter(self.h, int(ch), None, C.byref(a)))
        return a.value

    def SetupNoiseProc(self, on, threshold, width_us):
        check(self.L.cutesdr_bank_set_noiseproc(self.h, int(bool(on)), float(threshold), float(width_us)))

    def SetAudioRate(self, rate):
        check(self.L.cutesdr_bank_set_audio_rate(self.h, float(rate)))

    def SetStereo(self, on):
        check(self.L.cutesdr_bank_set_stereo(self.h, int(bool(on))))

    def block_length(self):
        n = C.c_int()
        check(self.L.cutesdr_bank_block_length(self.h, C.byref(n)))
        return n.value

    def launch_count(self):
        n = C.c_longlong()
        check(self.L.cutesdr_bank_launch_count(self.h, C.byref(n)))
        return n.value

    def kernel_timing(self, enable=True):
        check(self.L.cutesdr_bank_kernel_timing(self.h, int(bool(enable))))

    def kernel_model(self, which=0):
        """(kernel 1 runs on the tensor cores?, its GEMM flops per DSP block)."""
        t = C.c_int()
        f = C.c_double()
        check(self.L.cutesdr_bank_kernel_model(self.h, int(which), C.byref(t), C.byref(f)))
        return bool(t.value), f.value

    def kernel_time(self, which=0):
        ms, n = C.c_double(), C.c_longlong()
        check(self.L.cutesdr_bank_kernel_time(self.h, int(which), C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def last_block(self):
        """(device pointer, length) of the most recent DSP block after the noise blanker."""
        p, n = C.c_void_p(), C.c_int()
        check(self.L.cutesdr_bank_last_block(self.h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def stream(self):
        s = C.c_void_p()
        check(self.L.cutesdr_bank_stream(self.h, C.byref(s)))
        return s.value

    def synchronize(self):
        check(self.L.cutesdr_bank_synchronize(self.h))

    def join(self):
        check(self.L.cutesdr_bank_join(self.h))

    # --- processing
    def ProcessData(self, iq, audio=None, audio_stride=None):
        """iq: complex64 host array (any length). Returns (audio[n_channels, stride] float32, n_out[n_channels])."""
        iq = np.ascontiguousarray(iq, dtype=np.complex64)
        if audio is None:
            L = self.block_length()
            audio_stride = audio_stride or (len(iq) // L + 2) * 2048 * 2
            audio = np.zeros((self.n_channels, audio_stride), dtype=np.float32)
        else:
            audio_stride = audio.shape[1]
        n_out = np.zeros(self.n_channels, dtype=np.int32)
        check(self.L.cutesdr_bank_process(self.h, len(iq), iq.ctypes.data, audio.ctypes.data, int(audio_stride),
                                          n_out.ctypes.data_as(C.POINTER(C.c_int))))
        return audio, n_out

    def ProcessRaw(self, data, fmt, audio_stride=None):
        """Wire-format ingest: data = int16 array [n,2] (fmt 1) or uint8 array [n*6] of packed int24 (fmt 2)."""
        data = np.ascontiguousarray(data)
        n = data.size // 2 if fmt == 1 else (data.size // 6 if fmt == 2 else data.size)
        L = self.block_length()
        audio_stride = audio_stride or (n // L + 2) * 2048 * 2
        audio = np.zeros((self.n_channels, audio_stride), dtype=np.float32)
        n_out = np.zeros(self.n_channels, dtype=np.int32)
        check(self.L.cutesdr_bank_process_raw(self.h, int(n), data.ctypes.data, int(fmt), audio.ctypes.data, int(audio_stride),
                                              n_out.ctypes.data_as(C.POINTER(C.c_int))))
        return audio, n_out

    def ProcessPackets(self, packets, packet_bytes, audio_stride=None):
        """The radio's UDP datagrams as received (uint8 array, n * packet_bytes): 1028-byte 16-bit or 1444-byte 24-bit
        packets with their 4-byte headers (CUdpThread::OnreadyRead, interface/netiobase.cpp:464-534)."""
        packets = np.ascontiguousarray(packets, dtype=np.uint8)
        npk = packets.size // packet_bytes
        n = npk * (256 if packet_bytes == 1028 else 240)
        L = self.block_length()
        audio_stride = audio_stride or (n // L + 2) * 2048 * 2
        audio = np.zeros((self.n_channels, audio_stride), dtype=np.float32)
        n_out = np.zeros(self.n_channels, dtype=np.int32)
        check(self.L.cutesdr_bank_process_packets(self.h, int(npk), packets.ctypes.data, int(packet_bytes), audio.ctypes.data,
                                                  int(audio_stride), n_out.ctypes.data_as(C.POINTER(C.c_int))))
        return audio, n_out

    def MissedPackets(self, reset=False):
        """m_MissedPackets of the UDP front end (interface/netiobase.cpp:487-496)"""
        v = C.c_longlong()
        check(self.L.cutesdr_bank_missed_packets(self.h, C.byref(v), int(bool(reset))))
        return v.value

    def process_ptr(self, n_in, iq_ptr, audio_ptr, audio_stride, n_out_arr=None):
        """Raw-pointer form (pinned host memory) used by bench.py's end-to-end leg."""
        p = n_out_arr.ctypes.data_as(C.POINTER(C.c_int)) if n_out_arr is not None else None
        return check(self.L.cutesdr_bank_process(self.h, int(n_in), iq_ptr, audio_ptr, int(audio_stride), p))

    def process_async_ptr(self, n_in, iq_ptr, audio_ptr, audio_stride, n_out_arr=None):
        """Pipelined one-block form (pinned host pointers); results valid after synchronize()."""
        p = n_out_arr.ctypes.data_as(C.POINTER(C.c_int)) if n_out_arr is not None else None
        return check(self.L.cutesdr_bank_process_async(self.h, int(n_in), iq_ptr, audio_ptr, int(audio_stride), p))

    def process_async_device_ptr(self, n_in, d_iq_ptr, src_stream, audio_ptr, audio_stride, n_out_arr=None):
        """Pipelined form for a block already in device memory (stream-ordered with src_stream, a cudaStream_t)."""
        p = n_out_arr.ctypes.data_as(C.POINTER(C.c_int)) if n_out_arr is not None else None
        return check(self.L.cutesdr_bank_process_async_device(self.h, int(n_in), d_iq_ptr, src_stream, audio_ptr, int(audio_stride), p))

    def process_async_raw_ptr(self, n_in, data_ptr, fmt, audio_ptr, audio_stride, n_out_arr=None):
        """Pipelined one-block form for wire-format samples (fmt 1 = int16 pairs, 2 = packed int24)."""
        p = n_out_arr.ctypes.data_as(C.POINTER(C.c_int)) if n_out_arr is not None else None
        return check(self.L.cutesdr_bank_process_async_raw(self.h, int(n_in), data_ptr, int(fmt), audio_ptr, int(audio_stride), p))

    def process_async_bcast_ptr(self, mgpu, n_in, iq_ptr_rank0, fmt, audio_ptr, audio_stride, n_out_arr=None):
        """Multi-GPU pipelined form: rank 0 hands in the pinned host block, the library broadcasts it (NCCL)."""
        p = n_out_arr.ctypes.data_as(C.POINTER(C.c_int)) if n_out_arr is not None else None
        return check(self.L.cutesdr_bank_process_async_bcast(self.h, mgpu.h, int(n_in), iq_ptr_rank0, int(fmt), audio_ptr,
                                                             int(audio_stride), p))

    def process_device(self, d_iq_ptr, n_in, d_audio_ptr=None, audio_stride=0, fmt=0):
        m = C.c_int()
        check(self.L.cutesdr_bank_process_device_raw(self.h, d_iq_ptr, int(fmt), int(n_in), d_audio_ptr, int(audio_stride), C.byref(m)))
        return m.value

    # --- test-bench taps (PROFILE_1..4)
    def tap_enable(self, ch, profiles):
        mask = 0
        for p in profiles:
            mask |= 1 << p
        check(self.L.cutesdr_bank_tap_enable(self.h, int(ch), mask))

    def tap_read(self, ch, profile):
        n = C.c_long()
        check(self.L.cutesdr_bank_tap_size(self.h, int(ch), int(profile), C.byref(n)))
        out = np.empty(n.value, dtype=np.float32)
        if n.value:
            check(self.L.cutesdr_bank_tap_read(self.h, int(ch), int(profile), out.ctypes.data_as(C.POINTER(C.c_float)), n.value))
        if profile < 4:
            return out[0::2].astype(np.complex64) + 1j * out[1::2].astype(np.complex64)
        return out


    def tap_spectrum(self, ch, profile, fft, display_rate=10):
        """CTestBench's spectrum of a PROFILE tap (gui/testbench.cpp:583-611) on the device; fft=None detaches."""
        check(self.L.cutesdr_bank_tap_spectrum(self.h, int(ch), int(profile), fft.h if fft is not None else None, int(display_rate)))
        if fft is not None:
            fft._size = 2048

    def tap_spectrum_frames(self, ch, profile):
        n = C.c_longlong()
        check(self.L.cutesdr_bank_tap_spectrum_frames(self.h, int(ch), int(profile), C.byref(n)))
        return n.value


class MultiGpu:
    """cutesdr_mgpu: this process's rank in the multi-GPU receiver (one process per GPU). The 128-byte NCCL id is
    created by rank 0 (`MultiGpu.unique_id()`) and handed to the other processes by the host application."""

    def __init__(self, id_bytes, rank, world, device):
        self.L = load_library()
        h = C.c_void_p()
        buf = C.create_string_buffer(bytes(id_bytes), 128) if id_bytes is not None else None
        check(self.L.cutesdr_mgpu_init(C.byref(h), buf, int(rank), int(world), int(device)))
        self.h = h
        self.rank, self.world = int(rank), int(world)

    @staticmethod
    def unique_id():
        L = load_library()
        buf = C.create_string_buffer(128)
        check(L.cutesdr_mgpu_unique_id(buf))
        return buf.raw

    def info(self):
        r, w, v = C.c_int(), C.c_int(), C.c_int()
        nb, by = C.c_longlong(), C.c_longlong()
        check(self.L.cutesdr_mgpu_info(self.h, C.byref(r), C.byref(w), C.byref(v), C.byref(nb), C.byref(by)))
        return {"rank": r.value, "world": w.value, "nccl_version": v.value, "blocks": nb.value, "bytes_broadcast": by.value}

    def close(self):
        if getattr(self, "h", None):
            self.L.cutesdr_mgpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def channel_slice(n_channels, rank, world):
    """(first, count) of the contiguous channel slice rank `rank` owns (cutesdr_mgpu_channel_slice)."""
    f, n = C.c_int(), C.c_int()
    check(load_library().cutesdr_mgpu_channel_slice(int(n_channels), int(rank), int(world), C.byref(f), C.byref(n)))
    return f.value, n.value


def device_memory(device=0):
    """(free, total) bytes of the device, cudaMemGetInfo."""
    f, t = C.c_longlong(), C.c_longlong()
    check(load_library().cutesdr_device_memory(int(device), C.byref(f), C.byref(t)))
    return f.value, t.value


def microbench(which, device=0):
    """0 FP32 FMA TFLOP/s, 1 tcgen05 tf32 TFLOP/s, 2 tcgen05 f16 TFLOP/s, 3 HBM copy GB/s (cutesdr_microbench)."""
    v = C.c_double()
    check(load_library().cutesdr_microbench(int(device), int(which), C.byref(v)))
    return v.value


class CDemodulator(_Handle):
    """dsp/demodulator.h:56-100 (mono path)."""
    _create, _destroy = "cutesdr_demodulator_create", "cutesdr_demodulator_destroy"

    def SetInputSampleRate(self, rate):
        check(self.L.cutesdr_demodulator_set_input_sample_rate(self.h, float(rate)))

    def SetDemod(self, mode, info):
        s = _info_struct(info)
        check(self.L.cutesdr_demodulator_set_demod(self.h, int(mode), C.byref(s)))

    def SetDemodFreq(self, freq):
        check(self.L.cutesdr_demodulator_set_demod_freq(self.h, float(freq)))

    def GetOutputRate(self):
        r = C.c_double()
        check(self.L.cutesdr_demodulator_get_output_rate(self.h, C.byref(r)))
        return r.value

    def GetSMeterPeak(self):
        p = C.c_double()
        check(self.L.cutesdr_demodulator_get_smeter(self.h, C.byref(p), None))
        return p.value

    def GetSMeterAve(self):
        a = C.c_double()        # a NULL peak pointer leaves the held peak alone, like the reference's GetAve
        check(self.L.cutesdr_demodulator_get_smeter(self.h, None, C.byref(a)))
        return a.value

    def ProcessData(self, iq, stereo=False):
        buf = _cpx_to_f64(iq)
        if stereo:
            out = np.empty(2 * (len(iq) // 4 + 8192), dtype=np.float64)
            n = check(self.L.cutesdr_demodulator_process_stereo(self.h, len(iq), _dptr(buf), _dptr(out)))
            return _f64_to_cpx(out, n)
        out = np.empty(len(iq) // 4 + 8192, dtype=np.float64)
        n = check(self.L.cutesdr_demodulator_process(self.h, len(iq), _dptr(buf), _dptr(out)))
        return out[:n].copy()


class CDownConvert(_Handle):
    """dsp/downconvert.h:23-120."""
    _create, _destroy = "cutesdr_downconvert_create", "cutesdr_downconvert_destroy"

    def SetFrequency(self, f):
        check(self.L.cutesdr_downconvert_set_frequency(self.h, float(f)))

    def SetCwOffset(self, f):
        check(self.L.cutesdr_downconvert_set_cw_offset(self.h, float(f)))

    def SetDataRate(self, in_rate, max_bw):
        r = C.c_double()
        check(self.L.cutesdr_downconvert_set_data_rate(self.h, float(in_rate), float(max_bw), C.byref(r)))
        return r.value

    def stages(self):
        a = np.zeros(32, dtype=np.int32)
        n = C.c_int()
        check(self.L.cutesdr_downconvert_stages(self.h, a.ctypes.data_as(C.POINTER(C.c_int)), 32, C.byref(n)))
        return [int(v) for v in a[:n.value]]

    def ProcessData(self, x):
        buf = _cpx_to_f64(x)
        out = np.empty_like(buf)
        n = check(self.L.cutesdr_downconvert_process(self.h, len(x), _dptr(buf), _dptr(out)))
        return _f64_to_cpx(out, n)


class CFastFIR(_Handle):
    """dsp/fastfir.h:19-44."""
    _create, _destroy = "cutesdr_fastfir_create", "cutesdr_fastfir_destroy"

    def SetupParameters(self, lo, hi, offset, rate):
        check(self.L.cutesdr_fastfir_setup(self.h, float(lo), float(hi), float(offset), float(rate)))

    def ProcessData(self, x):
        buf = _cpx_to_f64(x)
        out = np.empty(2 * (len(x) + 2048), dtype=np.float64)
        n = check(self.L.cutesdr_fastfir_process(self.h, len(x), _dptr(buf), _dptr(out)))
        return _f64_to_cpx(out, n)


class CFft(_Handle):
    """dsp/fft.h:24-85 (display path)."""
    _create, _destroy = "cutesdr_fft_create", "cutesdr_fft_destroy"

    def SetFFTParams(self, size, invert, db_comp, sample_freq):
        check(self.L.cutesdr_fft_set_params(self.h, int(size), int(bool(invert)), float(db_comp), float(sample_freq)))
        self._size = min(max(int(size), 512), 65536)

    def SetFFTAve(self, ave):
        check(self.L.cutesdr_fft_set_ave(self.h, int(ave)))

    def ResetFFT(self):
        check(self.L.cutesdr_fft_reset(self.h))

    def PutInDisplayFFT(self, x):
        x = np.asarray(x)
        t = C.c_int()
        if x.dtype == np.complex64:
            x = np.ascontiguousarray(x)
            check(self.L.cutesdr_fft_put_f32(self.h, len(x), x.ctypes.data_as(C.POINTER(C.c_float)), C.byref(t)))
        else:
            buf = _cpx_to_f64(x)
            check(self.L.cutesdr_fft_put(self.h, len(x), _dptr(buf), C.byref(t)))
        return t.value

    def put_device(self, d_ptr, n):
        t = C.c_int()
        check(self.L.cutesdr_fft_put_device(self.h, int(n), d_ptr, C.byref(t)))
        return t.value

    def put_device_async(self, d_ptr, n, src_stream):
        """stream-ordered PutInDisplayFFT of a device frame (no host synchronisation); src_stream: cudaStream_t"""
        t = C.c_int()
        check(self.L.cutesdr_fft_put_device_async(self.h, int(n), d_ptr, src_stream, C.byref(t)))
        return t.value

    def GetScreenIntegerFFTData(self, max_h, max_w, max_db, min_db, start, stop):
        out = np.zeros(max_w, dtype=np.int32)
        ov = C.c_int()
        check(self.L.cutesdr_fft_get_screen(self.h, int(max_h), int(max_w), float(max_db), float(min_db), int(start),
                                            int(stop), out.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(ov)))
        return bool(ov.value), out

    def GetPlot(self, trace_h, width, max_db, min_db, start, stop):
        """CPlotter::draw's waterfall row (255 levels) and 2-D trace from one pass (gui/plotter.cpp:429-456)."""
        wf = np.zeros(width, dtype=np.int32)
        tr = np.zeros(width, dtype=np.int32)
        ov = C.c_int()
        check(self.L.cutesdr_fft_get_plot(self.h, int(trace_h), int(width), float(max_db), float(min_db), int(start), int(stop),
                                          wf.ctypes.data_as(C.POINTER(C.c_int32)), tr.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(ov)))
        return bool(ov.value), wf, tr

    def SetDcOffset(self, off_i, off_q):
        """m_NCOSpurOffsetI/Q of the display copy (interface/sdrinterface.cpp:889-894)."""
        check(self.L.cutesdr_fft_set_dc_offset(self.h, float(off_i), float(off_q)))

    def FwdFFT(self, x):
        buf = _cpx_to_f64(x)
        check(self.L.cutesdr_fft_fwd(self.h, _dptr(buf)))
        return _f64_to_cpx(buf, len(x))

    def RevFFT(self, x):
        buf = _cpx_to_f64(x)
        check(self.L.cutesdr_fft_rev(self.h, _dptr(buf)))
        return _f64_to_cpx(buf, len(x))

    def avebuf(self):
        out = np.empty(getattr(self, "_size", 2048), dtype=np.float32)
        n = check(self.L.cutesdr_fft_get_ave(self.h, out.ctypes.data_as(C.POINTER(C.c_float)), len(out)))
        return out[:n]

    def launch_count(self):
        n = C.c_longlong()
        check(self.L.cutesdr_fft_launch_count(self.h, C.byref(n)))
        return n.value


class CAgc(_Handle):
    """dsp/agc.h:18-63 (complex path)."""
    _create, _destroy = "cutesdr_agc_create", "cutesdr_agc_destroy"

    def SetParameters(self, on, hang, thresh, manual_gain, slope, decay, rate):
        check(self.L.cutesdr_agc_set_parameters(self.h, int(bool(on)), int(bool(hang)), int(thresh), int(manual_gain),
                                                int(slope), int(decay), float(rate)))

    def ProcessData(self, x):
        buf = _cpx_to_f64(x)
        out = np.empty_like(buf)
        check(self.L.cutesdr_agc_process(self.h, len(x), _dptr(buf), _dptr(out)))
        return _f64_to_cpx(out, len(x))


class CFractResampler(_Handle):
    """dsp/fractresampler.h:16-33."""
    _create, _destroy = "cutesdr_resampler_create", "cutesdr_resampler_destroy"

    def Init(self, max_input):
        check(self.L.cutesdr_resampler_init(self.h, int(max_input)))

    def Resample(self, x, rate, gain=None):
        x = np.asarray(x)
        cap = int(len(x) / rate) + 64
        if np.iscomplexobj(x):
            buf = _cpx_to_f64(x)
            if gain is None:
                out = np.empty(2 * cap, dtype=np.float64)
                n = check(self.L.cutesdr_resampler_cpx(self.h, len(x), float(rate), _dptr(buf), _dptr(out)))
                return _f64_to_cpx(out, n)
            out = np.empty(2 * cap, dtype=np.int16)
            n = check(self.L.cutesdr_resampler_stereo16(self.h, len(x), float(rate), _dptr(buf),
                                                        out.ctypes.data_as(C.POINTER(C.c_int16)), float(gain)))
            return out[:2 * n].reshape(n, 2)
        buf = np.ascontiguousarray(x, dtype=np.float64)
        if gain is None:
            out = np.empty(cap, dtype=np.float64)
            n = check(self.L.cutesdr_resampler_real(self.h, len(x), float(rate), _dptr(buf), _dptr(out)))
            return out[:n]
        out = np.empty(cap, dtype=np.int16)
        n = check(self.L.cutesdr_resampler_mono16(self.h, len(x), float(rate), _dptr(buf),
                                                  out.ctypes.data_as(C.POINTER(C.c_int16)), float(gain)))
        return out[:n]


class CNoiseProc(_Handle):
    """dsp/noiseproc.h:23-53."""
    _create, _destroy = "cutesdr_noiseproc_create", "cutesdr_noiseproc_destroy"

    def SetupBlanker(self, on, threshold, width_us, rate):
        check(self.L.cutesdr_noiseproc_setup(self.h, int(bool(on)), float(threshold), float(width_us), float(rate)))

    def ProcessBlanker(self, x):
        buf = _cpx_to_f64(x)
        out = np.empty_like(buf)
        check(self.L.cutesdr_noiseproc_process(self.h, len(x), _dptr(buf), _dptr(out)))
        return _f64_to_cpx(out, len(x))


class CIir(_Handle):
    """dsp/iir.h:16-40 -- one direct-form-2 biquad (InitLP/HP/BP/BR, ProcessFilter real or complex)."""
    _create, _destroy = "cutesdr_iir_create", "cutesdr_iir_destroy"

    def _init(self, kind, f0, q, rate):
        check(self.L.cutesdr_iir_init(self.h, kind, float(f0), float(q), float(rate)))

    def InitLP(self, f0, q, rate):
        self._init(0, f0, q, rate)

    def InitHP(self, f0, q, rate):
        self._init(1, f0, q, rate)

    def InitBP(self, f0, q, rate):
        self._init(2, f0, q, rate)

    def InitBR(self, f0, q, rate):
        self._init(3, f0, q, rate)

    def ProcessFilter(self, x):
        x = np.asarray(x)
        if np.iscomplexobj(x):
            buf = _cpx_to_f64(x)
            out = np.empty_like(buf)
            check(self.L.cutesdr_iir_process_cpx(self.h, len(x), _dptr(buf), _dptr(out)))
            return _f64_to_cpx(out, len(x))
        buf = np.ascontiguousarray(x, dtype=np.float64)
        out = np.empty_like(buf)
        check(self.L.cutesdr_iir_process_real(self.h, len(buf), _dptr(buf), _dptr(out)))
        return out
