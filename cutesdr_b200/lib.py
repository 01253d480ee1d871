"""ctypes loader for libcutesdr_cuda.so and the prototypes of every symbol include/cutesdr_cuda.h
declares. No compute happens at import or load time."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))


class CuteSdrError(RuntimeError):
    pass


class DemodInfo(C.Structure):
    """cutesdr_demod_info (tDemodInfo, dsp/demodulator.h:35-54)"""
    _fields_ = [(k, C.c_int) for k in ("HiCut", "HiCutmin", "HiCutmax", "LowCut", "LowCutmin", "LowCutmax", "Offset",
                                        "SquelchValue", "AgcSlope", "AgcThresh", "AgcManualGain", "AgcDecay", "AgcOn",
                                        "AgcHangOn")]


def library_path():
    return os.path.join(_HERE, "libcutesdr_cuda.so")


_vp = C.c_void_p
_pp = C.POINTER(C.c_void_p)
_dp = C.POINTER(C.c_double)
_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int)
_i16 = C.POINTER(C.c_int16)
_i32 = C.POINTER(C.c_int32)
_info = C.POINTER(DemodInfo)

# name -> (restype, argtypes): one entry per declaration in include/cutesdr_cuda.h
PROTOTYPES = {
    "cutesdr_last_error": (C.c_char_p, []),
    "cutesdr_version": (C.c_char_p, []),
    "cutesdr_device_count": (C.c_int, [_ip]),
    "cutesdr_microbench": (C.c_int, [C.c_int, C.c_int, _dp]),
    "cutesdr_host_alloc": (C.c_int, [_pp, C.c_size_t]),
    "cutesdr_host_free": (None, [_vp]),
    "cutesdr_device_memory": (C.c_int, [C.c_int, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]),
    "cutesdr_bank_create": (C.c_int, [_pp, C.c_int, C.c_double, C.c_int]),
    "cutesdr_bank_destroy": (None, [_vp]),
    "cutesdr_bank_set_demod": (C.c_int, [_vp, C.c_int, C.c_int, _info]),
    "cutesdr_bank_set_demod_freq": (C.c_int, [_vp, C.c_int, C.c_double]),
    "cutesdr_bank_get_output_rate": (C.c_int, [_vp, C.c_int, _dp]),
    "cutesdr_bank_block_length": (C.c_int, [_vp, _ip]),
    "cutesdr_bank_get_smeter": (C.c_int, [_vp, C.c_int, _dp, _dp]),
    "cutesdr_bank_set_noiseproc": (C.c_int, [_vp, C.c_int, C.c_double, C.c_double]),
    "cutesdr_bank_set_audio_rate": (C.c_int, [_vp, C.c_double]),
    "cutesdr_bank_set_stereo": (C.c_int, [_vp, C.c_int]),
    "cutesdr_bank_process": (C.c_int, [_vp, C.c_int, _vp, _vp, C.c_int, _ip]),
    "cutesdr_bank_process_async": (C.c_int, [_vp, C.c_int, _vp, _vp, C.c_int, _ip]),
    "cutesdr_bank_process_raw": (C.c_int, [_vp, C.c_int, _vp, C.c_int, _vp, C.c_int, _ip]),
    "cutesdr_bank_process_async_raw": (C.c_int, [_vp, C.c_int, _vp, C.c_int, _vp, C.c_int, _ip]),
    "cutesdr_bank_process_packets": (C.c_int, [_vp, C.c_int, _vp, C.c_int, _vp, C.c_int, _ip]),
    "cutesdr_bank_missed_packets": (C.c_int, [_vp, C.POINTER(C.c_longlong), C.c_int]),
    "cutesdr_bank_process_async_device": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, C.c_int, _ip]),
    "cutesdr_mgpu_unique_id": (C.c_int, [_vp]),
    "cutesdr_mgpu_init": (C.c_int, [_pp, _vp, C.c_int, C.c_int, C.c_int]),
    "cutesdr_mgpu_destroy": (None, [_vp]),
    "cutesdr_mgpu_info": (C.c_int, [_vp, _ip, _ip, _ip, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]),
    "cutesdr_mgpu_channel_slice": (C.c_int, [C.c_int, C.c_int, C.c_int, _ip, _ip]),
    "cutesdr_bank_process_async_bcast": (C.c_int, [_vp, _vp, C.c_int, _vp, C.c_int, _vp, C.c_int, _ip]),
    "cutesdr_bank_process_device": (C.c_int, [_vp, _vp, C.c_int, _vp, C.c_int, _ip]),
    "cutesdr_bank_process_device_raw": (C.c_int, [_vp, _vp, C.c_int, C.c_int, _vp, C.c_int, _ip]),
    "cutesdr_bank_synchronize": (C.c_int, [_vp]),
    "cutesdr_bank_join": (C.c_int, [_vp]),
    "cutesdr_bank_last_block": (C.c_int, [_vp, _pp, _ip]),
    "cutesdr_bank_stream": (C.c_int, [_vp, _pp]),
    "cutesdr_bank_launch_count": (C.c_int, [_vp, C.POINTER(C.c_longlong)]),
    "cutesdr_bank_kernel_timing": (C.c_int, [_vp, C.c_int]),
    "cutesdr_bank_kernel_time": (C.c_int, [_vp, C.c_int, _dp, C.POINTER(C.c_longlong)]),
    "cutesdr_bank_kernel_model": (C.c_int, [_vp, C.c_int, _ip, _dp]),
    "cutesdr_bank_tap_enable": (C.c_int, [_vp, C.c_int, C.c_uint]),
    "cutesdr_bank_tap_spectrum": (C.c_int, [_vp, C.c_int, C.c_int, _vp, C.c_int]),
    "cutesdr_bank_tap_spectrum_frames": (C.c_int, [_vp, C.c_int, C.c_int, C.POINTER(C.c_longlong)]),
    "cutesdr_bank_tap_size": (C.c_int, [_vp, C.c_int, C.c_int, C.POINTER(C.c_long)]),
    "cutesdr_bank_tap_read": (C.c_int, [_vp, C.c_int, C.c_int, _fp, C.c_long]),
    "cutesdr_downconvert_create": (C.c_int, [_pp, C.c_int]),
    "cutesdr_downconvert_destroy": (None, [_vp]),
    "cutesdr_downconvert_set_frequency": (C.c_int, [_vp, C.c_double]),
    "cutesdr_downconvert_set_cw_offset": (C.c_int, [_vp, C.c_double]),
    "cutesdr_downconvert_set_data_rate": (C.c_int, [_vp, C.c_double, C.c_double, _dp]),
    "cutesdr_downconvert_stages": (C.c_int, [_vp, _ip, C.c_int, _ip]),
    "cutesdr_downconvert_process": (C.c_int, [_vp, C.c_int, _dp, _dp]),
    "cutesdr_downconvert_process_f32": (C.c_int, [_vp, C.c_int, _fp, _fp]),
    "cutesdr_fastfir_create": (C.c_int, [_pp, C.c_int]),
    "cutesdr_fastfir_destroy": (None, [_vp]),
    "cutesdr_fastfir_setup": (C.c_int, [_vp, C.c_double, C.c_double, C.c_double, C.c_double]),
    "cutesdr_fastfir_process": (C.c_int, [_vp, C.c_int, _dp, _dp]),
    "cutesdr_fastfir_process_f32": (C.c_int, [_vp, C.c_int, _fp, _fp]),
    "cutesdr_fft_create": (C.c_int, [_pp, C.c_int]),
    "cutesdr_fft_destroy": (None, [_vp]),
    "cutesdr_fft_set_params": (C.c_int, [_vp, C.c_int, C.c_int, C.c_double, C.c_double]),
    "cutesdr_fft_set_ave": (C.c_int, [_vp, C.c_int]),
    "cutesdr_fft_reset": (C.c_int, [_vp]),
    "cutesdr_fft_put": (C.c_int, [_vp, C.c_int, _dp, _ip]),
    "cutesdr_fft_put_f32": (C.c_int, [_vp, C.c_int, _fp, _ip]),
    "cutesdr_fft_put_device": (C.c_int, [_vp, C.c_int, _vp, _ip]),
    "cutesdr_fft_put_device_async": (C.c_int, [_vp, C.c_int, _vp, _vp, _ip]),
    "cutesdr_fft_launch_count": (C.c_int, [_vp, C.POINTER(C.c_longlong)]),
    "cutesdr_fft_get_screen": (C.c_int, [_vp, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, _i32, _ip]),
    "cutesdr_fft_get_plot": (C.c_int, [_vp, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, _i32, _i32, _ip]),
    "cutesdr_fft_set_dc_offset": (C.c_int, [_vp, C.c_double, C.c_double]),
    "cutesdr_fft_fwd": (C.c_int, [_vp, _dp]),
    "cutesdr_fft_rev": (C.c_int, [_vp, _dp]),
    "cutesdr_fft_get_ave": (C.c_int, [_vp, _fp, C.c_int]),
    "cutesdr_agc_create": (C.c_int, [_pp, C.c_int]),
    "cutesdr_agc_destroy": (None, [_vp]),
    "cutesdr_agc_set_parameters": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double]),
    "cutesdr_agc_process": (C.c_int, [_vp, C.c_int, _dp, _dp]),
    "cutesdr_resampler_create": (C.c_int, [_pp, C.c_int]),
    "cutesdr_resampler_destroy": (None, [_vp]),
    "cutesdr_resampler_init": (C.c_int, [_vp, C.c_int]),
    "cutesdr_resampler_real": (C.c_int, [_vp, C.c_int, C.c_double, _dp, _dp]),
    "cutesdr_resampler_cpx": (C.c_int, [_vp, C.c_int, C.c_double, _dp, _dp]),
    "cutesdr_resampler_mono16": (C.c_int, [_vp, C.c_int, C.c_double, _dp, _i16, C.c_double]),
    "cutesdr_resampler_stereo16": (C.c_int, [_vp, C.c_int, C.c_double, _dp, _i16, C.c_double]),
    "cutesdr_iir_create": (C.c_int, [_pp, C.c_int]),
    "cutesdr_iir_destroy": (None, [_vp]),
    "cutesdr_iir_init": (C.c_int, [_vp, C.c_int, C.c_double, C.c_double, C.c_double]),
    "cutesdr_iir_process_real": (C.c_int, [_vp, C.c_int, _dp, _dp]),
    "cutesdr_iir_process_cpx": (C.c_int, [_vp, C.c_int, _dp, _dp]),
    "cutesdr_noiseproc_create": (C.c_int, [_pp, C.c_int]),
    "cutesdr_noiseproc_destroy": (None, [_vp]),
    "cutesdr_noiseproc_setup": (C.c_int, [_vp, C.c_int, C.c_double, C.c_double, C.c_double]),
    "cutesdr_noiseproc_process": (C.c_int, [_vp, C.c_int, _dp, _dp]),
    "cutesdr_noiseproc_process_f32": (C.c_int, [_vp, C.c_int, _fp, _fp]),
    "cutesdr_demodulator_create": (C.c_int, [_pp, C.c_int]),
    "cutesdr_demodulator_destroy": (None, [_vp]),
    "cutesdr_demodulator_set_input_sample_rate": (C.c_int, [_vp, C.c_double]),
    "cutesdr_demodulator_set_demod": (C.c_int, [_vp, C.c_int, _info]),
    "cutesdr_demodulator_set_demod_freq": (C.c_int, [_vp, C.c_double]),
    "cutesdr_demodulator_get_output_rate": (C.c_int, [_vp, _dp]),
    "cutesdr_demodulator_get_smeter": (C.c_int, [_vp, _dp, _dp]),
    "cutesdr_demodulator_process": (C.c_int, [_vp, C.c_int, _dp, _dp]),
    "cutesdr_demodulator_process_stereo": (C.c_int, [_vp, C.c_int, _dp, _dp]),
}

_lib = None


def load_library():
    """Loads libcutesdr_cuda.so (built by `python -m cutesdr_b200.build` / __graft_entry__.build()).
    Raises CuteSdrError if it is missing -- there is no fallback implementation."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise CuteSdrError("%s not found: build it with `python -m cutesdr_b200.build` (no CPU fallback exists)" % path)
    lib = C.CDLL(path)
    for name, (res, args) in PROTOTYPES.items():
        f = getattr(lib, name)          # AttributeError if the library does not export a declared symbol
        f.restype = res
        f.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc < 0:
        msg = load_library().cutesdr_last_error()
        raise CuteSdrError("libcutesdr_cuda error %d: %s" % (rc, msg.decode() if msg else "?"))
    return rc
