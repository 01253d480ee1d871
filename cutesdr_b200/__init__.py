"""cutesdr_b200 -- B200 (sm_100a) implementation of CuteSDR's receive DSP chain as a batched
multi-channel receiver. The product is the C-ABI library `libcutesdr_cuda.so`
(include/cutesdr_cuda.h); this package is its Python host-side mirror of the reference's
operator API (class names and argument meaning follow dsp/*.h).

There is no CPU fallback: importing the package works anywhere, but every object needs the
compiled CUDA library and a CUDA device, and fails loudly otherwise.
"""
from .modes import (DEMOD_AM, DEMOD_SAM, DEMOD_FM, DEMOD_USB, DEMOD_LSB, DEMOD_CWU, DEMOD_CWL, NUM_DEMODS,
                    MODE_NAMES, demod_info, max_bandwidth)
from .lib import load_library, library_path, CuteSdrError
from .dsp import (ReceiverBank, MultiGpu, channel_slice, microbench, device_memory, CDemodulator, CDownConvert, CFastFIR, CFft, CAgc, CFractResampler, CNoiseProc, CIir)

__all__ = ["DEMOD_AM", "DEMOD_SAM", "DEMOD_FM", "DEMOD_USB", "DEMOD_LSB", "DEMOD_CWU", "DEMOD_CWL", "NUM_DEMODS",
           "MODE_NAMES", "demod_info", "max_bandwidth", "load_library", "library_path", "CuteSdrError",
           "ReceiverBank", "MultiGpu", "channel_slice", "microbench", "device_memory", "CDemodulator", "CDownConvert", "CFastFIR", "CFft", "CAgc", "CFractResampler",
           "CNoiseProc", "CIir"]
