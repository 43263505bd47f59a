"""ctypes binding of liboctb200.so (include/octb200.h).

The CUDA library is the product; there is no CPU fallback.  Importing this module never needs a
GPU (the loader only dlopens the library and checks the exported symbols), but every compute entry
point fails loudly if the library or a B200 is missing.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("OCTB200_LIB", os.path.join(_HERE, "liboctb200.so"))   # override: development A/B builds only

# every symbol include/octb200.h declares (tests check the list against the header)
SYMBOLS = [
    "octb200_create", "octb200_destroy", "octb200_last_error", "octb200_version", "octb200_default_params",
    "octb200_effective_fft_mode", "octb200_query_fft_path", "octb200_set_params", "octb200_set_resample_curve", "octb200_set_dispersion_curve",
    "octb200_set_window_curve", "octb200_set_postprocess_background", "octb200_get_postprocess_background",
    "octb200_get_fpn_mean_line", "octb200_set_fpn_mean_line", "octb200_get_fpn_segment_stats", "octb200_make_resample_curve",
    "octb200_make_dispersion_curve", "octb200_make_window_curve", "octb200_make_sinusoidal_curve",
    "octb200_register_host_buffers", "octb200_unregister_host_buffers", "octb200_register_streaming_buffers",
    "octb200_unregister_streaming_buffers", "octb200_register_float_streaming_buffers",
    "octb200_unregister_float_streaming_buffers", "octb200_set_callbacks", "octb200_process_host",
    "octb200_process_device", "octb200_sync", "octb200_current_buffer_nr", "octb200_output_device_ptr",
    "octb200_copy_output", "octb200_bind_output", "octb200_bscan_frame", "octb200_enface_frame",
    "octb200_volume_u8", "octb200_float_to_output", "octb200_compute_stream", "octb200_event_record",
    "octb200_event_elapsed_ms", "octb200_launch_count", "octb200_time_kernel",
    "octb200_enface_gather_init", "octb200_enface_gather_connect", "octb200_enface_gather", "octb200_enface_gather_wait",
    "octb200_enface_gather_close", "octb200_enface_gather_auto", "octb200_enface_gather_status", "octb200_dispersion_sweep",
]
IPC_HANDLE_BYTES = 64

OK, ERR_INVALID, ERR_CUDA, ERR_NOMEM, ERR_NOT_READY = 0, -1, -2, -3, -4
FFT_AUTO, FFT_FUSED, FFT_SPLIT, FFT_CUFFT = 0, 1, 2, 3
INTERP_LINEAR, INTERP_CUBIC, INTERP_LANCZOS = 0, 1, 2
PACK_CONTAINER, PACK_12P = 0, 1
FLAG_SEPARATE_CONVERSION = 1
FLAG_NO_DEPENDENT_LAUNCH = 2
PATH_REGISTER_KERNEL, PATH_SHARED_MEMORY_KERNEL, PATH_CUFFT_CHAIN, PATH_CUFFT_CHAIN_SHARED_AVAILABLE = 1, 2, 3, 4


class Config(C.Structure):
    _fields_ = [("samplesPerLine", C.c_uint32), ("ascansPerBscan", C.c_uint32), ("bscansPerBuffer", C.c_uint32),
                ("buffersPerVolume", C.c_uint32), ("bitDepth", C.c_uint32), ("device", C.c_int32),
                ("rawSlots", C.c_int32), ("fftMode", C.c_int32), ("bscanIndexBase", C.c_uint32),
                ("inputPacking", C.c_uint32), ("flags", C.c_uint32), ("bscansInUnshardedBuffer", C.c_uint32)]


class Params(C.Structure):
    _fields_ = [("bitshift", C.c_int32), ("bscanFlip", C.c_int32), ("signalLogScaling", C.c_int32),
                ("sinusoidalScanCorrection", C.c_int32), ("signalGrayscaleMin", C.c_float),
                ("signalGrayscaleMax", C.c_float), ("signalMultiplicator", C.c_float), ("signalAddend", C.c_float),
                ("backgroundRemoval", C.c_int32), ("rollingAverageWindowSize", C.c_int32), ("resampling", C.c_int32),
                ("resamplingInterpolation", C.c_int32), ("dispersionCompensation", C.c_int32), ("windowing", C.c_int32),
                ("fixedPatternNoiseRemoval", C.c_int32), ("continuousFixedPatternNoiseDetermination", C.c_int32),
                ("redetermineFixedPatternNoise", C.c_int32), ("bscansForNoiseDetermination", C.c_uint32),
                ("postProcessBackgroundRemoval", C.c_int32), ("postProcessBackgroundRecordingRequested", C.c_int32),
                ("postProcessBackgroundWeight", C.c_float), ("postProcessBackgroundOffset", C.c_float),
                ("streamToHost", C.c_int32), ("streamingBuffersToSkip", C.c_uint32), ("streamFloatToHost", C.c_int32),
                ("reserved", C.c_uint32 * 3)]


class SweepConfig(C.Structure):
    _fields_ = [("lines", C.c_uint32), ("trials", C.c_uint32), ("metric", C.c_int32), ("metricThreshold", C.c_float),
                ("samplesToIgnore", C.c_int32), ("logScale", C.c_int32), ("logMin", C.c_float), ("logMax", C.c_float),
                ("logCoeff", C.c_float), ("logAddend", C.c_float), ("reserved", C.c_uint32 * 4)]


METRIC_SUM_ABOVE_THRESHOLD, METRIC_SAMPLES_ABOVE_THRESHOLD, METRIC_PEAK_VALUE, METRIC_MEAN_SOBEL = 0, 1, 2, 3
HOST_CALLBACK = C.CFUNCTYPE(None, C.c_void_p)

_lib = None


class Octb200Error(RuntimeError):
    pass


def load() -> C.CDLL:
    """dlopen the product library; raises (never falls back) when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise Octb200Error(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    missing = [s for s in SYMBOLS if not hasattr(lib, s)]
    if missing:
        raise Octb200Error(f"liboctb200.so lacks symbols: {missing}")
    P = C.c_void_p
    fp = C.POINTER(C.c_float)
    lib.octb200_create.argtypes = [C.POINTER(Config), C.POINTER(P)]
    lib.octb200_destroy.argtypes = [P]
    lib.octb200_last_error.argtypes = [P]; lib.octb200_last_error.restype = C.c_char_p
    lib.octb200_default_params.argtypes = [C.POINTER(Params)]; lib.octb200_default_params.restype = None
    lib.octb200_effective_fft_mode.argtypes = [P]
    lib.octb200_query_fft_path.argtypes = [C.c_uint32, C.c_uint32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    lib.octb200_set_params.argtypes = [P, C.POINTER(Params)]
    for n in ("octb200_set_resample_curve", "octb200_set_dispersion_curve", "octb200_set_window_curve",
              "octb200_set_postprocess_background", "octb200_get_postprocess_background",
              "octb200_get_fpn_mean_line", "octb200_set_fpn_mean_line"):
        getattr(lib, n).argtypes = [P, C.c_void_p, C.c_int]
    lib.octb200_get_fpn_segment_stats.argtypes = [P, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
    lib.octb200_make_resample_curve.argtypes = [C.c_int] + [C.c_float] * 4 + [C.c_void_p]
    lib.octb200_make_dispersion_curve.argtypes = [C.c_int] + [C.c_float] * 4 + [C.c_void_p]
    lib.octb200_make_window_curve.argtypes = [C.c_int, C.c_float, C.c_float, C.c_int, C.c_void_p]
    lib.octb200_make_sinusoidal_curve.argtypes = [C.c_int, C.c_void_p]
    lib.octb200_register_host_buffers.argtypes = [P, C.c_void_p, C.c_void_p]
    lib.octb200_unregister_host_buffers.argtypes = [P]
    lib.octb200_register_streaming_buffers.argtypes = [P, C.c_void_p, C.c_void_p, C.c_size_t]
    lib.octb200_unregister_streaming_buffers.argtypes = [P]
    lib.octb200_register_float_streaming_buffers.argtypes = [P, C.c_void_p, C.c_void_p, C.c_size_t]
    lib.octb200_unregister_float_streaming_buffers.argtypes = [P]
    lib.octb200_set_callbacks.argtypes = [P, HOST_CALLBACK, HOST_CALLBACK, HOST_CALLBACK]
    lib.octb200_process_host.argtypes = [P, C.c_void_p]
    lib.octb200_process_device.argtypes = [P, C.c_void_p]
    lib.octb200_sync.argtypes = [P]
    lib.octb200_current_buffer_nr.argtypes = [P]; lib.octb200_current_buffer_nr.restype = C.c_uint32
    lib.octb200_output_device_ptr.argtypes = [P, C.c_uint32]; lib.octb200_output_device_ptr.restype = C.c_void_p
    lib.octb200_copy_output.argtypes = [P, C.c_void_p, C.c_uint32]
    lib.octb200_bind_output.argtypes = [P, C.c_void_p]
    lib.octb200_bscan_frame.argtypes = [P, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p]
    lib.octb200_enface_frame.argtypes = [P, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p]
    lib.octb200_volume_u8.argtypes = [P, C.c_uint32, C.c_void_p]
    lib.octb200_float_to_output.argtypes = [P, C.c_uint32, C.c_void_p]
    lib.octb200_compute_stream.argtypes = [P]; lib.octb200_compute_stream.restype = C.c_void_p
    lib.octb200_event_record.argtypes = [P, C.c_int]
    lib.octb200_event_elapsed_ms.argtypes = [P, C.c_int, C.c_int, fp]
    lib.octb200_launch_count.argtypes = [P]; lib.octb200_launch_count.restype = C.c_uint64
    lib.octb200_time_kernel.argtypes = [P, C.c_void_p, C.c_int, fp]
    lib.octb200_enface_gather_init.argtypes = [P, C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_void_p]
    lib.octb200_enface_gather_connect.argtypes = [P, C.c_void_p]
    lib.octb200_enface_gather.argtypes = [P, C.c_uint32, C.c_uint32, C.c_int]
    lib.octb200_enface_gather_auto.argtypes = [P, C.c_int, C.c_uint32, C.c_uint32, C.c_int]
    lib.octb200_enface_gather_wait.argtypes = [P, C.POINTER(C.c_void_p)]
    lib.octb200_enface_gather_status.argtypes = [P, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    lib.octb200_enface_gather_close.argtypes = [P]
    lib.octb200_dispersion_sweep.argtypes = [P, C.c_void_p, C.POINTER(SweepConfig), C.c_void_p, C.c_void_p, C.c_void_p]
    _lib = lib
    return lib
