"""ctypes wrappers around the checker libraries (TEST INFRASTRUCTURE ONLY).

  liboct_oracle.so        our C restatement (oracle/oct_oracle.c)                    -- always available
  _ref/libref_luts.so     the reference's own host LUT code                          -- container only (needs /root/reference to build)
  _ref/libref_cpu.so      the reference's own CPU path + FFTW-API substitute         -- prebuilt, travels to the GPU box
  _ref/libref_cuda.so     the reference's unmodified cuda_code.cu for sm_100         -- prebuilt, runs only on a GPU

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


def build(ref: bool = False) -> None:
    subprocess.check_call(["make", "-s", "-C", HERE, "all"])
    if ref and os.path.isdir("/root/reference"):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


class OrcParams(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("samplesPerLine", "ascansPerBscan", "bscansPerBuffer", "bitDepth", "bitshift",
                                       "backgroundRemoval", "rollingAverageWindowSize", "resampling", "interpolation",
                                       "windowing", "dispersionCompensation", "fixedPatternNoiseRemoval",
                                       "bscansForNoiseDetermination", "signalLogScaling")] + \
               [(n, C.c_float) for n in ("signalGrayscaleMin", "signalGrayscaleMax", "signalMultiplicator", "signalAddend")] + \
               [("bscanFlip", C.c_int), ("sinusoidalScanCorrection", C.c_int), ("postProcessBackgroundRemoval", C.c_int),
                ("postProcessBackgroundWeight", C.c_float), ("postProcessBackgroundOffset", C.c_float)]


_orc = None


def lib() -> C.CDLL:
    global _orc
    if _orc is None:
        path = os.path.join(HERE, "liboct_oracle.so")
        if not os.path.exists(path):
            build()
        _orc = C.CDLL(path)
        _orc.orc_resample_curve.argtypes = [C.c_int] + [C.c_float] * 4 + [C.c_void_p]
        _orc.orc_dispersion_curve.argtypes = [C.c_int] + [C.c_float] * 4 + [C.c_void_p]
        _orc.orc_window_curve.argtypes = [C.c_int, C.c_float, C.c_float, C.c_int, C.c_void_p]
        _orc.orc_sinusoidal_curve.argtypes = [C.c_int, C.c_void_p]
        _orc.orc_process.argtypes = [C.POINTER(OrcParams)] + [C.c_void_p] * 6 + [C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        _orc.orc_postprocess_background.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        _orc.orc_bscan_frame.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_uint, C.c_uint, C.c_int, C.c_void_p]
        _orc.orc_enface_frame.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_uint, C.c_uint, C.c_int, C.c_void_p]
        _orc.orc_float_to_output.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
    return _orc


def resample_curve(n, c0, c1, c2, c3):
    out = np.empty(n, np.float32); lib().orc_resample_curve(n, c0, c1, c2, c3, out.ctypes.data); return out


def dispersion_curve(n, d0, d1, d2, d3):
    out = np.empty(n, np.float32); lib().orc_dispersion_curve(n, d0, d1, d2, d3, out.ctypes.data); return out


def window_curve(wtype, center, fill, n):
    out = np.empty(n, np.float32); lib().orc_window_curve(wtype, center, fill, n, out.ctypes.data); return out


def sinusoidal_curve(a):
    out = np.empty(a, np.float32); lib().orc_sinusoidal_curve(a, out.ctypes.data); return out


def params_from(q) -> OrcParams:
    """q: octproz_b200.params.OctAlgorithmParameters (duck typed)"""
    p = OrcParams()
    p.samplesPerLine, p.ascansPerBscan, p.bscansPerBuffer, p.bitDepth = q.samplesPerLine, q.ascansPerBscan, q.bscansPerBuffer, q.bitDepth
    p.bitshift = int(q.bitshift); p.backgroundRemoval = int(q.backgroundRemoval)
    p.rollingAverageWindowSize = max(1, int(q.rollingAverageWindowSize))
    p.resampling = int(q.resampling); p.interpolation = int(q.resamplingInterpolation)
    p.windowing = int(q.windowing); p.dispersionCompensation = int(q.dispersionCompensation)
    p.fixedPatternNoiseRemoval = int(q.fixedPatternNoiseRemoval); p.bscansForNoiseDetermination = int(q.bscansForNoiseDetermination)
    p.signalLogScaling = int(q.signalLogScaling)
    p.signalGrayscaleMin, p.signalGrayscaleMax = q.signalGrayscaleMin, q.signalGrayscaleMax
    p.signalMultiplicator, p.signalAddend = q.signalMultiplicator, q.signalAddend
    p.bscanFlip = int(q.bscanFlip); p.sinusoidalScanCorrection = int(q.sinusoidalScanCorrection)
    p.postProcessBackgroundRemoval = int(q.postProcessBackgroundRemoval)
    p.postProcessBackgroundWeight, p.postProcessBackgroundOffset = q.postProcessBackgroundWeight, q.postProcessBackgroundOffset
    return p


def process(q, raw: np.ndarray, mean_line: np.ndarray | None = None, determine_fpn: bool = True, precision: int = 64,
            want_complex: bool = False, pp_background: np.ndarray | None = None):
    """run the oracle chain on one raw buffer.  Curves are taken from q (q.resampleCurve ...).
    returns (out [B][A][N/2] float32, mean_line [N][2] float64, complex or None)"""
    n, a, b = int(q.samplesPerLine), int(q.ascansPerBscan), int(q.bscansPerBuffer)
    raw = np.ascontiguousarray(raw)
    assert raw.size == n * a * b
    p = params_from(q)
    out = np.empty((b, a, n // 2), np.float32)
    ml = np.zeros((n, 2), np.float64) if mean_line is None else np.ascontiguousarray(mean_line, np.float64).copy()
    cplx = np.empty((b, a, n, 2), np.float64) if want_complex else None
    f32 = lambda x: None if x is None else np.ascontiguousarray(x, np.float32)
    rs, ds, ws, bg = f32(q.resampleCurve), f32(q.dispersionCurve), f32(q.windowCurve), f32(pp_background)
    rc = lib().orc_process(C.byref(p), raw.ctypes.data,
                           rs.ctypes.data if rs is not None else None, ds.ctypes.data if ds is not None else None,
                           ws.ctypes.data if ws is not None else None, bg.ctypes.data if bg is not None else None,
                           ml.ctypes.data, int(determine_fpn), out.ctypes.data,
                           cplx.ctypes.data if cplx is not None else None, precision)
    if rc != 0:
        raise RuntimeError(f"orc_process failed: {rc}")
    return out, ml, cplx


def postprocess_background(processed: np.ndarray, half_n: int, a: int) -> np.ndarray:
    out = np.empty(half_n, np.float32)
    lib().orc_postprocess_background(np.ascontiguousarray(processed, np.float32).ctypes.data, half_n, a, out.ctypes.data)
    return out


def bscan_frame(vol, half_n, a, btot, frame, nframes, fn):
    out = np.zeros(half_n * a, np.float32)
    lib().orc_bscan_frame(np.ascontiguousarray(vol, np.float32).ctypes.data, half_n, a, btot, frame, nframes, fn, out.ctypes.data)
    return out


def enface_frame(vol, half_n, a, btot, frame, nframes, fn):
    out = np.zeros(a * btot, np.float32)
    lib().orc_enface_frame(np.ascontiguousarray(vol, np.float32).ctypes.data, half_n, a, btot, frame, nframes, fn, out.ctypes.data)
    return out


def float_to_output(x: np.ndarray, bit_depth: int) -> np.ndarray:
    dt = np.uint8 if bit_depth <= 8 else (np.uint16 if bit_depth <= 16 else np.uint32)
    x = np.ascontiguousarray(x, np.float32)
    out = np.empty(x.shape, dt)
    lib().orc_float_to_output(x.ctypes.data, x.size, bit_depth, out.ctypes.data)
    return out


# ----------------------------------------------------------------------------- reference-built libraries
def have_ref(name: str) -> bool:
    return os.path.exists(os.path.join(REF_DIR, name))


def ref_luts(n, c, d, wtype, center, fill):
    """LUTs from the reference's own octalgorithmparameters.cpp / polynomial.cpp / windowfunction.cpp"""
    L = C.CDLL(os.path.join(REF_DIR, "libref_luts.so"))
    L.ref_luts.argtypes = [C.c_int] + [C.c_float] * 8 + [C.c_int, C.c_float, C.c_float] + [C.c_void_p] * 3
    r, dd, w = (np.empty(n, np.float32) for _ in range(3))
    L.ref_luts(n, *c, *d, wtype, center, fill, r.ctypes.data, dd.ctypes.data, w.ctypes.data)
    return r, dd, w


class RefCpu:
    """the reference's CPU path (processor.tpp) + FFTW-API substitute: TIMED BASELINE, not a parity oracle"""

    def __init__(self):
        self.L = C.CDLL(os.path.join(REF_DIR, "libref_cpu.so"))
        self.L.refcpu_process.argtypes = [C.c_void_p] + [C.c_int] * 10 + [C.c_void_p, C.c_void_p] + [C.c_float] * 4 + [C.c_int, C.c_void_p]
        self.max_threads = int(self.L.refcpu_max_threads())

    def process(self, q, raw: np.ndarray, threads: int = 1) -> np.ndarray:
        n, a = int(q.samplesPerLine), int(q.ascansPerBscan)
        raw = np.ascontiguousarray(raw)
        b = raw.size // (n * a)
        out = np.empty((b, a, n // 2), np.float32)
        c = np.array([q.c0, q.c1, q.c2, q.c3], np.float32); d = np.array([q.d0, q.d1, q.d2, q.d3], np.float32)
        self.L.refcpu_process(raw.ctypes.data, int(q.bitDepth), n, a, b, int(q.rollingAverageWindowSize), int(q.backgroundRemoval),
                              int(q.resampling), int(q.dispersionCompensation), int(q.windowing), int(q.signalLogScaling),
                              c.ctypes.data, d.ctypes.data, q.signalMultiplicator, q.signalGrayscaleMin, q.signalGrayscaleMax,
                              q.signalAddend, threads, out.ctypes.data)
        return out


class RefCudaCfg(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("samplesPerLine", "ascansPerBscan", "bscansPerBuffer", "buffersPerVolume", "bitDepth",
                                       "bitshift", "bscanFlip", "signalLogScaling", "sinusoidalScanCorrection")] + \
               [(n, C.c_float) for n in ("signalGrayscaleMin", "signalGrayscaleMax", "signalMultiplicator", "signalAddend")] + \
               [("backgroundRemoval", C.c_int), ("rollingAverageWindowSize", C.c_int), ("resampling", C.c_int), ("resamplingInterpolation", C.c_int)] + \
               [(n, C.c_float) for n in ("c0", "c1", "c2", "c3")] + [("dispersionCompensation", C.c_int)] + \
               [(n, C.c_float) for n in ("d0", "d1", "d2", "d3")] + [("windowing", C.c_int), ("windowType", C.c_int)] + \
               [("windowCenter", C.c_float), ("windowFillFactor", C.c_float)] + \
               [("fixedPatternNoiseRemoval", C.c_int), ("continuousFixedPatternNoiseDetermination", C.c_int), ("bscansForNoiseDetermination", C.c_int),
                ("postProcessBackgroundRemoval", C.c_int), ("postProcessBackgroundWeight", C.c_float), ("postProcessBackgroundOffset", C.c_float),
                ("streamToHost", C.c_int), ("saveAs32bitFloat", C.c_int)]


class RefCuda:
    """the reference's unmodified cuda_code.cu (sm_100, --use_fast_math) behind a headless driver.  GPU only."""

    def __init__(self, lib: str = "libref_cuda.so"):
        """lib = "libadapter_api.so": the same harness and the reference's own parameter code, but the kernels.h symbols
        come from integration/octproz_kernels_adapter.cpp on top of liboctb200.so (drop-in check)"""
        self.L = C.CDLL(os.path.join(REF_DIR, lib))
        self.L.refcuda_configure.argtypes = [C.POINTER(RefCudaCfg)]
        self.L.refcuda_get_curves.argtypes = [C.c_void_p] * 3
        self.L.refcuda_init.argtypes = [C.c_void_p, C.c_void_p]
        self.L.refcuda_process.argtypes = [C.c_void_p]
        self.L.refcuda_copy_output.argtypes = [C.c_void_p, C.c_int]
        self.L.refcuda_time.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]; self.L.refcuda_time.restype = C.c_double
        self.L.refcuda_set_postprocess_background.argtypes = [C.c_void_p, C.c_int]
        self.L.refcuda_register_streaming.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        self.L.refcuda_callback_counts.argtypes = [C.POINTER(C.c_int)] * 3
        self.q = None
        self._bufs = None

    def configure(self, q) -> None:
        c = RefCudaCfg()
        for f, _ in RefCudaCfg._fields_:
            src = {"windowType": "window", "saveAs32bitFloat": "saveAs32bitFloat"}.get(f, f)
            v = getattr(q, src)
            setattr(c, f, float(v) if isinstance(getattr(c, f), float) else int(v))
        self.L.refcuda_configure(C.byref(c))
        self.q = q

    def curves(self):
        n = int(self.q.samplesPerLine)
        r, d, w = (np.zeros(n, np.float32) for _ in range(3))
        self.L.refcuda_get_curves(r.ctypes.data, d.ctypes.data, w.ctypes.data)
        return r, d, w

    def init(self, h1: np.ndarray, h2: np.ndarray) -> None:
        self._bufs = (h1, h2)
        if self.L.refcuda_init(h1.ctypes.data, h2.ctypes.data) != 0:
            raise RuntimeError("reference initializeCuda failed")

    def process(self, h_in: np.ndarray | None) -> None:
        self.L.refcuda_process(h_in.ctypes.data if h_in is not None else None)

    def output(self, buffer_nr: int = 0) -> np.ndarray:
        q = self.q
        out = np.empty((q.bscansPerBuffer, q.ascansPerBscan, q.samplesPerLine // 2), np.float32)
        rc = self.L.refcuda_copy_output(out.ctypes.data, buffer_nr)
        if rc != 0:
            raise RuntimeError(f"reference copy_output: cuda error {rc}")
        return out

    def mean_line(self) -> np.ndarray:
        n = int(self.q.samplesPerLine)
        out = np.empty((n, 2), np.float32)
        self.L.refcuda_get_mean_line.argtypes = [C.c_void_p, C.c_int]
        self.L.refcuda_get_mean_line(out.ctypes.data, n)
        return out

    def time(self, h_a, h_b, iters: int, warmup: int) -> float:
        return float(self.L.refcuda_time(h_a.ctypes.data if h_a is not None else None,
                                         h_b.ctypes.data if h_b is not None else None, iters, warmup))

    def cleanup(self) -> None:
        self.L.refcuda_cleanup()
        self._bufs = None
