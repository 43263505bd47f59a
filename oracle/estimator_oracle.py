"""Oracle of the dispersion-estimator path (TEST INFRASTRUCTURE ONLY -- never imported by the product).

numpy restatement, in the reference's operation order, of
  * OCTSignalProcessing::Processor<float>::processRawData  (octproz-dispersion-estimator-extension/src/octprocessor/processor.tpp:241-321
    with :124-133 window, :135-174 dispersive phase, :176-196 resample curve, :323-345 rolling DC removal, :347-380 cubic
    k-linearisation, :382-407 dispersion, :409-414 window, :416-434 IFFT / N, :436-470 log scale),
  * AscanMetricCalculator::calculateMetric                 (src/ascanmetriccalculator.cpp:22-128),
  * DispersionEstimationEngine::startDispersionEstimation  (src/dispersionestimationengine.cpp:21-116, the search loop).
Pinned against the reference's own code compiled in place (oracle/_ref/libref_cpu.so, libref_metric.so: tests/test_estimator.py,
container only) and against tests/golden/estimator.npz produced from those libraries (tests/golden/make_golden_estimator.py).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
f32 = np.float32

SUM_ABOVE_THRESHOLD, SAMPLES_ABOVE_THRESHOLD, PEAK_VALUE, MEAN_SOBEL = 0, 1, 2, 3
# ProcessorController::processData builds Processor(samplesPerSpectrum, settings.rollingAverageWindowSize) (processorcontroller.cpp:116),
# but the constructor's second parameter is `windowSize` (processor.h:26): the rolling window is ALWAYS the default, 10
CPU_PATH_ROLLING_WINDOW = 10


def cpu_window(n: int) -> np.ndarray:
    """processor.tpp:124-133: Hanning over n-1, evaluated in float"""
    factor = f32(2.0 * np.pi / (n - 1))
    i = np.arange(n, dtype=f32)
    return (f32(0.5) * (f32(1) - np.cos(factor * i, dtype=f32))).astype(f32)


def cpu_resample_curve(n: int, c) -> np.ndarray:
    """processor.tpp:176-196"""
    c0 = f32(c[0]); c1 = f32(c[1]) / (f32(n) - f32(1)); c2 = f32(c[2]) / (f32(n - 1.0) * f32(n - 1.0))
    c3 = f32(c[3]) / (f32(n - 1.0) * f32(n - 1.0) * f32(n - 1.0))
    x = np.arange(n, dtype=f32)
    val = (c0 + x * (c1 + x * (c2 + x * c3))).astype(f32)
    return np.clip(val, f32(0), f32(n - 3)).astype(f32)


def cpu_phase(n: int, d) -> np.ndarray:
    """processor.tpp:135-174 -> complex64 phasors"""
    denom = f32(n - 1)
    k = [f32(d[0]), f32(d[1]) / denom, f32(d[2]) / (denom * denom), f32(d[3]) / (denom * denom * denom)]
    i = np.arange(n, dtype=f32)
    ph = (k[0] + i * (k[1] + i * (k[2] + i * k[3]))).astype(f32)
    return (np.cos(ph, dtype=f32) + 1j * np.sin(ph, dtype=f32)).astype(np.complex64)


def cpu_process(raw: np.ndarray, n: int, *, remove_dc=False, rolling_window=CPU_PATH_ROLLING_WINDOW, resample=True, c=(0, 0, 0, 0),
                dispersion=True, d=(0, 0, 0, 0), window=True, log_scale=True, coeff=1.0, vmin=0.0, vmax=100.0, addend=0.0) -> np.ndarray:
    """raw: [lines][n] unsigned containers -> [lines][n/2] float32 in the CPU path's units"""
    x = np.ascontiguousarray(raw).reshape(-1, n).astype(f32)
    lines = x.shape[0]
    if remove_dc:                                             # processor.tpp:323-345
        w = int(rolling_window)
        cs = np.concatenate([np.zeros((lines, 1), f32), np.cumsum(x, axis=1, dtype=f32)], axis=1)
        idx = np.arange(n)
        s = np.where(idx >= w - 1, idx - (w - 1), 0); e = np.minimum(idx + w, n - 1)
        mean = (cs[:, e + 1] - cs[:, s]) / (e - s + 1).astype(f32)
        x = (x - mean).astype(f32)
    if resample:                                              # processor.tpp:347-380 + :472-481
        curve = cpu_resample_curve(n, c)
        n1 = curve.astype(np.int64)
        n0 = np.minimum(np.abs(n1 - 1), n - 1); n2 = np.minimum(n1 + 1, n - 1); n3 = np.minimum(n1 + 2, n - 1); n1c = np.minimum(n1, n - 1)
        y0, y1, y2, y3 = x[:, n0], x[:, n1c], x[:, n2], x[:, n3]
        pos = (curve - n1.astype(f32)).astype(f32)
        a_ = (-y0 + f32(3.0) * (y1 - y2) + y3); b_ = (f32(2.0) * y0 - f32(5.0) * y1 + f32(4.0) * y2 - y3); c_ = (-y0 + y2)
        pos2 = pos * pos
        x = (f32(0.5) * pos * (a_ * pos2 + b_ * pos + c_) + y1).astype(f32)
    z = x.astype(np.complex64)
    if dispersion:
        z = (z.real[..., None] * np.stack([cpu_phase(n, d).real, cpu_phase(n, d).imag], -1)).astype(f32)
        z = (z[..., 0] + 1j * z[..., 1]).astype(np.complex64)
    if window:
        z = (z * cpu_window(n)).astype(np.complex64)
    spec = (np.fft.ifft(z.astype(np.complex128), axis=1) * n).astype(np.complex128)     # fftw backward, unnormalised ...
    spec = (spec.astype(np.complex64) * f32(1.0 / n)).astype(np.complex64)                # ... then * 1/N in float
    if log_scale:                                             # processor.tpp:436-470
        mag2 = (spec.real * spec.real + spec.imag * spec.imag).astype(f32)
        with np.errstate(divide="ignore"):
            val = (f32(10.0) * np.log10(mag2 / f32(n), dtype=f32)).astype(f32)
        rng = f32(vmax) - f32(vmin)
        out = (f32(coeff) * ((val - f32(vmin)) / rng + f32(addend))).astype(f32)
    else:
        out = np.abs(spec).astype(f32)
    return out[:, : n // 2].copy()


def ascan_metric(data: np.ndarray, samples_per_line: int, metric: int, threshold: float, ignore: int) -> np.float32:
    """ascanmetriccalculator.cpp:22-128 with float accumulators and the reference's summation order"""
    data = np.ascontiguousarray(data, f32).reshape(-1)
    if samples_per_line <= 0 or data.size == 0:
        return f32(0)
    lines = data.size // samples_per_line
    ig = min(ignore, samples_per_line) if ignore > 0 else ignore
    valid = samples_per_line - ig
    total = f32(0)
    thr = f32(threshold)
    for l in range(lines):
        if valid <= 0:
            continue
        d = data[l * samples_per_line + ig: l * samples_per_line + ig + valid]
        if metric == SUM_ABOVE_THRESHOLD:
            sel = d[d > thr]
            m = np.cumsum(sel, dtype=f32)[-1] if sel.size else f32(0)        # sequential float sum
        elif metric == SAMPLES_ABOVE_THRESHOLD:
            m = f32(int((d > thr).sum()))
        elif metric == PEAK_VALUE:
            m = f32(max(0.0, float(d.max())))
        elif metric == MEAN_SOBEL:
            if valid < 3:
                m = f32(0)
            else:
                g = np.abs((d[2:] - d[:-2]) * f32(0.5)).astype(f32)
                m = np.cumsum(g, dtype=f32)[-1] / f32(g.size)
        else:
            m = f32(0)
        total = f32(total + f32(m))
    return f32(total)


def estimate(process_trials, params: dict) -> dict:
    """dispersionestimationengine.cpp:63-116: d2 sweep with d3 = 0, then d3 sweep at the best d2; strict '<' from 0.
    process_trials(list of (d2, d3) as float32 pairs) -> sequence of metric values."""
    n = int(params["numberOfDispersionSamples"])
    step2 = abs(params["d2end"] - params["d2start"]) / float(n)
    step3 = abs(params["d3end"] - params["d3start"]) / float(n)
    d2s, d = [], float(params["d2start"])
    for _ in range(n):
        d2s.append(d); d += step2
    m2 = [float(v) for v in process_trials([(f32(v), f32(0.0)) for v in d2s])]
    best2, bm2 = 0.0, 0.0
    for v, m in zip(d2s, m2):
        if bm2 < m:
            bm2, best2 = m, v
    d3s, d = [], float(params["d3start"])
    for _ in range(n):
        d3s.append(d); d += step3
    m3 = [float(v) for v in process_trials([(f32(best2), f32(v)) for v in d3s])]
    best3, bm3 = 0.0, 0.0
    for v, m in zip(d3s, m3):
        if bm3 < m:
            bm3, best3 = m, v
    return {"bestD2": best2, "bestD3": best3, "d2": d2s, "metricD2": m2, "d3": d3s, "metricD3": m3,
            "calculatedD1": -(best2 + best3) if params.get("autoCalcD1") else None}


def center_lines(lines_per_frame: int, number_of_center_ascans: int):
    """dispersionestimationengine.cpp:37-47 -> (offset, count)"""
    c = min(int(number_of_center_ascans), int(lines_per_frame))
    off = (lines_per_frame - c) // 2 if c < lines_per_frame else 0
    return off, c


# ----------------------------------------------------------------------------- the reference's own code, compiled in place
def have_ref_metric() -> bool:
    return os.path.exists(os.path.join(REF_DIR, "libref_metric.so"))


def ref_metric(data: np.ndarray, samples_per_line: int, metric: int, threshold: float, ignore: int) -> float:
    L = C.CDLL(os.path.join(REF_DIR, "libref_metric.so"))
    L.refmetric_calculate.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int]; L.refmetric_calculate.restype = C.c_float
    d = np.ascontiguousarray(data, f32).reshape(-1)
    return float(L.refmetric_calculate(d.ctypes.data, d.size, samples_per_line, metric, float(threshold), ignore))
