"""Headless mirror of the reference's dispersion estimator (octproz-dispersion-estimator-extension), on the GPU.

The reference's `DispersionEstimationEngine::startDispersionEstimation` (src/dispersionestimationengine.cpp:21-116) sweeps
d2 (with d3 = 0), then d3 at the best d2, and for EVERY trial value re-runs its CPU processing path on the center A-scans
(`processDispersionMetric`, :118-158) and scores the result with `AscanMetricCalculator` (src/ascanmetriccalculator.cpp).
Here each sweep is ONE call of `octb200_dispersion_sweep`: all trial values are processed by a single launch of the fused
sm_100a kernel (one table image per trial) plus one metric kernel -- two launches per coefficient instead of
numberOfDispersionSamples CPU passes.  Names, parameters and search semantics follow the reference; this module owns no
arithmetic beyond the loop that walks the trial values (the engine's own job).
"""
from __future__ import annotations

import copy
from dataclasses import dataclass

import numpy as np

from . import _lib

CPU_PATH_ROLLING_WINDOW = 10      # processor.h:26 default; see startDispersionEstimation

# ASCAN_SHARPNESS_METRIC (src/dispersionestimatorparameters.h:51-56)
SUM_ABOVE_THRESHOLD, SAMPLES_ABOVE_THRESHOLD, PEAK_VALUE, MEAN_SOBEL = 0, 1, 2, 3


@dataclass
class DispersionEstimatorParameters:
    """src/dispersionestimatorparameters.h:58-76 (GUI-only fields omitted)"""
    numberOfCenterAscans: int = 10
    useLinearAscans: bool = True
    numberOfAscanSamplesToIgnore: int = 0
    autoCalcD1: bool = False
    sharpnessMetric: int = SUM_ABOVE_THRESHOLD
    metricThreshold: float = 0.0
    d2start: float = -100.0
    d2end: float = 100.0
    d3start: float = -100.0
    d3end: float = 100.0
    numberOfDispersionSamples: int = 100


def cpu_path_window(n: int) -> np.ndarray:
    """the window the estimator's processing path applies whatever the main window setting is: Hanning over n-1 points,
    evaluated in float (octprocessor/processor.tpp:124-133)"""
    factor = np.float32(2.0 * np.pi / (n - 1))
    i = np.arange(n, dtype=np.float32)
    return (np.float32(0.5) * (np.float32(1) - np.cos(factor * i, dtype=np.float32))).astype(np.float32)


class DispersionEstimationEngine:
    def __init__(self, pipeline):
        """pipeline: an initialised OctPipeline whose parameter object carries the processing settings the reference reads from
        its settings file (processorcontroller.cpp:38-92): resampling + coefficients, windowing, background removal, d0/d1,
        log min/max/coeff/addend."""
        self.pipe = pipeline
        self.params = DispersionEstimatorParameters()
        self.bestD2 = self.bestD3 = 0.0
        self.bestMetricValueD2 = self.bestMetricValueD3 = 0.0
        self.calculatedD1 = 0.0
        self.metricsD2: list[tuple[float, float]] = []
        self.metricsD3: list[tuple[float, float]] = []
        self.ascanWithoutDispersionCompensation = None
        self.ascanWithBestDispersion = None

    def setParams(self, params: DispersionEstimatorParameters) -> None:
        self.params = params

    # ------------------------------------------------------------------
    def _sweep(self, raw, pairs, want_ascans=False):
        q = self.pipe.params
        co = np.array([[q.d0, q.d1, d2, d3] for d2, d3 in pairs], np.float32)
        prm = self.params
        return self.pipe.dispersion_sweep(raw, co, prm.sharpnessMetric, prm.metricThreshold, prm.numberOfAscanSamplesToIgnore,
                                          log_scale=not prm.useLinearAscans, log_min=q.signalGrayscaleMin, log_max=q.signalGrayscaleMax,
                                          log_coeff=q.signalMultiplicator, log_addend=q.signalAddend, want_ascans=want_ascans)

    def startDispersionEstimation(self, frameBuffer: np.ndarray, bitDepth: int, samplesPerLine: int, linesPerFrame: int) -> dict:
        """src/dispersionestimationengine.cpp:21-116.  frameBuffer: one raw frame [linesPerFrame][samplesPerLine]."""
        prm = self.params
        frame = np.ascontiguousarray(frameBuffer).reshape(linesPerFrame, samplesPerLine)
        center = min(int(prm.numberOfCenterAscans), int(linesPerFrame))                    # :37
        offset = (linesPerFrame - center) // 2 if center < linesPerFrame else 0            # :42-47
        raw = np.ascontiguousarray(frame[offset:offset + center])

        # the estimator's path has its own window (processor.tpp:124-133) and always compensates dispersion with the trial values;
        # everything else comes from the main settings.  The handle's curves are restored afterwards.
        q = self.pipe.params
        saved = copy.copy(q)
        saved_window = None if q.windowCurve is None else q.windowCurve.copy()
        try:
            q.bitshift = False
            # quirk reproduced: the reference's controller hands the window-size setting to the constructor's `windowSize` slot
            # (processorcontroller.cpp:116 vs processor.h:26), so its DC removal always runs with the default window of 10
            q.rollingAverageWindowSize = CPU_PATH_ROLLING_WINDOW
            q.windowCurve = cpu_path_window(samplesPerLine); q.windowUpdated = True
            self.pipe.push_params()
            n = int(prm.numberOfDispersionSamples)
            stepD2 = abs(prm.d2end - prm.d2start) / float(n)                               # :69
            stepD3 = abs(prm.d3end - prm.d3start) / float(n)                               # :70
            # ---- d2, with d3 = 0 (:78-82) ----
            d2s, d = [], float(prm.d2start)
            for _ in range(n):
                d2s.append(d); d += stepD2
            m2 = self._sweep(raw, [(v, 0.0) for v in d2s])
            self.bestD2, self.bestMetricValueD2 = 0.0, 0.0
            for v, m in zip(d2s, m2):
                if self.bestMetricValueD2 < float(m):                                      # :140-143, strict
                    self.bestMetricValueD2, self.bestD2 = float(m), v
            self.metricsD2 = list(zip(d2s, (float(x) for x in m2)))
            # ---- d3, at the best d2 (:85-90) ----
            d3s, d = [], float(prm.d3start)
            for _ in range(n):
                d3s.append(d); d += stepD3
            m3 = self._sweep(raw, [(self.bestD2, v) for v in d3s])
            self.bestD3, self.bestMetricValueD3 = 0.0, 0.0
            for v, m in zip(d3s, m3):
                if self.bestMetricValueD3 < float(m):                                      # :147-150
                    self.bestMetricValueD3, self.bestD3 = float(m), v
            self.metricsD3 = list(zip(d3s, (float(x) for x in m3)))
            # ---- the two A-scans the GUI plots (:93-96, processFirstLineOnly :160-192; its offset is applied a second time
            #      inside the already extracted center block, reproduced here) ----
            off2 = (center - min(int(prm.numberOfCenterAscans), center)) // 2
            first = raw[off2:off2 + 1]
            _, a = self._sweep(first, [(0.0, 0.0), (self.bestD2, self.bestD3)], want_ascans=True)
            self.ascanWithoutDispersionCompensation, self.ascanWithBestDispersion = a[0, 0].copy(), a[1, 0].copy()
            if prm.autoCalcD1:
                self.calculatedD1 = -(self.bestD2 + self.bestD3)                           # :101
        finally:
            q.bitshift = saved.bitshift
            q.rollingAverageWindowSize = saved.rollingAverageWindowSize
            q.windowCurve = saved_window; q.windowUpdated = True
            self.pipe.push_params()
        return {"bestD2": self.bestD2, "bestD3": self.bestD3, "calculatedD1": self.calculatedD1 if prm.autoCalcD1 else None,
                "metricsD2": self.metricsD2, "metricsD3": self.metricsD3}
