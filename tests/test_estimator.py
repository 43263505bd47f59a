"""Dispersion-estimator path on CPU: the oracle (oracle/estimator_oracle.py) against golden vectors produced by the reference's
own code (tests/golden/estimator.npz, tests/golden/make_golden_estimator.py), against that code itself when oracle/_ref is
present, and the host-side engine mirror (octproz_b200/dispersion_estimator.py) with the GPU call replaced by the oracle."""
import os

import numpy as np
import pytest

from octproz_b200 import benchmark_params
from octproz_b200.dispersion_estimator import (DispersionEstimationEngine, DispersionEstimatorParameters, PEAK_VALUE,
                                              cpu_path_window)
from oracle import estimator_oracle as eo

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "estimator.npz"))
N = 1024
EST = dict(numberOfDispersionSamples=16, d2start=-160.0, d2end=0.0, d3start=-40.0, d3end=40.0, autoCalcD1=True)


def oracle_ascans(d2, d3, log, raw=None):
    lg = G["log"]
    return eo.cpu_process(G["raw"] if raw is None else raw, N, c=G["c"], d=(G["d01"][0], G["d01"][1], d2, d3), log_scale=bool(log),
                          vmin=lg[0], vmax=lg[1], coeff=lg[2], addend=lg[3])


@pytest.mark.parametrize("log", [0, 1])
def test_oracle_cpu_path_matches_reference_golden(log):
    for k, (d2, d3) in enumerate(G["trials"]):
        ref = G[f"ascans_log{log}"][k]
        got = oracle_ascans(d2, d3, log)
        if log:
            # log output is unbounded at spectral nulls: compare where the reference is above its 5 % quantile
            # (5e-4 of the 130 dB display range = 0.065 dB, the fp32 floor of the small bins next to the DC peak)
            keep = ref > np.quantile(ref, 0.05)
            assert np.abs(got - ref)[keep].max() < 5e-4
        else:
            # fp32 round-off scales with the largest bin (the DC peak), not with the local value
            assert np.all(np.abs(got - ref) <= 1e-4 * np.abs(ref) + 1e-5 * np.abs(ref).max())


@pytest.mark.parametrize("log", [0, 1])
def test_oracle_metric_is_the_reference_metric(log):
    """same processed data in -> bit-identical metric value out, all four metrics, with and without ignored samples"""
    a = G[f"ascans_log{log}"]
    thr = float(G[f"thr_log{log}"])
    for m in range(4):
        for j, ig in enumerate((0, 15)):
            for k in range(a.shape[0]):
                assert np.float32(eo.ascan_metric(a[k], N // 2, m, thr, ig)) == G[f"metrics_log{log}"][m, j, k], (m, ig, k)
    # degenerate inputs (ascanmetriccalculator.cpp:24-26, 44-47)
    assert eo.ascan_metric(np.zeros(0, np.float32), 512, 0, 0.0, 0) == 0
    assert eo.ascan_metric(a[0], N // 2, 0, thr, N) == 0          # everything ignored


@pytest.mark.parametrize("log", [0, 1])
def test_search_restatement_reproduces_reference_search(log):
    thr = float(G[f"thr_log{log}"])
    res = eo.estimate(lambda pairs: [eo.ascan_metric(oracle_ascans(d2, d3, log), N // 2, eo.PEAK_VALUE, thr, 15) for d2, d3 in pairs], EST)
    assert (res["bestD2"], res["bestD3"], res["calculatedD1"]) == tuple(G[f"search_log{log}"])
    assert np.allclose(res["metricD2"], G[f"search_metricD2_log{log}"], rtol=2e-4, atol=1e-3)
    assert np.allclose(res["metricD3"], G[f"search_metricD3_log{log}"], rtol=2e-4, atol=1e-3)


def test_center_lines_and_window():
    assert eo.center_lines(512, 10) == (251, 10) and eo.center_lines(8, 10) == (0, 8)
    assert np.array_equal(cpu_path_window(N), eo.cpu_window(N))


@pytest.mark.skipif(not (eo.have_ref_metric() and os.path.exists(os.path.join(eo.REF_DIR, "libref_cpu.so"))), reason="oracle/_ref not built")
def test_oracle_against_live_reference_code():
    import copy
    from oracle import oracle as orc
    rng = np.random.default_rng(5)
    q = benchmark_params(N, 4, 1); q.update_all_curves()
    raw = rng.integers(0, 4096, (4, N)).astype(np.uint16)
    rc = orc.RefCpu()
    for log in (0, 1):
        for kw in (dict(), dict(backgroundRemoval=True, rollingAverageWindowSize=17), dict(resampling=False), dict(windowing=False)):
            qq = copy.copy(q); qq.signalLogScaling = bool(log)
            for k, v in kw.items():
                setattr(qq, k, v)
            ref = rc.process(qq, raw, threads=1).reshape(4, N // 2)
            # quirk reproduced: ProcessorController passes the window-size setting into the constructor's `windowSize` slot
            # (processorcontroller.cpp:116 vs processor.h:26), so the rolling window is always the default 10
            got = eo.cpu_process(raw, N, remove_dc=qq.backgroundRemoval, rolling_window=eo.CPU_PATH_ROLLING_WINDOW, resample=qq.resampling,
                                 c=(q.c0, q.c1, q.c2, q.c3), d=(q.d0, q.d1, q.d2, q.d3), window=qq.windowing, log_scale=bool(log),
                                 coeff=q.signalMultiplicator, vmin=q.signalGrayscaleMin, vmax=q.signalGrayscaleMax, addend=q.signalAddend)
            if log:
                keep = ref > np.quantile(ref, 0.05)
                assert np.abs(got - ref)[keep].max() < 1e-3, kw
            else:
                assert np.all(np.abs(got - ref) <= 2e-4 * np.abs(ref) + 1e-4 * np.sqrt(np.mean(ref.astype(np.float64) ** 2))), kw   # fp32 floor ~ rms of the line
            for m in range(4):
                assert np.float32(eo.ascan_metric(ref, N // 2, m, 0.3 if log else 30.0, 7)) == np.float32(eo.ref_metric(ref, N // 2, m, 0.3 if log else 30.0, 7))


class OraclePipeline:
    """stand-in for OctPipeline (TEST ONLY): dispersion_sweep served by the oracle, records what the engine asked for"""

    def __init__(self, q):
        self.params, self.calls, self.pushed = q, [], 0

    def push_params(self):
        self.pushed += 1

    def dispersion_sweep(self, raw, coeffs, metric, threshold, samples_to_ignore, log_scale, log_min, log_max, log_coeff, log_addend, want_ascans=False):
        q = self.params
        assert np.array_equal(q.windowCurve, cpu_path_window(N)), "the estimator path uses its own Hanning window"
        self.calls.append((np.array(raw).shape, np.array(coeffs)))
        a = np.stack([eo.cpu_process(raw, N, c=(q.c0, q.c1, q.c2, q.c3), d=tuple(co), log_scale=log_scale, coeff=log_coeff,
                                     vmin=log_min, vmax=log_max, addend=log_addend) for co in np.array(coeffs)])
        m = np.array([eo.ascan_metric(x, N // 2, metric, threshold, samples_to_ignore) for x in a], np.float32)
        return (m, a) if want_ascans else m


def test_engine_mirror_follows_the_reference_search():
    q = benchmark_params(N, 6, 1); q.update_all_curves()
    win_before = q.windowCurve.copy()
    pipe = OraclePipeline(q)
    eng = DispersionEstimationEngine(pipe)
    prm = DispersionEstimatorParameters(numberOfCenterAscans=4, useLinearAscans=True, numberOfAscanSamplesToIgnore=15, autoCalcD1=True,
                                        sharpnessMetric=PEAK_VALUE, metricThreshold=40.0, d2start=-160.0, d2end=0.0, d3start=-40.0,
                                        d3end=40.0, numberOfDispersionSamples=16)
    eng.setParams(prm)
    res = eng.startDispersionEstimation(G["raw"], 12, N, 6)
    # reference semantics on the same center block: offset (6-4)/2 = 1
    block = G["raw"][1:5]
    want = eo.estimate(lambda pairs: [eo.ascan_metric(oracle_ascans(d2, d3, 0, raw=block), N // 2, eo.PEAK_VALUE, 40.0, 15) for d2, d3 in pairs], EST)
    assert (res["bestD2"], res["bestD3"]) == (want["bestD2"], want["bestD3"]) and res["calculatedD1"] == -(want["bestD2"] + want["bestD3"])
    assert [m for _, m in res["metricsD2"]] == pytest.approx(want["metricD2"]) and [d for d, _ in res["metricsD2"]] == want["d2"]
    # two sweeps of 16 trials on the 4 center lines + the two plotted A-scans of one line
    assert [c[0] for c in pipe.calls] == [(4, N), (4, N), (1, N)] and [len(c[1]) for c in pipe.calls] == [16, 16, 2]
    assert np.all(pipe.calls[0][1][:, 3] == 0) and np.all(pipe.calls[1][1][:, 2] == np.float32(want["bestD2"]))
    assert np.all(pipe.calls[0][1][:, 0] == np.float32(q.d0)) and np.all(pipe.calls[0][1][:, 1] == np.float32(q.d1))
    assert eng.ascanWithBestDispersion.shape == (N // 2,) and eng.ascanWithBestDispersion.max() > eng.ascanWithoutDispersionCompensation[15:].max()
    # the main window setting is restored
    assert np.array_equal(q.windowCurve, win_before)
