"""GPU dispersion sweep (octb200_dispersion_sweep, octproz_b200/dispersion_estimator.py) through the C ABI against the oracle of
the reference's estimator path (oracle/estimator_oracle.py, pinned to the reference's own CPU code) and its golden vectors."""
import copy
import os

import numpy as np
import pytest

from octproz_b200 import OctPipeline, _lib, benchmark_params, synth
from octproz_b200.dispersion_estimator import (DispersionEstimationEngine, DispersionEstimatorParameters, PEAK_VALUE,
                                              cpu_path_window)
from oracle import estimator_oracle as eo

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "estimator.npz"))


def make_pipe(n, lines, **kw):
    q = benchmark_params(n, lines, 1)
    for k, v in kw.items():
        setattr(q, k, v)
    q.update_all_curves()
    q.windowCurve = cpu_path_window(n); q.windowUpdated = True          # the estimator path's own window
    p = OctPipeline(fft_mode=_lib.FFT_FUSED)
    assert p.initializeCuda(None, None, q), getattr(p, "_create_error", "")
    return p, q


def amp_close(got, ref, rel=1e-4):
    """linear amplitudes: 1e-4 relative (north_star) plus the fp32 floor of the reference's own float path (~1e-4 of the line rms)"""
    rms = np.sqrt(np.mean(ref.astype(np.float64) ** 2))
    bad = np.abs(got.astype(np.float64) - ref) > rel * np.abs(ref) + 1e-4 * rms
    return bad.mean()


@pytest.mark.parametrize("log", [0, 1])
def test_sweep_matches_reference_cpu_path_golden(log):
    n, lines = 1024, G["raw"].shape[0]
    p, q = make_pipe(n, lines)
    lg = G["log"]
    co = np.array([[G["d01"][0], G["d01"][1], d2, d3] for d2, d3 in G["trials"]], np.float32)
    thr = float(G[f"thr_log{log}"])
    for m in range(4):
        for j, ig in enumerate((0, 15)):
            metrics, a = p.dispersion_sweep(G["raw"], co, m, thr, ig, bool(log), lg[0], lg[1], lg[2], lg[3], want_ascans=True)
            ref = G[f"ascans_log{log}"]
            if log:
                keep = ref > np.quantile(ref, 0.05)
                assert np.abs(a - ref)[keep].max() < 5e-4
            else:
                assert amp_close(a, ref) == 0
            # the metric kernel sums in the reference's order: bit-identical to the reference metric code on the SAME A-scans ...
            for k in range(len(co)):
                assert metrics[k] == np.float32(eo.ascan_metric(a[k], n // 2, m, thr, ig)), (m, ig, k)
            # ... and close to the reference's value on ITS A-scans (thresholded metrics flip single samples at the threshold)
            want = G[f"metrics_log{log}"][m, j]
            tol = 2e-3 if m in (0, 1) else 3e-4
            assert np.allclose(metrics, want, rtol=tol, atol=1.0 if m == 1 else 1e-3), (m, ig, metrics, want)
    p.cleanupCuda()


@pytest.mark.parametrize("n,kw", [(1024, {}), (2048, {}), (1024, dict(backgroundRemoval=True, rollingAverageWindowSize=10)),
                                  (1024, dict(resampling=False)), (2048, dict(windowing=False)), (1024, dict(resamplingInterpolation=0))])
def test_sweep_matches_oracle_and_single_trials(n, kw):
    lines, trials = 5, 9
    p, q = make_pipe(n, lines, **kw)
    raw = synth.make_volume(n, lines, 1, 12, resample=q.resampleCurve, dispersion=-q.dispersionCurve).reshape(lines, n)
    rng = np.random.default_rng(3)
    co = np.stack([np.full(trials, q.d0), np.full(trials, q.d1), rng.uniform(-150, 50, trials), rng.uniform(-30, 30, trials)], 1).astype(np.float32)
    metrics, a = p.dispersion_sweep(raw, co, PEAK_VALUE, 0.0, 20, False, want_ascans=True)
    if q.resamplingInterpolation == 1 or not q.resampling:          # the CPU path only knows the cubic interpolator
        for k in range(trials):
            ref = eo.cpu_process(raw, n, remove_dc=q.backgroundRemoval, rolling_window=q.rollingAverageWindowSize, resample=q.resampling,
                                 c=(q.c0, q.c1, q.c2, q.c3), d=tuple(co[k]), window=q.windowing, log_scale=False)
            assert amp_close(a[k], ref) < 1e-4, (k, amp_close(a[k], ref))
    # batched launch == one launch per trial, bit for bit (same kernel, same tables)
    for k in (0, trials // 2, trials - 1):
        m1, a1 = p.dispersion_sweep(raw, co[k:k + 1], PEAK_VALUE, 0.0, 20, False, want_ascans=True)
        assert np.array_equal(a1[0], a[k]) and m1[0] == metrics[k]
    # device-resident raw data is accepted too
    import torch
    m2 = p.dispersion_sweep(torch.from_numpy(raw.view(np.int16)).cuda(), co, PEAK_VALUE, 0.0, 20, False)
    assert np.array_equal(m2, metrics)
    p.cleanupCuda()


def test_engine_finds_the_dispersion_of_the_sample_and_agrees_with_the_reference_search():
    n, lines = 1024, G["raw"].shape[0]
    p, q = make_pipe(n, lines)
    eng = DispersionEstimationEngine(p)
    eng.setParams(DispersionEstimatorParameters(numberOfCenterAscans=lines, useLinearAscans=True, numberOfAscanSamplesToIgnore=15,
                                                autoCalcD1=True, sharpnessMetric=PEAK_VALUE, metricThreshold=40.0, d2start=-160.0,
                                                d2end=0.0, d3start=-40.0, d3end=40.0, numberOfDispersionSamples=16))
    res = eng.startDispersionEstimation(G["raw"], 12, n, lines)
    # the golden search ran every trial through the reference's CPU path and metric code
    assert (res["bestD2"], res["bestD3"], res["calculatedD1"]) == tuple(G["search_log0"])
    assert np.allclose([m for _, m in res["metricsD2"]], G["search_metricD2_log0"], rtol=3e-4)
    assert np.allclose([m for _, m in res["metricsD3"]], G["search_metricD3_log0"], rtol=3e-4)
    # finer grid: the optimum sits at the dispersion the sample was synthesised with (d2 = -96.625, d3 = -0.375)
    eng.setParams(DispersionEstimatorParameters(numberOfCenterAscans=lines, useLinearAscans=False, numberOfAscanSamplesToIgnore=15,
                                                sharpnessMetric=PEAK_VALUE, d2start=-120.0, d2end=-70.0, d3start=-10.0, d3end=10.0,
                                                numberOfDispersionSamples=200))
    res = eng.startDispersionEstimation(G["raw"], 12, n, lines)
    assert abs(res["bestD2"] - q.d2) < 3.0 and abs(res["bestD3"] - q.d3) < 3.0, res["bestD2"]
    assert eng.ascanWithBestDispersion[15:].max() > eng.ascanWithoutDispersionCompensation[15:].max()
    p.cleanupCuda()


def test_sweep_error_behaviour():
    q = benchmark_params(1664, 8, 1); q.update_all_curves()
    p = OctPipeline()
    assert p.initializeCuda(None, None, q)
    with pytest.raises(_lib.Octb200Error, match="1024 or 2048"):
        p.dispersion_sweep(np.zeros((8, 1664), np.uint16), np.zeros((2, 4), np.float32), 0, 0.0, 0, False)
    p.cleanupCuda()
    p, q = make_pipe(1024, 4)
    with pytest.raises(_lib.Octb200Error):
        p.dispersion_sweep(np.zeros((4, 1024), np.uint16), np.zeros((2, 4), np.float32), 7, 0.0, 0, False)      # unknown metric
    with pytest.raises(_lib.Octb200Error):
        p.dispersion_sweep(np.zeros((4, 1024), np.uint16), np.zeros((2, 4), np.float32), 0, 0.0, 0, True, 5.0, 5.0)   # empty log range
    p.cleanupCuda()
