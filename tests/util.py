"""Shared parity metric of the tests.

north_star: "results match the reference CUDA path within 1e-4 relative float tolerance".
The displayed value is a log of the power, which is unbounded at spectral nulls, so the tolerance is stated on
the LINEAR AMPLITUDE both implementations computed before the log:

    | amp_a - amp_b |  <=  RTOL * amp_b  +  ATOL_FRAC * median(amp_b)        RTOL = ATOL_FRAC = 1e-4

i.e. 1e-4 relative, with an absolute floor of 1e-4 of the buffer's median (noise-floor) amplitude -- the size of
fp32 FFT round-off, which two different fp32 FFTs (cuFFT vs ours) cannot agree below.  For linear scaling the
output is proportional to the amplitude and the same formula is applied to it directly.
"""
from __future__ import annotations

import numpy as np

RTOL = 1e-4
ATOL_FRAC = 1e-4


def amplitude(out: np.ndarray, q) -> np.ndarray:
    """invert postProcessTruncateLog / Lin (cuda_code.cu:718 / :739) back to |X|"""
    o = out.astype(np.float64)
    h = q.samplesPerLine / 2.0
    v = (o / q.signalMultiplicator - q.signalAddend) * (q.signalGrayscaleMax - q.signalGrayscaleMin) + q.signalGrayscaleMin
    if q.signalLogScaling:
        return np.sqrt(h) * 10.0 ** (v / 20.0)
    return v * h


def parity_report(out, ref, q, rtol=RTOL, atol_frac=ATOL_FRAC, saturated=False, atol_abs=0.0):
    if saturated:   # after postProcessBackgroundRemoval the value is clamped to [0,1]: compare the displayed value
        a, b = out.astype(np.float64), ref.astype(np.float64)
        tol = rtol * np.abs(b) + atol_frac
    else:
        a, b = amplitude(out, q), amplitude(ref, q)
        tol = rtol * np.abs(b) + atol_frac * np.median(np.abs(b)) + atol_abs
    both_nonfinite = ~np.isfinite(a) & ~np.isfinite(b)
    d = np.abs(a - b)
    d[both_nonfinite] = 0.0
    ratio = d / tol
    return {"max_ratio": float(np.nanmax(ratio)), "frac_outside": float((~(ratio <= 1.0)).mean()),
            "median_amp": float(np.median(np.abs(b)))}


def assert_parity(out, ref, q, rtol=RTOL, atol_frac=ATOL_FRAC, max_frac_outside=0.0, saturated=False, what="", atol_abs=0.0):
    """atol_abs: extra absolute amplitude floor for cases whose round-off is set by a much larger cancelled term
    (fixed-pattern-noise subtraction: 4 eps32 |meanLine|max; degenerate all-in-one-bin inputs: eps32 * max amplitude)"""
    r = parity_report(out, ref, q, rtol, atol_frac, saturated, atol_abs)
    assert r["frac_outside"] <= max_frac_outside, f"{what}: {r} (rtol={rtol}, atol_frac={atol_frac})"
    return r
