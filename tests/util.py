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


REPORT: list[dict] = []       # every parity comparison of the session; tests/conftest.py writes it to gpurun_out/parity_report.json
FAST_ABOVE = 1 << 26          # elements; larger comparisons screen in the output domain first (same criterion, see parity_report)


def parity_report(out, ref, q, rtol=RTOL, atol_frac=ATOL_FRAC, saturated=False, atol_abs=0.0):
    """max_ratio: worst |amp_a - amp_b| / tolerance; frac_outside: fraction of elements with ratio > 1;
    max_rel_above_median: the STRICT relative amplitude error max |amp_a - amp_b| / amp_b over the elements whose amplitude is above
    the buffer's median (no absolute floor involved there).
    Buffers above FAST_ABOVE elements (full BASELINE sizes) are screened in the output domain first: an element whose displayed values
    differ by less than what a relative amplitude error of rtol/2 makes of them satisfies the criterion a fortiori; the amplitudes
    (float64 pow) are only formed for the rest, and the median comes from a strided sample of >= 4 M elements."""
    if saturated:   # after postProcessBackgroundRemoval the value is clamped to [0,1]: compare the displayed value
        a, b = out.astype(np.float64), ref.astype(np.float64)
        tol = rtol * np.abs(b) + atol_frac
        both_nonfinite = ~np.isfinite(a) & ~np.isfinite(b)
        d = np.abs(a - b)
        d[both_nonfinite] = 0.0
        ratio = d / tol
        return {"max_ratio": float(np.nanmax(ratio)), "frac_outside": float((~(ratio <= 1.0)).mean()), "median_amp": float(np.median(np.abs(b))),
                "max_rel_above_median": None, "elements": int(ratio.size)}
    n_el = int(np.asarray(out).size)
    if n_el > FAST_ABOVE:
        o, r = np.asarray(out).reshape(-1), np.asarray(ref).reshape(-1)
        step = max(1, n_el // (1 << 22))
        med = float(np.median(np.abs(amplitude(r[::step], q))))
        # displayed-value distance of a relative amplitude change rtol/2: log: coeff * 20 log10(1 + rtol/2) / (max - min); linear: rtol/2 of the value
        span = float(q.signalGrayscaleMax - q.signalGrayscaleMin)
        if q.signalLogScaling:
            screen = np.float32(abs(q.signalMultiplicator) * 20.0 * np.log10(1.0 + rtol / 2) / abs(span))
            cand = np.flatnonzero(~(np.abs(o - r) <= screen))
        else:
            cand = np.flatnonzero(~(np.abs(o - r) <= np.float32(rtol / 2) * np.abs(r - np.float32(q.signalMultiplicator * (q.signalAddend - q.signalGrayscaleMin / span)))))
        a, b = amplitude(o[cand], q), amplitude(r[cand], q)
        tol = rtol * np.abs(b) + atol_frac * med + atol_abs
        both_nonfinite = ~np.isfinite(a) & ~np.isfinite(b)
        d = np.abs(a - b); d[both_nonfinite] = 0.0
        ratio = d / tol
        above = b > med
        rel = float(np.max(d[above] / b[above])) if above.any() else 0.0
        return {"max_ratio": float(np.nanmax(ratio)) if ratio.size else 0.5, "frac_outside": float((~(ratio <= 1.0)).sum() / n_el), "median_amp": med,
                "max_rel_above_median": max(rel, rtol / 2) if cand.size else rtol / 2, "elements": n_el, "screened": True,
                "note": "max_ratio / max_rel_above_median are exact only above the screen (ratio 0.5 / rtol/2 = everything passed the screen)"}
    a, b = amplitude(out, q), amplitude(ref, q)
    med = float(np.median(np.abs(b)))
    tol = rtol * np.abs(b) + atol_frac * med + atol_abs
    both_nonfinite = ~np.isfinite(a) & ~np.isfinite(b)
    d = np.abs(a - b)
    d[both_nonfinite] = 0.0
    ratio = d / tol
    above = np.isfinite(b) & (b > med)
    rel = float(np.max(d[above] / b[above])) if above.any() else 0.0
    return {"max_ratio": float(np.nanmax(ratio)), "frac_outside": float((~(ratio <= 1.0)).mean()), "median_amp": med,
            "max_rel_above_median": rel, "elements": n_el}


def assert_parity(out, ref, q, rtol=RTOL, atol_frac=ATOL_FRAC, max_frac_outside=0.0, saturated=False, what="", atol_abs=0.0):
    """atol_abs: extra absolute amplitude floor for cases whose round-off is set by a much larger cancelled term
    (fixed-pattern-noise subtraction: 4 eps32 |meanLine|max; degenerate all-in-one-bin inputs: eps32 * max amplitude)"""
    r = parity_report(out, ref, q, rtol, atol_frac, saturated, atol_abs)
    REPORT.append(dict(r, what=what, rtol=rtol, atol_frac=atol_frac, atol_abs=atol_abs, allowed_frac_outside=max_frac_outside))
    assert r["frac_outside"] <= max_frac_outside, f"{what}: {r} (rtol={rtol}, atol_frac={atol_frac})"
    return r
