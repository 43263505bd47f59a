"""The parameter cases behind the golden fixtures (shared by the generators and the tests)."""
from __future__ import annotations

import copy

from octproz_b200 import benchmark_params

GOLDEN_GEOMETRY = (16, 2)   # ascansPerBscan, bscansPerBuffer of the chain fixtures

LUT_CASES = [
    # (N, (c0..c3), (d0..d3), window type, centre, fill)
    (1024, (0.535239, 871.817574, -170.633784, 97.249716), (0.0, 97.0, -96.625, -0.375), 0, 0.5, 0.95),   # benchmark INI
    (2048, (0.5, 1800.0, -300.0, 150.0), (1.0, -40.0, 25.5, 3.25), 1, 0.45, 0.8),
    (1664, (0.0, 1664.0, 0.0, 0.0), (0.0, 0.0, 0.0, 0.0), 2, 0.5, 1.0),                                   # default settings.ini width
    (1024, (-3.0, 1100.0, 10.0, -5.0), (0.3, 10.0, 0.0, -2.0), 3, 0.3, 0.5),                              # clamps at both ends
    (512, (2.0, 400.0, 50.0, 20.0), (0.0, 5.0, 5.0, 5.0), 4, 0.9, 0.4),
    (1024, (0.0, 1000.0, 0.0, 0.0), (0.0, 0.0, 50.0, 0.0), 5, 0.5, 0.9),
    (100, (1.0, 90.0, 3.0, 1.0), (0.0, 1.0, 2.0, 3.0), 0, 1.5, 0.7),                                      # centre clamped to 1
]


def chain_cases(n: int = 1024):
    a, b = GOLDEN_GEOMETRY
    base = benchmark_params(n, a, b, 12)
    base.fixedPatternNoiseRemoval = False
    cases = {}
    cases["benchmark_nofpn"] = copy.deepcopy(base)
    q = copy.deepcopy(base); q.fixedPatternNoiseRemoval = True; q.bscansForNoiseDetermination = 2; cases["benchmark_fpn"] = q
    q = copy.deepcopy(base); q.resamplingInterpolation = 0; cases["linear"] = q
    q = copy.deepcopy(base); q.resamplingInterpolation = 2; cases["lanczos"] = q
    q = copy.deepcopy(base); q.resampling = False; cases["noresample"] = q
    q = copy.deepcopy(base); q.windowing = False; q.dispersionCompensation = False; cases["klin_only"] = q
    q = copy.deepcopy(base); q.resampling = False; q.windowing = False; q.dispersionCompensation = False; cases["fft_only"] = q
    q = copy.deepcopy(base); q.windowing = False; cases["klin_disp"] = q
    q = copy.deepcopy(base); q.backgroundRemoval = True; q.rollingAverageWindowSize = 16; cases["rolling16"] = q
    q = copy.deepcopy(base); q.bscanFlip = True; q.sinusoidalScanCorrection = True; cases["flip_sinus"] = q
    q = copy.deepcopy(base); q.signalLogScaling = False; q.signalGrayscaleMin = 0.0; q.signalGrayscaleMax = 400.0; cases["linscale"] = q
    q = copy.deepcopy(base); q.bitshift = True; q.bitDepth = 16; cases["bitshift16"] = q
    q = copy.deepcopy(base); q.postProcessBackgroundRemoval = True; q.postProcessBackgroundWeight = 0.5; q.postProcessBackgroundOffset = 0.01; cases["ppbg"] = q
    return cases
