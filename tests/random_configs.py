"""Seeded random parameter sets over the whole configuration vocabulary of the path, shared by the CPU
tests (curve generators against the reference's own host code) and the GPU tests (chain against the oracle)."""
from __future__ import annotations

import numpy as np

from octproz_b200 import OctAlgorithmParameters


def random_lut_case(rng: np.random.Generator):
    """(N, (c0..c3), (d0..d3), windowType, center, fill) like tests/golden/cases.LUT_CASES"""
    n = int(rng.choice([64, 100, 512, 1024, 1664, 2048, 4096]))
    c = (float(rng.uniform(-2, 4)), float(rng.uniform(0.5, 1.1) * (n - 1)), float(rng.uniform(-0.3, 0.3) * (n - 1)), float(rng.uniform(-0.2, 0.2) * (n - 1)))
    d = tuple(float(x) for x in rng.uniform(-120, 120, 4))
    return n, c, d, int(rng.integers(0, 6)), float(rng.uniform(0.2, 0.8)), float(rng.uniform(0.1, 1.0))


def random_chain_config(rng: np.random.Generator, n: int, a: int = 27, b: int = 2) -> tuple[OctAlgorithmParameters, dict]:
    """a random but valid processing configuration at line length n; returns (params with curves built, extras).
    a >= 27 keeps at least three A-scans in each of the nine segments of the FPN determination (cuda_code.cu:531): with one A-scan per
    segment the "mean" is that A-scan itself and its own line becomes log(0)."""
    bits = int(rng.choice([10, 12, 14, 16]))
    q = OctAlgorithmParameters(samplesPerLine=n, ascansPerBscan=a, bscansPerBuffer=b, buffersPerVolume=1, bitDepth=bits)
    q.bitshift = bool(bits == 16 and rng.random() < 0.5)
    q.resampling = bool(rng.random() < 0.8)
    q.resamplingInterpolation = int(rng.integers(0, 3))
    s = n / 1024.0
    q.c0, q.c1 = float(rng.uniform(0, 3)), float(rng.uniform(780, 1000) * s)
    q.c2, q.c3 = float(rng.uniform(-200, 100) * s), float(rng.uniform(-50, 120) * s)
    q.dispersionCompensation = bool(rng.random() < 0.7)
    # |phase| stays below ~25 rad, the range of the published benchmark coefficients: the product (like the reference, under
    # --use_fast_math) takes cos/sin of the phase from the MUFU approximations, whose range reduction loses ~|phase| * 2^-24 rad --
    # against the oracle's exact cos/sin that is only below the 1e-4 tolerance for moderate phases
    q.d0, q.d1, q.d2, q.d3 = 0.0, float(rng.uniform(-12, 12)), float(rng.uniform(-10, 10)), float(rng.uniform(-3, 3))
    q.windowing = bool(rng.random() < 0.8)
    q.window = int(rng.integers(0, 6)); q.windowFillFactor = float(rng.uniform(0.4, 1.0)); q.windowCenter = float(rng.uniform(0.35, 0.65))
    q.backgroundRemoval = bool(rng.random() < 0.3); q.rollingAverageWindowSize = int(rng.integers(1, 120))
    q.signalLogScaling = bool(rng.random() < 0.7)
    if q.signalLogScaling:
        q.signalGrayscaleMin, q.signalGrayscaleMax = float(rng.uniform(-40, 20)), float(rng.uniform(60, 120))
    else:
        q.signalGrayscaleMin, q.signalGrayscaleMax = 0.0, float(rng.uniform(50, 4000))
    q.signalMultiplicator, q.signalAddend = float(rng.uniform(0.5, 2.0)), float(rng.uniform(-0.2, 0.2))
    q.bscanFlip = bool(rng.random() < 0.4)
    q.sinusoidalScanCorrection = bool(rng.random() < 0.3)
    q.fixedPatternNoiseRemoval = bool(rng.random() < 0.5); q.bscansForNoiseDetermination = int(rng.integers(1, b + 1))
    q.postProcessBackgroundRemoval = bool(rng.random() < 0.25)
    q.postProcessBackgroundWeight, q.postProcessBackgroundOffset = float(rng.uniform(0.2, 1.2)), float(rng.uniform(-0.05, 0.05))
    q.update_all_curves()
    extras = {"pp_background": (rng.uniform(0.0, 0.3, n // 2)).astype(np.float32) if q.postProcessBackgroundRemoval else None}
    return q, extras


def describe(q) -> str:
    keys = ("samplesPerLine", "bitDepth", "bitshift", "resampling", "resamplingInterpolation", "dispersionCompensation", "windowing", "window",
            "backgroundRemoval", "rollingAverageWindowSize", "signalLogScaling", "bscanFlip", "sinusoidalScanCorrection",
            "fixedPatternNoiseRemoval", "bscansForNoiseDetermination", "postProcessBackgroundRemoval")
    return ", ".join(f"{k}={getattr(q, k)}" for k in keys)
