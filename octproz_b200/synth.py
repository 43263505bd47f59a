"""Deterministic synthetic OCT raw data (the reference's test dataset is an external download and the
reference holds no fixtures for this path).

Spectral interferograms of a few tilted reflectors, sampled non-linearly in k (inverse of the benchmark
resampling polynomial, so k-linearisation matters), chirped by the negative benchmark dispersion
polynomial (so dispersion compensation matters), with a fixed-pattern term identical in every A-scan
(so FPN removal matters) and additive Gaussian noise.  Layout = the Virtual OCT System's raw file:
headerless little-endian containers [B][A][N] (virtualoctsystem.cpp:163-184, docs/docs/faq.md:5).
"""
from __future__ import annotations

import numpy as np

SEED = 0x0C7B200

REFLECTORS = ((60.0, 1.0), (150.3, 0.3), (300.7, 0.1), (420.0, 0.03))


def container_dtype(bit_depth: int):
    return np.uint8 if bit_depth <= 8 else (np.uint16 if bit_depth <= 16 else np.uint32)


def _k_axis(n: int, resample: np.ndarray | None) -> np.ndarray:
    m = np.arange(n, dtype=np.float64)
    if resample is None:
        return m
    r = resample.astype(np.float64)
    r_mono = np.maximum.accumulate(r)
    return np.interp(m, r_mono, np.arange(n, dtype=np.float64))


def make_volume(n: int, a: int, b: int, bit_depth: int = 12, seed: int = SEED, resample: np.ndarray | None = None,
                dispersion: np.ndarray | None = None, b_offset: int = 0, noise_lsb: float | None = None) -> np.ndarray:
    """returns [b][a][n] raw samples in the container type of `bit_depth`"""
    full = float(2 ** bit_depth - 1)
    dc = 0.45 * (2 ** bit_depth)
    sigma = noise_lsb if noise_lsb is not None else (8.0 if bit_depth <= 12 else 8.0 * 2 ** (bit_depth - 12))
    m = np.arange(n, dtype=np.float64)
    kappa = _k_axis(n, resample)
    fwhm = 0.6 * n
    env = np.exp(-4.0 * np.log(2.0) * ((m - n / 2.0) / fwhm) ** 2)
    theta = np.zeros(n)
    if dispersion is not None:
        theta = -np.interp(kappa, np.arange(n, dtype=np.float64), dispersion.astype(np.float64))
    fp = 0.05 * np.cos(2.0 * np.pi * 200.0 * m / n)
    out = np.empty((b, a, n), container_dtype(bit_depth))
    k32 = (2.0 * np.pi * kappa / n).astype(np.float32)
    th32 = theta.astype(np.float32)
    base = (dc * env).astype(np.float32)
    fp32 = fp.astype(np.float32)
    aa = np.arange(a, dtype=np.float32)[:, None]
    for bi in range(b):
        bg = bi + b_offset
        rng = np.random.Generator(np.random.Philox(key=seed + bg))
        acc = np.ones((a, n), np.float32) + fp32[None, :]
        for z0, rho in REFLECTORS:
            z = z0 * n / 1024.0 + 0.02 * aa * (n / 1024.0) + 0.05 * bg * (n / 1024.0)     # slow tilt in a and b
            acc += np.float32(rho) * np.cos(z * k32[None, :] + th32[None, :])
        sig = base[None, :] * acc + rng.standard_normal((a, n), dtype=np.float32) * np.float32(sigma)
        out[bi] = np.clip(np.rint(sig), 0.0, full).astype(out.dtype)
    return out


def adversarial_lines(n: int, bit_depth: int = 12) -> np.ndarray:
    """[5][n]: all-zero, full-scale, single impulse, Nyquist tone, ramp"""
    full = 2 ** bit_depth - 1
    dt = container_dtype(bit_depth)
    z = np.zeros((5, n), dt)
    z[1] = full
    z[2, n // 3] = full
    z[3, ::2] = full
    z[4] = (np.arange(n) * full // max(n - 1, 1)).astype(dt)
    return z
