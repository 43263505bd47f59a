"""Host curve generators: product (liboctb200.so octb200_make_*) and oracle (oracle/oct_oracle.c) against golden
LUTs produced by the reference's own host code (tests/golden/luts.npz, tests/golden/make_golden_luts.py).
Bit-exact: these LUTs feed every kernel."""
import ctypes as C
import os

import numpy as np
import pytest

from octproz_b200 import _lib
from oracle import oracle as orc
from tests.golden.cases import LUT_CASES

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "luts.npz"))


@pytest.mark.parametrize("i", range(len(LUT_CASES)))
def test_product_curves_match_reference_bit_exact(i):
    n, c, d, wt, ce, fi = LUT_CASES[i]
    L = _lib.load()
    r, dd, w = (np.empty(n, np.float32) for _ in range(3))
    assert L.octb200_make_resample_curve(n, *c, r.ctypes.data) == 0
    assert L.octb200_make_dispersion_curve(n, *d, dd.ctypes.data) == 0
    assert L.octb200_make_window_curve(wt, ce, fi, n, w.ctypes.data) == 0
    assert np.array_equal(r, GOLD[f"resample_{i}"])
    assert np.array_equal(dd, GOLD[f"dispersion_{i}"])
    assert np.array_equal(w, GOLD[f"window_{i}"])
    assert r.min() >= 0.0 and r.max() <= n - 3        # octalgorithmparameters.cpp:167


@pytest.mark.parametrize("i", range(len(LUT_CASES)))
def test_oracle_curves_match_reference_bit_exact(i):
    n, c, d, wt, ce, fi = LUT_CASES[i]
    assert np.array_equal(orc.resample_curve(n, *c), GOLD[f"resample_{i}"])
    assert np.array_equal(orc.dispersion_curve(n, *d), GOLD[f"dispersion_{i}"])
    assert np.array_equal(orc.window_curve(wt, ce, fi, n), GOLD[f"window_{i}"])


def test_known_answers_benchmark_window():
    # Hann 0.95/0.5, N=1024 -> w[26]=0, w[27]~1.0e-5, w[512]~0.999997 (zeroing rule xiNorm<0.0001, windowfunction.cpp:156)
    w = GOLD["window_0"]
    assert w[26] == 0.0 and 0.9e-5 < w[27] < 1.2e-5 and abs(w[512] - 0.999997) < 2e-6


@pytest.mark.parametrize("a", [1, 7, 512, 1024])
def test_sinusoidal_curve(a):
    L = _lib.load()
    s = np.empty(a, np.float32)
    assert L.octb200_make_sinusoidal_curve(a, s.ctypes.data) == 0
    assert np.array_equal(s, orc.sinusoidal_curve(a))
    k = np.arange(a, dtype=np.float64)
    expect = (a / np.pi) * np.arccos(1.0 - 2.0 * k / a)     # cuda_code.cu:519
    assert np.allclose(s, expect, rtol=2e-6, atol=2e-4)
    assert s[0] == 0.0 and np.all(np.diff(s) > 0) and s[-1] < a - 1 + 1e-3


def test_generators_reject_bad_arguments():
    L = _lib.load()
    buf = np.empty(8, np.float32)
    assert L.octb200_make_resample_curve(2, 0, 0, 0, 0, buf.ctypes.data) == _lib.ERR_INVALID
    assert L.octb200_make_window_curve(9, 0.5, 0.5, 8, buf.ctypes.data) == _lib.ERR_INVALID
    assert L.octb200_make_window_curve(0, 0.5, 0.5, 8, None) == _lib.ERR_INVALID
