"""Randomised curve-generator parity: the product's generators (octb200_make_*, called without a GPU) and the oracle's against
each other on 200 seeded parameter sets, and -- where the reference's own host code was built (oracle/_ref/libref_luts.so, this
container and any box the .so travelled to) -- against that code directly.  Bit-exact: these LUTs feed every kernel."""
import numpy as np
import pytest

from octproz_b200 import _lib
from oracle import oracle as orc
from tests.random_configs import random_lut_case

SEED = 0x0C7B200


def product(n, c, d, wt, ce, fi):
    L = _lib.load()
    r, dd, w = (np.empty(n, np.float32) for _ in range(3))
    assert L.octb200_make_resample_curve(n, *c, r.ctypes.data) == 0
    assert L.octb200_make_dispersion_curve(n, *d, dd.ctypes.data) == 0
    assert L.octb200_make_window_curve(wt, ce, fi, n, w.ctypes.data) == 0
    return r, dd, w


def test_product_and_oracle_generators_agree_on_random_parameters():
    rng = np.random.default_rng(SEED)
    for i in range(200):
        n, c, d, wt, ce, fi = random_lut_case(rng)
        r, dd, w = product(n, c, d, wt, ce, fi)
        assert np.array_equal(r, orc.resample_curve(n, *c)), (i, n, c)
        assert np.array_equal(dd, orc.dispersion_curve(n, *d)), (i, n, d)
        assert np.array_equal(w, orc.window_curve(wt, ce, fi, n)), (i, n, wt, ce, fi)
        assert r.min() >= 0.0 and r.max() <= n - 3        # Polynomial::clamp, octalgorithmparameters.cpp:167
        assert np.isfinite(w).all() and w.min() >= -0.1 and w.max() <= 1.0 + 1e-6


@pytest.mark.skipif(not orc.have_ref("libref_luts.so"), reason="oracle/_ref/libref_luts.so not built (needs /root/reference)")
def test_generators_match_the_reference_host_code_on_random_parameters():
    rng = np.random.default_rng(SEED + 1)
    for i in range(120):
        n, c, d, wt, ce, fi = random_lut_case(rng)
        want = orc.ref_luts(n, c, d, wt, ce, fi)
        got = product(n, c, d, wt, ce, fi)
        for name, g, w in zip(("resample", "dispersion", "window"), got, want):
            assert np.array_equal(g, w), (i, name, n, c, d, wt, ce, fi, float(np.abs(g - w).max()))
