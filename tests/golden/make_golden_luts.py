"""Writes tests/golden/luts.npz from the REFERENCE'S OWN host code
(octalgorithmparameters.cpp, polynomial.cpp, windowfunction.cpp compiled in place into
oracle/_ref/libref_luts.so by oracle/Makefile).  Run in the container that has /root/reference:
    python tests/golden/make_golden_luts.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as orc  # noqa: E402
from tests.golden.cases import LUT_CASES  # noqa: E402

out = {}
for i, (n, c, d, wt, ce, fi) in enumerate(LUT_CASES):
    r, dd, w = orc.ref_luts(n, c, d, wt, ce, fi)
    out[f"resample_{i}"], out[f"dispersion_{i}"], out[f"window_{i}"] = r, dd, w
np.savez_compressed(os.path.join(HERE, "luts.npz"), **out)
print("wrote luts.npz:", len(LUT_CASES), "cases")
