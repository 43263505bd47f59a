"""CPU execution of the fused kernel's per-lane phases (octproz_b200/csrc/oct_phases.cuh, compiled for the host by
tests/emu/emu_phases.cpp): validates the register / lane / shared-memory index maps of the 32x32 four-step FFT, the
R=2 radix-2 combine, the LUT folding of stage A and the epilogue against numpy and the oracle -- without a GPU.
TEST-ONLY emulator; the product never runs these functions on the host."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from octproz_b200 import benchmark_params, synth
from oracle import oracle as orc
from tests.util import assert_parity

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("emu") / "libemu.so")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-I/usr/local/cuda/include",
                           "-I" + os.path.join(ROOT, "octproz_b200", "csrc"), os.path.join(ROOT, "tests", "emu", "emu_phases.cpp"), "-o", so])
    L = C.CDLL(so)
    f = C.c_float
    L.emu_fused_line.argtypes = [C.c_int] * 3 + [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, f, f, f, f,
                                 C.c_void_p, C.c_void_p, f, f, C.c_void_p, C.c_void_p]
    return L


def test_fft32_register_network(emu):
    rng = np.random.default_rng(0)
    x = (rng.standard_normal(32) + 1j * rng.standard_normal(32)).astype(np.complex64)
    o = np.zeros(32, np.complex64)
    emu.emu_fft32(x.ctypes.data_as(C.c_void_p), o.ctypes.data_as(C.c_void_p))
    ref = np.fft.ifft(x.astype(np.complex128)) * 32
    assert np.abs(o - ref).max() < 5e-7 * np.abs(ref).max()


def test_four_step_1024(emu):
    rng = np.random.default_rng(1)
    x = (rng.standard_normal(1024) + 1j * rng.standard_normal(1024)).astype(np.complex64)
    o = np.zeros(1024, np.complex64)
    emu.emu_ifft_1024(x.ctypes.data_as(C.c_void_p), o.ctypes.data_as(C.c_void_p))
    ref = np.fft.ifft(x.astype(np.complex128)) * 1024
    assert np.abs(o - ref).max() < 1e-6 * np.abs(ref).max()


@pytest.mark.parametrize("n", [1024, 2048])
@pytest.mark.parametrize("case", ["cubic", "linear", "lanczos", "none"])
def test_fused_line_matches_oracle(emu, n, case):
    interp, sa = {"cubic": (1, 3), "linear": (0, 1), "lanczos": (2, 2), "none": (0, 0)}[case]     # SA_CUBIC=3, SA_LINEAR=1, SA_LANCZOS=2, SA_NONE=0
    q = benchmark_params(n, 1, 1); q.fixedPatternNoiseRemoval = False
    q.resampling = case != "none"; q.resamplingInterpolation = interp
    q.update_all_curves()
    raw = synth.make_volume(n, 1, 1, 12, resample=q.resampleCurve, dispersion=q.dispersionCurve)
    ref, _, _ = orc.process(q, raw)
    fs = np.zeros(n + 64, np.float32); fs[16:16 + n] = raw.reshape(-1)
    d = q.dispersionCurve.astype(np.float64)
    ph = np.stack([np.cos(d), np.sin(d)], 1).astype(np.float32)
    out = np.zeros(n // 2, np.float32)
    rc = emu.emu_fused_line(n, sa, interp, fs[16:].ctypes.data, 8 if sa == 2 else 0,
                            q.resampleCurve.ctypes.data if q.resampling else None, q.windowCurve.ctypes.data, ph.ctypes.data,
                            1, q.signalGrayscaleMin, q.signalGrayscaleMax, q.signalMultiplicator, q.signalAddend,
                            None, None, 0.0, 0.0, out.ctypes.data, None)
    assert rc == 0
    assert_parity(out.reshape(1, 1, -1), ref, q, what=f"emulated fused line N={n} {case}")


@pytest.mark.parametrize("n", [8, 16, 48, 100, 130, 512, 640, 896, 1024, 1408, 1536, 1664, 2560, 3072, 4096, 4160, 5632, 8190, 8192])
def test_generic_mixed_radix_transform(emu, n):
    """the shared-memory kernel's inverse transform (octproz_b200/csrc/generic_fft.cuh: Stockham autosort passes over padded line
    buffers, radices 13 / 11 / 7 / 5 / 3 / 8 / 4 / 2, per-pass twiddle tables, j mod Ns by multiply-high) executed on the CPU with the
    kernel's own thread counts against numpy, for every radix and the lengths the GPU tests use"""
    rng = np.random.default_rng(n)
    x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    ref = np.fft.ifft(x.astype(np.complex128)) * n
    for threads in {64 if n < 1024 else (128 if n < 2048 else 256), 32}:
        o = np.zeros(n, np.complex64)
        rad = (C.c_int * 16)(); cnt = C.c_int()
        rc = emu.emu_generic_ifft(n, threads, x.ctypes.data_as(C.c_void_p), o.ctypes.data_as(C.c_void_p), rad, C.byref(cnt))
        assert rc == 0
        assert int(np.prod(list(rad[: cnt.value]))) == n
        assert np.abs(o - ref).max() < 2e-6 * np.abs(ref).max(), (n, threads, list(rad[: cnt.value]))


def test_generic_transform_refuses_large_prime_factors(emu):
    o = np.zeros(34, np.complex64)
    assert emu.emu_generic_ifft(34, 32, o.ctypes.data_as(C.c_void_p), o.ctypes.data_as(C.c_void_p), None, None) == -1      # 2 * 17
