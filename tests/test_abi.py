"""The C-ABI library loads on a machine without a GPU and exports exactly what include/octb200.h declares."""
import ctypes as C
import os
import re

import pytest

from octproz_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "octb200.h")).read()
    return re.findall(r"OCTB200_API\s+[\w\s\*]+?\b(octb200_\w+)\s*\(", txt)


def test_every_declared_symbol_is_exported():
    L = _lib.load()
    declared = header_symbols()
    assert len(declared) >= 40
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, missing
    assert sorted(set(declared)) == sorted(set(_lib.SYMBOLS))


def test_version_and_defaults():
    L = _lib.load()
    assert L.octb200_version() == 100
    p = _lib.Params()
    L.octb200_default_params(C.byref(p))
    # octalgorithmparameters.cpp:36-112
    assert p.signalGrayscaleMax == 60.0 and p.signalMultiplicator == 1.0 and p.rollingAverageWindowSize == 1
    assert p.bscansForNoiseDetermination == 1 and p.postProcessBackgroundWeight == 1.0 and p.resampling == 0


def test_struct_layout_is_plain_c():
    assert C.sizeof(_lib.Config) == 48
    assert C.sizeof(_lib.Params) == 112


def test_create_rejects_bad_geometry_without_touching_the_gpu():
    L = _lib.load()
    h = C.c_void_p()
    cfg = _lib.Config(7, 4, 1, 1, 12, -1, 0, 0, 0)            # odd / tiny samplesPerLine
    assert L.octb200_create(C.byref(cfg), C.byref(h)) == _lib.ERR_INVALID and not h
    assert b"geometry" in L.octb200_last_error(None)
    cfg = _lib.Config(1 << 16, 1 << 10, 1 << 6, 1, 12, -1, 0, 0, 0)   # >= 2^31 samples (the reference's int indexing limit)
    assert L.octb200_create(C.byref(cfg), C.byref(h)) == _lib.ERR_INVALID
    assert L.octb200_create(None, C.byref(h)) == _lib.ERR_INVALID


def test_no_cpu_fallback_without_gpu():
    from tests.conftest import has_gpu
    if has_gpu():
        pytest.skip("GPU present")
    L = _lib.load()
    h = C.c_void_p()
    cfg = _lib.Config(1024, 16, 2, 1, 12, -1, 0, 0, 0)
    rc = L.octb200_create(C.byref(cfg), C.byref(h))
    assert rc == _lib.ERR_CUDA and not h        # fails loudly, never computes on the host


def test_header_is_plain_c99_and_links(tmp_path):
    """include/octb200.h compiles as strict C99 and a C program drives the library through it (tests/abi/abi_c99.c)"""
    import subprocess
    exe = str(tmp_path / "abi_c99")
    libdir = os.path.join(ROOT, "octproz_b200")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Wpedantic", "-Werror", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "abi", "abi_c99.c"), "-o", exe, "-L" + libdir, "-loctb200", "-Wl,-rpath," + libdir])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
    assert "abi ok" in out.stdout
    assert f"sizeof_config {C.sizeof(_lib.Config)} sizeof_params {C.sizeof(_lib.Params)}" in out.stdout


def test_adapter_gl_interop_branch_compiles():
    """the adapter's CUDA-GL interop branch (cuda_registerGlBuffer*, PBO mapping in octCudaPipeline, the volume view
    through octb200_volume_u8 + cudaMemcpy3DAsync, cuda_code.cu:1310-1355,1607-1695) is compiled by oracle/Makefile against the toolkit's
    cuda_gl_interop.h and the <GL/gl.h> stand-in (there is no libGL in the image: compile-only); the object must define the three
    registration entry points with the GL branch's code behind them"""
    import subprocess
    obj = os.path.join(ROOT, "oracle", "_ref", "adapter_gl.o")
    if not os.path.exists(obj):
        pytest.skip("oracle/_ref/adapter_gl.o is built where /root/reference exists")
    syms = subprocess.run(["nm", "-C", obj], capture_output=True, text=True, check=True).stdout
    for name in ("cuda_registerGlBufferBscan", "cuda_registerGlBufferEnFaceView", "cuda_registerGlBufferVolumeView", "octCudaPipeline"):
        assert f" T {name}" in syms, name
    for need in ("cudaGraphicsGLRegisterBuffer", "cudaGraphicsGLRegisterImage", "cudaGraphicsSubResourceGetMappedArray", "cudaMemcpy3DAsync", "octb200_volume_u8"):
        assert f" U {need}" in syms, need


def test_fft_path_query_and_plans_without_a_gpu():
    """host logic of the kernel selection (include/octb200.h octb200_query_fft_path): the Stockham plan of the shared-memory kernel
    multiplies out to the line length with radices from {2, 3, 4, 5, 7, 8, 11, 13}, odd primes first; AUTO's policy per geometry"""
    lib = _lib.load()
    def q(n, bits=12):
        rad = (C.c_int32 * 16)(); cnt = C.c_int32()
        rc = lib.octb200_query_fft_path(n, bits, rad, C.byref(cnt))
        return rc, list(rad[: cnt.value])
    assert q(1024)[0] == _lib.PATH_REGISTER_KERNEL and q(2048, 8)[0] == _lib.PATH_REGISTER_KERNEL and q(1024, 32)[0] == _lib.PATH_REGISTER_KERNEL
    assert q(1664) == (_lib.PATH_SHARED_MEMORY_KERNEL, [13, 8, 4, 4])          # the reference's default line length: 13 * 128
    assert q(4096) == (_lib.PATH_SHARED_MEMORY_KERNEL, [8, 8, 8, 8])
    assert q(8192) == (_lib.PATH_SHARED_MEMORY_KERNEL, [8, 8, 8, 4, 4])
    assert q(1536) == (_lib.PATH_SHARED_MEMORY_KERNEL, [3, 8, 8, 8])
    assert q(8190)[1] == [13, 7, 5, 3, 3, 2]
    assert q(512) == (_lib.PATH_CUFFT_CHAIN_SHARED_AVAILABLE, [8, 8, 8]) and q(16) == (_lib.PATH_CUFFT_CHAIN_SHARED_AVAILABLE, [4, 4])
    assert q(1006)[0] == _lib.PATH_CUFFT_CHAIN and q(34)[0] == _lib.PATH_CUFFT_CHAIN and q(8232)[0] == _lib.PATH_CUFFT_CHAIN and q(16384)[0] == _lib.PATH_CUFFT_CHAIN
    assert q(1001)[0] == _lib.ERR_INVALID and q(4)[0] == _lib.ERR_INVALID and q(1024, 40)[0] == _lib.ERR_INVALID
    for n in range(8, 8193, 2):
        rc, rad = q(n)
        if rc in (_lib.PATH_SHARED_MEMORY_KERNEL, _lib.PATH_CUFFT_CHAIN_SHARED_AVAILABLE):
            prod = 1
            for r in rad:
                assert r in (2, 3, 4, 5, 7, 8, 11, 13)
                prod *= r
            assert prod == n and len(rad) <= 16, (n, rad)
            odd = [r for r in rad if r % 2]
            assert rad[: len(odd)] == odd == sorted(odd, reverse=True), (n, rad)      # odd primes first, descending
            assert sum(1 for r in rad if r == 2) <= 1
        elif rc == _lib.PATH_CUFFT_CHAIN:
            m = n
            for pr in (2, 3, 5, 7, 11, 13):
                while m % pr == 0:
                    m //= pr
            assert m != 1, n                 # only lengths with a prime factor above 13 are left to cuFFT below 8192
