"""Randomised chain parity on the GPU: 36 seeded random configurations over the whole parameter vocabulary (interpolators, windows,
dispersion, bit depths / bitshift, rolling background, FPN, flip, sinusoidal correction, log / linear scaling, background removal),
N = 1024 and 2048, every FFT mode, against the oracle on the same raw buffer.  Tolerances are those of tests/util.py with the
documented exceptions of DESIGN.md section 4 (Lanczos vs the fp64 oracle; FPN cancellation floor).  Where the reference's own
CUDA build travelled to the box (oracle/_ref/libref_cuda.so) the same configurations are also run through the UNMODIFIED
cuda_code.cu on the same raw buffer and compared at the strict 1e-4 tolerance, Lanczos included."""
import copy

import numpy as np
import pytest

from octproz_b200 import OctPipeline, _lib, synth
from oracle import oracle as orc
from tests.random_configs import describe, random_chain_config
from tests.util import assert_parity

pytestmark = pytest.mark.gpu
MODES = {"fused": _lib.FFT_FUSED, "split": _lib.FFT_SPLIT, "cufft": _lib.FFT_CUFFT}
SEED = 0x0C7B200 + 7


def oracle_reference(q, raw, extras):
    return orc.process(q, raw, pp_background=extras["pp_background"])


def gpu_run(q, raw, mode, mean_line, extras):
    qq = copy.deepcopy(q)
    p = OctPipeline(fft_mode=mode)
    assert p.initializeCuda(None, None, qq), getattr(p, "_create_error", "")
    if q.fixedPatternNoiseRemoval:
        p.set_fpn_mean_line(np.asarray(mean_line, np.float32))      # the determination itself is covered (and conditioned) elsewhere
    if extras["pp_background"] is not None:
        qq.loadPostProcessingBackground(extras["pp_background"])
    p.octCudaPipeline(np.ascontiguousarray(raw)); p.sync()
    out = p.copy_output(0)
    p.cleanupCuda()
    return out


@pytest.mark.parametrize("n", [1024, 2048])
@pytest.mark.parametrize("i", range(18))
def test_random_configuration_matches_oracle(i, n):
    rng = np.random.default_rng([SEED, n, i])
    q, extras = random_chain_config(rng, n)
    raw = synth.make_volume(n, q.ascansPerBscan, q.bscansPerBuffer, q.bitDepth, resample=q.resampleCurve if q.resampling else None,
                            dispersion=q.dispersionCurve if q.dispersionCompensation else None)
    ref, ml, _ = oracle_reference(q, raw, extras)
    lanczos = q.resampling and q.resamplingInterpolation == 2
    floor = 4e-6 * float(np.abs(ml).max()) if q.fixedPatternNoiseRemoval else 0.0
    for name, mode in MODES.items():
        out = gpu_run(q, raw, mode, ml, extras)
        # Lanczos weights come from __sinf like the reference's: against the fp64 oracle the floor is ~1e-2 of the median amplitude and a few
        # 1e-4 of the bins sit up to 3x above it once the rolling mean has removed the DC term (against the reference CUDA build
        # itself the strict bound holds: the test below and the golden vectors)
        assert_parity(out, ref, q, atol_frac=2e-2 if lanczos else 1e-4, max_frac_outside=1e-3 if lanczos else 1e-4, saturated=bool(q.postProcessBackgroundRemoval),
                      atol_abs=floor, what=f"random #{i} N={n} {name}: {describe(q)}")


@pytest.mark.skipif(not orc.have_ref("libref_cuda.so"), reason="oracle/_ref/libref_cuda.so not built")
@pytest.mark.parametrize("n", [1024, 2048])
@pytest.mark.parametrize("i", range(18))
def test_random_configuration_matches_live_reference_cuda(i, n):
    """the same seeded configurations through the reference's unmodified cuda_code.cu (its own curve generators, its own FPN
    determination, its own kernels) and through ours, same raw buffer, same box: 1e-4 relative on the amplitude for every mode"""
    rng = np.random.default_rng([SEED, n, i])
    q, extras = random_chain_config(rng, n)
    rc = orc.RefCuda(); rc.configure(q)
    ours = (q.resampleCurve, q.dispersionCurve, q.windowCurve)
    q.resampleCurve, q.dispersionCurve, q.windowCurve = rc.curves()
    # libref_cuda.so carries the reference's host code as nvcc's host pass compiles it (no -O2): its Gauss window differs from the
    # qmake-release (-O2) build of the same source by 1-3 ulp in a few elements.  Bit-identity of our generators is pinned against
    # the -O2 build (tests/test_random_curves.py, tests/golden/luts.npz); here the curves only have to be the same curves.
    for mine, theirs in zip(ours, (q.resampleCurve, q.dispersionCurve, q.windowCurve)):
        assert np.allclose(mine, theirs, rtol=1e-6, atol=1e-7), "curve generators vs the reference's host code"
    raw = synth.make_volume(n, q.ascansPerBscan, q.bscansPerBuffer, q.bitDepth, resample=q.resampleCurve if q.resampling else None,
                            dispersion=q.dispersionCurve if q.dispersionCompensation else None)
    h1 = np.ascontiguousarray(raw).copy(); h2 = h1.copy()
    rc.init(h1, h2)
    if extras["pp_background"] is not None:
        rc.L.refcuda_set_postprocess_background(extras["pp_background"].ctypes.data, n // 2)
    rc.process(h1)
    ref = rc.output(0)
    ml = rc.mean_line() if q.fixedPatternNoiseRemoval else None
    rc.cleanup()
    floor = 4e-6 * float(np.abs(ml).max()) if ml is not None else 0.0
    for name, mode in MODES.items():
        out = gpu_run(q, raw, mode, ml, extras)
        assert_parity(out, ref, q, max_frac_outside=1e-4, saturated=bool(q.postProcessBackgroundRemoval), atol_abs=floor,
                      what=f"random #{i} N={n} {name} vs reference CUDA: {describe(q)}")
