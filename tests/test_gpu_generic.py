"""The shared-memory fused kernel (octproz_b200/csrc/k_generic.cu): OCTB200_FFT_FUSED for the line lengths and containers the
register kernels do not take -- any even N <= 8192 with prime factors <= 13 (the reference's default geometry is N = 1664 = 2^7 * 13,
octproz/default/settings.ini:62), u8 / u16 / u32 containers.  Checked against the oracle, against the cuFFT chain of the same library
and against the live reference CUDA build.  Tolerance: tests/util.py."""
import copy

import numpy as np
import pytest

from octproz_b200 import OctPipeline, _lib, benchmark_params, synth
from oracle import oracle as orc
from tests.util import assert_parity

pytestmark = pytest.mark.gpu


def run(q, raw, mode, mean_line=None, pp_background=None):
    q = copy.deepcopy(q)
    p = OctPipeline(fft_mode=mode)
    assert p.initializeCuda(None, None, q), getattr(p, "_create_error", "")
    eff = p.fft_mode
    if mean_line is not None:
        p.set_fpn_mean_line(np.asarray(mean_line, np.float32))
    if pp_background is not None:
        q.loadPostProcessingBackground(pp_background)
    l0 = p.launch_count()
    p.octCudaPipeline(np.ascontiguousarray(raw)); p.sync()
    launches = p.launch_count() - l0
    out = p.copy_output(0)
    ml = p.fpn_mean_line()
    p.cleanupCuda()
    return out, ml, eff, launches


# 2^a * {1, 3, 5, 7, 11, 13} mixes, the smallest and the largest supported length
LENGTHS = [8, 48, 100, 512, 640, 896, 1408, 1536, 1664, 2560, 3072, 4096, 4160, 8190, 8192]


@pytest.mark.parametrize("n", LENGTHS)
def test_line_lengths_match_the_oracle_in_one_launch(n):
    a, b = (12, 3) if n < 4000 else (4, 3)       # (the oracle's transform is O(N^2) for lengths that are not powers of two)
    q = benchmark_params(n, a, b, 12); q.fixedPatternNoiseRemoval = False; q.bscanFlip = True
    q.update_all_curves()
    raw = synth.make_volume(n, a, b, 12, resample=q.resampleCurve, dispersion=q.dispersionCurve)
    ref, _, _ = orc.process(q, raw)
    out, _, eff, launches = run(q, raw, _lib.FFT_FUSED)
    assert eff == _lib.FFT_FUSED
    # the first call also builds the phasor table (one fill_phase launch); the chain itself is one kernel
    assert launches <= 2, launches
    assert_parity(out, ref, q, max_frac_outside=1e-4, what=f"generic fused kernel N={n}")
    cu, _, eff_cu, _ = run(q, raw, _lib.FFT_CUFFT)
    assert eff_cu == _lib.FFT_CUFFT
    # two fp32 transforms that each meet 1e-4 against the oracle may differ from one another by twice that
    assert_parity(out, cu, q, rtol=2e-4, atol_frac=2e-4, max_frac_outside=1e-4, what=f"generic fused kernel vs cuFFT chain N={n}")


VARIANTS = {
    "linear": dict(resamplingInterpolation=0), "lanczos": dict(resamplingInterpolation=2), "noresample": dict(resampling=False),
    "fft_only": dict(resampling=False, windowing=False, dispersionCompensation=False), "klin_only": dict(windowing=False, dispersionCompensation=False),
    "rolling64": dict(backgroundRemoval=True, rollingAverageWindowSize=64), "rolling8_lanczos": dict(backgroundRemoval=True, rollingAverageWindowSize=8, resamplingInterpolation=2),
    "flip_sinus": dict(bscanFlip=True, sinusoidalScanCorrection=True), "linscale": dict(signalLogScaling=False, signalGrayscaleMin=0.0, signalGrayscaleMax=400.0),
    "fpn": dict(fixedPatternNoiseRemoval=True), "bitshift16": dict(bitshift=True, bitDepth=16),
}


@pytest.mark.parametrize("name", list(VARIANTS))
@pytest.mark.parametrize("n", [1664, 512])
def test_stage_variants_at_the_default_geometry(n, name):
    a, b = 18, 3
    q = benchmark_params(n, a, b, 12); q.fixedPatternNoiseRemoval = False
    for k, v in VARIANTS[name].items():
        setattr(q, k, v)
    q.update_all_curves()
    raw = synth.make_volume(n, a, b, q.bitDepth, resample=q.resampleCurve, dispersion=q.dispersionCurve)
    ref, ml, _ = orc.process(q, raw)
    lanczos = q.resampling and q.resamplingInterpolation == 2
    # FPN: the oracle's line injected (the determination itself: test_own_determination_* below and tests/test_gpu_reference_full.py)
    out, _, eff, _ = run(q, raw, _lib.FFT_FUSED, mean_line=ml if q.fixedPatternNoiseRemoval else None)
    assert eff == _lib.FFT_FUSED
    assert_parity(out, ref, q, atol_frac=2e-2 if lanczos else 1e-4, max_frac_outside=1e-4, what=f"generic N={n} {name}",
                  atol_abs=4e-6 * float(np.abs(ml).max()) if q.fixedPatternNoiseRemoval else 0.0)


@pytest.mark.parametrize("bits,n", [(8, 1024), (8, 2048), (8, 1664), (32, 1024), (32, 2048), (32, 1664), (24, 512), (20, 1024)])
def test_u8_and_u32_containers_take_a_fused_path(bits, n):
    """round 1 sent u8 / u32 containers through the three-kernel chain; they are one launch now: N = 1024 / 2048 through the register
    kernel's SRC_RAW8 / SRC_RAW32 slot conversions, other lengths through the shared-memory kernel"""
    q = benchmark_params(n, 8, 2, bits); q.fixedPatternNoiseRemoval = False; q.update_all_curves()
    raw = synth.make_volume(n, 8, 2, min(bits, 20), resample=q.resampleCurve, dispersion=q.dispersionCurve).astype(synth.container_dtype(bits))
    ref, _, _ = orc.process(q, raw)
    out, _, eff, launches = run(q, raw, _lib.FFT_FUSED)
    assert eff == _lib.FFT_FUSED and launches <= 2
    assert_parity(out, ref, q, max_frac_outside=1e-4, what=f"fused, container of {bits} bits, N={n}")
    if bits > 16:
        q.bitshift = True
        ref, _, _ = orc.process(q, raw)
        out, _, _, _ = run(q, raw, _lib.FFT_FUSED)
        assert_parity(out, ref, q, max_frac_outside=1e-4, what=f"fused, u32 bitshift, N={n}")


CONTAINER_VARIANTS = {
    "linear": dict(resamplingInterpolation=0), "noresample": dict(resampling=False), "fft_only": dict(resampling=False, windowing=False, dispersionCompensation=False),
    "flip_sinus": dict(bscanFlip=True, sinusoidalScanCorrection=True), "linscale": dict(signalLogScaling=False, signalGrayscaleMin=0.0, signalGrayscaleMax=40.0),
    "bitshift": dict(bitshift=True), "fpn": dict(fixedPatternNoiseRemoval=True),
    # the container kernels cover the 4-tap and plain stages; these two take the split chain (pre kernel + register FFT kernel) and must still be right
    "lanczos": dict(resamplingInterpolation=2), "rolling16": dict(backgroundRemoval=True, rollingAverageWindowSize=16),
}


@pytest.mark.parametrize("name", list(CONTAINER_VARIANTS))
@pytest.mark.parametrize("bits,n", [(8, 1024), (8, 2048), (32, 1024), (32, 2048)])
def test_container_kernels_stage_variants(bits, n, name):
    a, b = 20, 3
    q = benchmark_params(n, a, b, bits); q.fixedPatternNoiseRemoval = False
    for k, v in CONTAINER_VARIANTS[name].items():
        setattr(q, k, v)
    q.update_all_curves()
    raw = synth.make_volume(n, a, b, min(bits, 20), resample=q.resampleCurve, dispersion=q.dispersionCurve).astype(synth.container_dtype(bits))
    ref, ml, _ = orc.process(q, raw)
    out, _, eff, launches = run(q, raw, _lib.FFT_AUTO, mean_line=ml if q.fixedPatternNoiseRemoval else None)
    assert eff == _lib.FFT_FUSED
    lanczos = q.resampling and q.resamplingInterpolation == 2
    assert_parity(out, ref, q, atol_frac=2e-2 if lanczos else 1e-4, max_frac_outside=1e-4, what=f"container of {bits} bits, N={n}, {name}",
                  atol_abs=4e-6 * float(np.abs(ml).max()) if q.fixedPatternNoiseRemoval else 0.0)
    # bit-identical to the u16 kernel on the same values wherever those fit a u16 container
    if bits == 8 and name not in ("lanczos", "rolling16"):
        q16 = copy.deepcopy(q); q16.bitDepth = 16
        out16, _, _, _ = run(q16, raw.astype(np.uint16), _lib.FFT_FUSED, mean_line=ml if q.fixedPatternNoiseRemoval else None)
        assert np.array_equal(out, out16), "u8 container kernel differs from the u16 kernel on the same sample values"


def test_auto_mode_policy():
    """AUTO: register kernels for 1024 / 2048 (every container); the shared-memory kernel where it beats the cuFFT chain (not a power of
    two, or N > 2048); the cuFFT chain for short power-of-two lines and for lengths with large prime factors"""
    want = {(1024, 12): _lib.FFT_FUSED, (2048, 8): _lib.FFT_FUSED, (1024, 32): _lib.FFT_FUSED, (1664, 12): _lib.FFT_FUSED, (4096, 12): _lib.FFT_FUSED,
            (1536, 8): _lib.FFT_FUSED, (512, 12): _lib.FFT_CUFFT, (256, 12): _lib.FFT_CUFFT, (1006, 12): _lib.FFT_CUFFT}
    for (n, bits), mode in want.items():
        q = benchmark_params(n, 4, 2, bits); q.update_all_curves()
        p = OctPipeline(fft_mode=_lib.FFT_AUTO)
        assert p.initializeCuda(None, None, q), getattr(p, "_create_error", "")
        assert p.fft_mode == mode, (n, bits, p.fft_mode)
        p.cleanupCuda()


def test_lengths_with_large_prime_factors_keep_the_cufft_chain():
    for n in (1006, 34, 8232):              # 2 * 503, 2 * 17, 2^3 * 3 * 7^3 = above the shared-memory limit
        q = benchmark_params(n, 4, 2, 12); q.fixedPatternNoiseRemoval = False; q.update_all_curves()
        raw = synth.make_volume(n, 4, 2, 12, resample=q.resampleCurve, dispersion=q.dispersionCurve)
        ref, _, _ = orc.process(q, raw)
        out, _, eff, _ = run(q, raw, _lib.FFT_AUTO)
        assert eff == _lib.FFT_CUFFT
        assert_parity(out, ref, q, max_frac_outside=1e-4, what=f"cuFFT chain N={n}")
        p = OctPipeline(fft_mode=_lib.FFT_FUSED)
        assert not p.initializeCuda(None, None, copy.deepcopy(q))          # asked for explicitly: a loud refusal, no silent fallback


@pytest.mark.skipif(not orc.have_ref("libref_cuda.so"), reason="oracle/_ref/libref_cuda.so not built")
@pytest.mark.parametrize("shape", [(1664, 512, 16, 12), (512, 256, 8, 12), (4096, 128, 4, 16)])
def test_matches_the_live_reference_cuda_build(shape):
    """the reference CUDA path (cuFFT plan of any length, cuda_code.cu:1140) and ours on the same raw buffer, benchmark settings incl. FPN"""
    from tests.test_gpu_reference_full import classify_fpn_bins, reference_run
    n, a, b, bits = shape
    q = benchmark_params(n, a, b, bits)
    q.update_all_curves()
    raw = synth.make_volume(n, a, b, bits, resample=q.resampleCurve, dispersion=q.dispersionCurve)
    ref, ref_ml = reference_run(q, raw)
    out, _, eff, _ = run(q, raw, _lib.FFT_FUSED, mean_line=ref_ml)
    assert eff == _lib.FFT_FUSED
    assert_parity(out, ref, q, max_frac_outside=1e-4, what=f"generic fused kernel vs live reference {shape}", atol_abs=4e-6 * float(np.abs(ref_ml).max()))
    # own determination through the generic kernel's complex-output pass
    qq = copy.deepcopy(q)
    p = OctPipeline(fft_mode=_lib.FFT_FUSED)
    assert p.initializeCuda(None, None, qq)
    p.octCudaPipeline(np.ascontiguousarray(raw)); p.sync()
    own = p.copy_output(0); ml = p.fpn_mean_line(); stats, seg_len = p.fpn_segment_stats()
    p.cleanupCuda()
    c = classify_fpn_bins(stats, seg_len, ml, ref_ml)
    assert c["identification_error"] < 5e-4 and np.all(c["gap_over_bound"] <= 1.0), (c["differ"], c["gap_over_bound"])
    assert_parity(own[..., c["same"]], ref[..., c["same"]], q, max_frac_outside=1e-4, what=f"generic fused kernel, own FPN line, vs live reference {shape}",
                  atol_abs=4e-6 * float(np.abs(ref_ml).max()))


def test_full_size_default_geometry_properties():
    """1664 x 512 x 256 (the default settings' line length at the BASELINE volume shape): determinism, flip as an exact permutation,
    agreement with the cuFFT chain"""
    n, a, b = 1664, 512, 256
    q = benchmark_params(n, a, b, 12); q.fixedPatternNoiseRemoval = False; q.update_all_curves()
    small = synth.make_volume(n, a, 8, 12, resample=q.resampleCurve, dispersion=q.dispersionCurve)
    raw = np.ascontiguousarray(np.tile(small, (b // 8, 1, 1)))
    out, _, eff, _ = run(q, raw, _lib.FFT_FUSED)
    assert eff == _lib.FFT_FUSED
    again, _, _, _ = run(q, raw, _lib.FFT_FUSED)
    assert np.array_equal(out, again), "not deterministic"
    assert np.array_equal(out[:8], out[8:16]) and np.array_equal(out[:8], out[-8:]), "periodic input must give periodic output"
    ref, _, _ = orc.process(benchmark_params_like(q, 1), small[:1])        # (the oracle's transform is O(N^2) for this length: one B-scan)
    assert_parity(out[:1], ref, q, max_frac_outside=1e-4, what="generic fused kernel, full size 1664x512x256, first B-scan vs oracle")
    qf = copy.deepcopy(q); qf.bscanFlip = True
    fl, _, _, _ = run(qf, raw, _lib.FFT_FUSED)
    assert np.array_equal(fl[0::2], out[0::2, ::-1]) and np.array_equal(fl[1::2], out[1::2])
    cu, _, _, _ = run(q, raw, _lib.FFT_CUFFT)
    assert_parity(out, cu, q, max_frac_outside=1e-5, what="generic fused kernel vs cuFFT chain, full size 1664x512x256")


def benchmark_params_like(q, bscans):
    qq = copy.deepcopy(q); qq.bscansPerBuffer = bscans
    return qq
