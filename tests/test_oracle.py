"""The oracle (oracle/oct_oracle.c) against known answers, numpy restatements of single stages, and -- when the
fixtures exist -- golden vectors produced by the reference's unmodified CUDA file on a B200
(tests/golden/refcuda_*.npz, tests/golden/make_golden_refcuda.py)."""
import copy
import glob
import os

import numpy as np
import pytest

from octproz_b200 import OctAlgorithmParameters, benchmark_params, synth
from oracle import oracle as orc
from tests.golden.cases import chain_cases
from tests.util import assert_parity

GOLD_DIR = os.path.join(os.path.dirname(__file__), "golden")


def plain(n, a, b, bits=12):
    q = OctAlgorithmParameters(samplesPerLine=n, ascansPerBscan=a, bscansPerBuffer=b, bitDepth=bits)
    q.signalLogScaling = False; q.signalGrayscaleMin = 0.0; q.signalGrayscaleMax = 1.0
    return q


def test_fft_only_matches_numpy_unnormalised_inverse():
    n, a, b = 1024, 4, 2
    q = plain(n, a, b)
    raw = synth.make_volume(n, a, b, 12)
    out, _, cplx = orc.process(q, raw, want_complex=True)
    X = np.fft.ifft(raw.astype(np.float64), axis=-1) * n            # cufftExecC2C INVERSE is unnormalised (cuda_code.cu:1515)
    got = cplx[..., 0] + 1j * cplx[..., 1]
    assert np.allclose(got, X, rtol=1e-12, atol=1e-6)
    assert np.allclose(out, np.abs(X[..., : n // 2]) / (n / 2), rtol=1e-6)     # cuda_code.cu:739, min 0 max 1


def test_non_power_of_two_line_length():
    n, a, b = 100, 3, 1
    q = plain(n, a, b)
    raw = synth.make_volume(n, a, b, 12)
    out, _, _ = orc.process(q, raw)
    X = np.fft.ifft(raw.astype(np.float64), axis=-1) * n
    assert np.allclose(out, np.abs(X[..., : n // 2]) / (n / 2), rtol=1e-6)


def test_impulse_known_answer_log_scale():
    n = 1024
    q = plain(n, 1, 1); q.signalLogScaling = True; q.signalGrayscaleMin = -30.0; q.signalGrayscaleMax = 100.0
    raw = np.zeros((1, 1, n), np.uint16); raw[0, 0, 0] = 1000
    out, _, _ = orc.process(q, raw)
    expect = ((10.0 * np.log10(1000.0 ** 2 / (n / 2))) + 30.0) / 130.0      # flat spectrum of an impulse, cuda_code.cu:718
    assert np.allclose(out, expect, rtol=0, atol=1e-6)


def test_all_zero_input_is_minus_infinity_in_log_mode():
    q = plain(1024, 1, 1); q.signalLogScaling = True
    out, _, _ = orc.process(q, np.zeros((1, 1, 1024), np.uint16))
    assert np.all(np.isneginf(out))        # no clamping, cuda_code.cu:718


def test_convert_bitshift_and_containers():
    n = 64
    for bits, dt in ((8, np.uint8), (16, np.uint16), (32, np.uint32)):
        q = plain(n, 1, 1, bits)
        rng = np.random.default_rng(bits)
        raw = rng.integers(0, np.iinfo(dt).max, size=(1, 1, n), dtype=dt)
        for shift in (False, True):
            q.bitshift = shift
            out, _, cplx = orc.process(q, raw, want_complex=True)
            x = raw.astype(np.float64)
            x = (x / 4294967296.0) if (shift and bits == 32) else (np.floor(x / 16) if shift else x)    # cuda_code.cu:138-144
            X = np.fft.ifft(x, axis=-1) * n
            assert np.allclose(cplx[..., 0] + 1j * cplx[..., 1], X, rtol=1e-10, atol=1e-9 * max(1.0, x.max()) * n)


def test_rolling_background_clips_window_to_the_line():
    n, w = 64, 5
    q = plain(n, 2, 1); q.backgroundRemoval = True; q.rollingAverageWindowSize = w
    raw = synth.make_volume(n, 2, 1, 12)
    _, _, cplx = orc.process(q, raw, want_complex=True)
    x = raw.astype(np.float64)
    y = np.empty_like(x)
    for m in range(n):
        s, e = max(0, m - w + 1), min(n - 1, m + w)            # cuda_code.cu:183-185
        y[..., m] = x[..., m] - x[..., s:e + 1].mean(axis=-1)
    X = np.fft.ifft(y, axis=-1) * n
    assert np.allclose(cplx[..., 0] + 1j * cplx[..., 1], X, rtol=1e-9, atol=1e-6)


@pytest.mark.parametrize("interp", [0, 1])
def test_interpolators_reproduce_a_linear_ramp(interp):
    n = 256
    q = plain(n, 1, 1); q.resampling = True; q.resamplingInterpolation = interp
    q.c0, q.c1, q.c2, q.c3 = 0.7, 200.0, 30.0, -10.0
    q.updateResampleCurve()
    raw = (10 + 3 * np.arange(n)).astype(np.uint16).reshape(1, 1, n)
    _, _, cplx = orc.process(q, raw, want_complex=True)
    y = np.fft.fft(cplx[0, 0, :, 0] + 1j * cplx[0, 0, :, 1]).real / n
    r = q.resampleCurve.astype(np.float64)
    inner = r >= 1.0                       # n0 = |n1-1| mirrors at n1 = 0 (cuda_code.cu:284), exact only for n1 >= 1
    assert np.allclose(y[inner], 10 + 3 * r[inner], rtol=0, atol=1e-6)


def test_cubic_mirror_tap_at_the_line_start():
    n = 64
    q = plain(n, 1, 1); q.resampling = True; q.resamplingInterpolation = 1
    q.resampleCurve = np.full(n, 0.25, np.float32)
    x = np.arange(n, dtype=np.float64) ** 2 + 5
    _, _, cplx = orc.process(q, x.astype(np.uint16).reshape(1, 1, n), want_complex=True)
    y = np.fft.fft(cplx[0, 0, :, 0] + 1j * cplx[0, 0, :, 1]).real / n
    y0, y1, y2, y3, t = x[1], x[0], x[1], x[2], 0.25          # y0 read at abs(0-1) = 1
    a = -y0 + 3 * (y1 - y2) + y3; b = 2 * y0 - 5 * y1 + 4 * y2 - y3; c = -y0 + y2
    assert np.allclose(y, 0.5 * t * (a * t * t + b * t + c) + y1, atol=1e-9)


def test_lanczos_offset_clamp_shifts_the_first_line_by_eight():
    n, a = 128, 3
    q = plain(n, a, 1); q.resampling = True; q.resamplingInterpolation = 2
    q.resampleCurve = np.clip(np.arange(n, dtype=np.float32), 0, n - 3)     # integer positions: kernel = delta
    raw = synth.make_volume(n, a, 1, 12)
    _, _, cplx = orc.process(q, raw, want_complex=True)
    y = np.fft.fft(cplx[0, :, :, 0] + 1j * cplx[0, :, :, 1], axis=-1).real / n
    flat = raw.reshape(-1).astype(np.float64)
    m = np.arange(n - 16)
    assert np.allclose(y[0, m], flat[8 + m], atol=1e-3)                      # cuda_code.cu:313: offset = max(offset, 8)
    assert np.allclose(y[1, m], flat[n + m], atol=1e-3)


def test_fpn_minimum_variance_segment_and_half_line_subtraction():
    n, a, b = 64, 18, 2
    q = plain(n, a, b); q.fixedPatternNoiseRemoval = True; q.bscansForNoiseDetermination = 2
    raw = synth.make_volume(n, a, b, 12)
    out, ml, cplx = orc.process(q, raw, want_complex=True)
    X = (cplx[..., 0] + 1j * cplx[..., 1]).reshape(a * b, n)
    L = (2 * a) // 9                                                          # integer division, cuda_code.cu:531
    best = np.zeros(n, complex)
    for z in range(n):
        mv = np.finfo(np.float32).max
        for s in range(9):
            seg = X[s * L:(s + 1) * L, z]
            mu = seg.mean(); var = (np.abs(seg) ** 2).mean() - abs(mu) ** 2
            if var < mv:
                mv, best[z] = var, mu
    assert np.allclose(ml[:, 0] + 1j * ml[:, 1], best, rtol=1e-9, atol=1e-9)
    Y = X.copy(); Y[:, : n // 2] -= best[None, : n // 2]                      # only the first N/2 bins, cuda_code.cu:567-584
    assert np.allclose(out.reshape(a * b, n // 2), np.abs(Y[:, : n // 2]) / (n / 2), rtol=1e-5, atol=1e-7)


def emulate_reference_bscan_flip(vol):
    """cuda_bscanFlip thread by thread (cuda_code.cu:787-807, launched at :1547 with halfSamplesInVolume = samplesPerBuffer/4):
    in-place swaps, every (index, mirrorIndex) pair is touched by exactly one thread"""
    b, a, h = vol.shape
    out = vol.copy().reshape(-1)
    samples_per_bscan = h * a
    for index0 in range((2 * h * a * b) // 4):
        bscan = (index0 // samples_per_bscan) * 2
        index = bscan * samples_per_bscan + index0 % samples_per_bscan
        sample = index % samples_per_bscan
        ascan = sample // h
        mirror = bscan * samples_per_bscan + ((a - 1) - ascan) * h + sample % h
        if ascan >= a // 2:
            out[mirror], out[index] = out[index], out[mirror]
    return out.reshape(b, a, h)


def test_flip_even_bscans_and_sinusoidal_last_line():
    n, a, b = 64, 6, 3
    q = plain(n, a, b)
    raw = synth.make_volume(n, a, b, 12)
    base, _, _ = orc.process(q, raw)
    q.bscanFlip = True
    flipped, _, _ = orc.process(q, raw)
    # odd number of B-scans: the reference's kernel covers samplesPerBuffer/4 elements, so the LAST even B-scan stays as it is
    assert np.array_equal(flipped[0], base[0, ::-1]) and np.array_equal(flipped[1], base[1]) and np.array_equal(flipped[2], base[2])
    for bb, aa in ((1, 6), (3, 6), (4, 6), (5, 5), (3, 5), (2, 7)):
        qq = plain(n, aa, bb); rr = synth.make_volume(n, aa, bb, 12)
        plain_out, _, _ = orc.process(qq, rr)
        qq.bscanFlip = True
        got, _, _ = orc.process(qq, rr)
        assert np.array_equal(got, emulate_reference_bscan_flip(plain_out)), (bb, aa)
    q.sinusoidalScanCorrection = True
    sc, _, _ = orc.process(q, raw)
    curve = orc.sinusoidal_curve(a).astype(np.float64)
    T = flipped.astype(np.float64)
    for k in range(a):
        qf = int(curve[k]); fr = curve[k] - qf
        expect = T[:, qf] + (T[:, min(qf + 1, a - 1)] - T[:, qf]) * fr
        if k == a - 1:
            assert np.allclose(sc[:b - 1, k], expect[:b - 1], rtol=1e-6, atol=1e-7)
            assert np.array_equal(sc[b - 1, k], flipped[b - 1, k])            # last line untouched (cuda_code.cu:499)
        else:
            assert np.allclose(sc[:, k], expect, rtol=1e-6, atol=1e-7)


def test_postprocess_background_and_saturation():
    n, a, b = 64, 4, 2
    q = plain(n, a, b); q.signalGrayscaleMax = 400.0
    raw = synth.make_volume(n, a, b, 12)
    base, _, _ = orc.process(q, raw)
    bg = orc.postprocess_background(base, n // 2, a)
    assert np.allclose(bg, base[0].mean(axis=0), rtol=1e-6)                   # first B-scan only, cuda_code.cu:743-755
    q.postProcessBackgroundRemoval = True; q.postProcessBackgroundWeight = 0.5; q.postProcessBackgroundOffset = 0.01
    out, _, _ = orc.process(q, raw, pp_background=bg)
    assert np.allclose(out, np.clip(base - (0.5 * bg + 0.01), 0, 1), atol=1e-7)


def test_display_extraction_and_output_conversion():
    h, a, btot = 8, 5, 4
    vol = np.random.default_rng(3).random((btot, a, h), dtype=np.float32)
    F = h * a
    f2 = orc.bscan_frame(vol, h, a, btot, 2, 1, 0)
    assert np.array_equal(f2, vol[2].reshape(-1)[::-1])                       # disp[i] = vol[f*F + F-1-i], cuda_code.cu:857
    assert np.allclose(orc.bscan_frame(vol, h, a, btot, 2, 5, 0), vol[2:4].mean(axis=0).reshape(-1)[::-1], rtol=1e-6)   # clipped to the volume
    assert np.array_equal(orc.bscan_frame(vol, h, a, btot, 1, 2, 1), vol[1:3].max(axis=0).reshape(-1)[::-1])
    assert np.array_equal(orc.bscan_frame(vol, h, a, btot, 99, 1, 0), vol[0].reshape(-1)[::-1])                          # frameNr >= depth -> 0
    e = orc.enface_frame(vol, h, a, btot, 3, 1, 0)
    assert np.array_equal(e, vol[:, :, 3].reshape(-1)[::-1])                  # cuda_code.cu:909
    assert np.allclose(orc.enface_frame(vol, h, a, btot, 6, 4, 0), vol[:, :, 6:8].mean(axis=2).reshape(-1)[::-1], rtol=1e-6)
    x = np.array([-0.5, 0.0, 0.25, 0.999, 1.0, 7.0, np.nan], np.float32)
    assert list(orc.float_to_output(x, 8)) == [0, 0, 63, 254, 255, 255, 0]    # saturate, truncating cast, NaN -> 0
    assert list(orc.float_to_output(x, 12)) == [0, 0, 1023, 4090, 4095, 4095, 0]
    assert list(orc.float_to_output(x, 16)) == [0, 0, 16383, 65469, 65535, 65535, 0]
    assert orc.float_to_output(x, 32)[4] == 0xFFFFFFFF


def test_fp32_mode_tracks_fp64_mode():
    q = benchmark_params(1024, 16, 2); q.fixedPatternNoiseRemoval = False; q.update_all_curves()
    raw = synth.make_volume(1024, 16, 2, 12, resample=q.resampleCurve, dispersion=q.dispersionCurve)
    o64, _, _ = orc.process(q, raw, precision=64)
    o32, _, _ = orc.process(q, raw, precision=32)
    assert_parity(o32, o64, q, what="oracle fp32 vs fp64")


GOLDEN = sorted(glob.glob(os.path.join(GOLD_DIR, "refcuda_*.npz")))


@pytest.mark.skipif(not GOLDEN, reason="reference-CUDA golden vectors not generated yet (tests/golden/make_golden_refcuda.py needs a GPU)")
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[8:-4] for p in GOLDEN])
def test_oracle_against_reference_cuda_golden_vectors(path):
    """pins the oracle to outputs of the reference itself (unmodified cuda_code.cu, sm_100, --use_fast_math)"""
    name = os.path.basename(path)[8:-4]
    n = int(name.split("_")[0][1:]); case = name.split("_", 1)[1]
    g = np.load(path)
    q = copy.deepcopy(chain_cases(n)[case])
    q.resampleCurve, q.dispersionCurve, q.windowCurve = g["resample"], g["dispersion"], g["window"]
    kw = {}
    if "pp_background" in g.files:
        kw["pp_background"] = g["pp_background"]
    if "mean_line" in g.files:
        # the reference's single-pass fp32 variance is ill-conditioned at DC-dominated bins: compare the
        # chain with the reference's own mean line, and the determination separately where it is well conditioned
        out, ml, _ = orc.process(q, g["raw"], mean_line=g["mean_line"].astype(np.float64), determine_fpn=False, **kw)
        _, ml_own, _ = orc.process(q, g["raw"], **kw)
        h = n // 2
        same = np.isclose(ml_own[:h], g["mean_line"][:h], rtol=1e-3, atol=1e-2 * np.abs(g["mean_line"][:h]).max() * 1e-3).all(axis=1)
        assert same.mean() > 0.9, f"min-variance mean line agrees on only {same.mean():.2%} of the bins"
    else:
        out, _, _ = orc.process(q, g["raw"], **kw)
    # FPN: X - M cancels a term ~1e3 x larger than the result at the fixed-pattern / DC bins; the reference's fp32 round-off
    # of that term (~eps32 |M|) is the floor there
    floor = 4e-6 * float(np.abs(g["mean_line"]).max()) if "mean_line" in g.files else 0.0
    lanczos = case == "lanczos"          # the reference's Lanczos weights come from __sinf (fast-math); error floor ~1e-3 of the median amplitude
    assert_parity(out, g["out"], q, atol_frac=2e-2 if lanczos else 1e-4, saturated=bool(q.postProcessBackgroundRemoval),
                  max_frac_outside=1e-4, what=name, atol_abs=floor)


def test_float_to_output_known_answers_and_special_values():
    """floatToOutput (cuda_code.cu:943-967): (container)((double)saturate(x) * (2^bits - 1)), i.e. truncation of the exact product"""
    below_one = np.nextafter(np.float32(1.0), np.float32(0.0))
    x = np.array([0.0, -0.0, 1.0, below_one, 0.5, -3.0, 7.0, np.nan, np.inf, -np.inf, 1.0 / 4095.0, 2047.5 / 4095.0], np.float32)
    got = orc.float_to_output(x, 12)
    third = int(np.floor(float(np.float32(1.0 / 4095.0)) * 4095.0))        # 1/4095 is not representable: 0 or 1 depending on its rounding
    half = int(np.floor(float(np.float32(2047.5 / 4095.0)) * 4095.0))
    assert got.dtype == np.uint16
    assert got.tolist() == [0, 0, 4095, 4094, 2047, 0, 4095, 0, 4095, 0, third, half]
    assert orc.float_to_output(np.array([1.0, below_one, 0.5], np.float32), 8).tolist() == [255, 254, 127]
    assert orc.float_to_output(np.array([1.0, below_one, 0.25], np.float32), 10).tolist() == [1023, 1022, 255]
    assert orc.float_to_output(np.array([1.0, below_one, 0.25], np.float32), 16).tolist() == [65535, 65534, 16383]
    assert orc.float_to_output(np.array([1.0, 0.5], np.float32), 24).dtype == np.uint32


def test_round_toward_zero_fma_identity_behind_the_fused_conversion():
    """The fused kernel converts with ONE fp32 FMA rounded toward zero: low 16 bits of RZ(sat(x) * K + 2^23) (oct_tmem.cuh).  Checked here in
    exact rational arithmetic: for every fp32 s in [0, 1] and K in {1023, 4095, 65535}, RZ32(s * K + 2^23) - 2^23 == floor(s * K), which is
    what the reference's double-precision product truncates to.  (The GPU tests check the kernel itself bit for bit.)"""
    from fractions import Fraction

    def rz32(v: Fraction) -> Fraction:          # round a positive rational toward zero onto the fp32 grid
        e = v.numerator.bit_length() - v.denominator.bit_length()
        if Fraction(2) ** e > v:
            e -= 1
        ulp = Fraction(2) ** (e - 23)
        return (v // ulp) * ulp

    rng = np.random.default_rng(5)
    samples = np.concatenate([rng.random(3000, dtype=np.float32), np.float32(1.0) - rng.random(500, dtype=np.float32) * np.float32(1e-4),
                              rng.random(500, dtype=np.float32) * np.float32(1e-3),
                              np.array([0.0, 1.0, np.nextafter(np.float32(1), np.float32(0)), 0.5, 1e-30, 1e-45], np.float32)])
    for K in (1023, 4095, 65535):
        ks = np.arange(1, K + 1, max(1, K // 257), dtype=np.float64) / K                 # values right at the integer boundaries k / K ...
        edge = np.concatenate([ks.astype(np.float32), np.nextafter(ks.astype(np.float32), np.float32(0)), np.nextafter(ks.astype(np.float32), np.float32(2))])
        for s in np.concatenate([samples, edge[edge <= 1.0]]):
            exact = Fraction(float(s)) * K
            fused = rz32(exact + 2 ** 23) - 2 ** 23
            assert fused == exact.numerator // exact.denominator, (float(s), K)
        want = orc.float_to_output(samples, {1023: 10, 4095: 12, 65535: 16}[K])
        mine = np.array([int(Fraction(float(s)) * K) for s in samples], np.uint16)
        assert np.array_equal(want, mine)
