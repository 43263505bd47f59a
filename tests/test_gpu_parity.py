"""Parity of the CUDA path (through the C ABI) against the oracle, the reference-CUDA golden vectors, the live
reference CUDA build (when oracle/_ref/libref_cuda.so travelled to the box) and size-independent properties at
the full BASELINE size.  Tolerance: tests/util.py (1e-4 relative on the amplitude + 1e-4 of the median amplitude)."""
import copy
import ctypes as C
import glob
import os

import numpy as np
import pytest

from octproz_b200 import OctAlgorithmParameters, OctPipeline, _lib, benchmark_params, synth
from oracle import oracle as orc
from tests.golden.cases import chain_cases
from tests.util import assert_parity, parity_report

pytestmark = pytest.mark.gpu
GOLD_DIR = os.path.join(os.path.dirname(__file__), "golden")
MODES = {"fused": _lib.FFT_FUSED, "split": _lib.FFT_SPLIT, "cufft": _lib.FFT_CUFFT}


def run(q, raw, mode=_lib.FFT_AUTO, mean_line=None, pp_background=None, calls=1):
    q = copy.deepcopy(q)
    p = OctPipeline(fft_mode=mode)
    assert p.initializeCuda(None, None, q), getattr(p, "_create_error", "")
    if mean_line is not None:
        p.set_fpn_mean_line(np.asarray(mean_line, np.float32))
    if pp_background is not None:
        q.loadPostProcessingBackground(pp_background)
    h = np.ascontiguousarray(raw)
    for _ in range(calls):
        p.octCudaPipeline(h)
    p.sync()
    out = p.copy_output(0)
    ml = p.fpn_mean_line()
    p.cleanupCuda()
    return out, ml


def variants(n, a=32, b=4):
    v = {}
    base = benchmark_params(n, a, b); base.fixedPatternNoiseRemoval = False
    v["benchmark_nofpn"] = base
    for name, kw in {
        "linear": dict(resamplingInterpolation=0), "lanczos": dict(resamplingInterpolation=2), "noresample": dict(resampling=False),
        "klin_only": dict(windowing=False, dispersionCompensation=False), "klin_disp": dict(windowing=False),
        "fft_only": dict(resampling=False, windowing=False, dispersionCompensation=False),
        "win_only": dict(resampling=False, dispersionCompensation=False), "disp_only": dict(resampling=False, windowing=False),
        "rolling64": dict(backgroundRemoval=True, rollingAverageWindowSize=64),
        "rolling1": dict(backgroundRemoval=True, rollingAverageWindowSize=1),
        "rolling8_lanczos": dict(backgroundRemoval=True, rollingAverageWindowSize=8, resamplingInterpolation=2),
        "flip": dict(bscanFlip=True), "flip_sinus": dict(bscanFlip=True, sinusoidalScanCorrection=True),
        "linscale": dict(signalLogScaling=False, signalGrayscaleMin=0.0, signalGrayscaleMax=400.0),
        "gauss_window": dict(window=1, windowFillFactor=0.7), "coeff_addend": dict(signalMultiplicator=0.8, signalAddend=0.1),
        "bitshift16": dict(bitshift=True, bitDepth=16),
    }.items():
        q = copy.deepcopy(base)
        for k, val in kw.items():
            setattr(q, k, val)
        v[name] = q
    return v


@pytest.mark.parametrize("mode", list(MODES))
@pytest.mark.parametrize("n", [1024, 2048])
@pytest.mark.parametrize("name", list(variants(1024)))
def test_matches_oracle(n, name, mode):
    q = variants(n)[name]; q.update_all_curves()
    raw = synth.make_volume(n, q.ascansPerBscan, q.bscansPerBuffer, q.bitDepth, resample=q.resampleCurve, dispersion=q.dispersionCurve)
    ref, _, _ = orc.process(q, raw)
    out, _ = run(q, raw, MODES[mode])
    lanczos = q.resampling and q.resamplingInterpolation == 2
    # Lanczos weights come from __sinf like the reference's (cuda_code.cu:297-302 under --use_fast_math): vs the fp64 oracle the
    # error floor is ~1e-2 of the median amplitude (N=2048: arguments up to 8 pi); against the reference CUDA build itself the 1e-4 bound holds (golden test)
    assert_parity(out, ref, q, atol_frac=2e-2 if lanczos else 1e-4, max_frac_outside=1e-4, what=f"N={n} {name} {mode}")


GOLDEN = sorted(glob.glob(os.path.join(GOLD_DIR, "refcuda_*.npz")))


@pytest.mark.skipif(not GOLDEN, reason="no reference-CUDA golden vectors committed yet")
@pytest.mark.parametrize("mode", list(MODES))
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[8:-4] for p in GOLDEN])
def test_matches_reference_cuda_golden(path, mode):
    """same raw buffer, same LUTs: our output vs the unmodified reference cuda_code.cu (sm_100, --use_fast_math)"""
    name = os.path.basename(path)[8:-4]
    n = int(name.split("_")[0][1:]); case = name.split("_", 1)[1]
    g = np.load(path)
    q = copy.deepcopy(chain_cases(n)[case])
    q.resampleCurve, q.dispersionCurve, q.windowCurve = g["resample"], g["dispersion"], g["window"]
    out, ml = run(q, g["raw"], MODES[mode], mean_line=g["mean_line"] if "mean_line" in g.files else None,
                  pp_background=g["pp_background"] if "pp_background" in g.files else None)
    floor = 4e-6 * float(np.abs(g["mean_line"]).max()) if "mean_line" in g.files else 0.0     # round-off of the cancelled FPN term
    assert_parity(out, g["out"], q, saturated=bool(q.postProcessBackgroundRemoval), max_frac_outside=1e-4, what=f"{name} {mode}", atol_abs=floor)


@pytest.mark.skipif(not orc.have_ref("libref_cuda.so"), reason="oracle/_ref/libref_cuda.so not built")
@pytest.mark.parametrize("shape", [(1024, 512, 64, 12), (2048, 256, 32, 16)])
def test_matches_live_reference_cuda_large(shape):
    """the reference CUDA path and ours on the same raw buffer, benchmark settings incl. FPN, same box"""
    n, a, b, bits = shape
    q = benchmark_params(n, a, b, bits)
    rc = orc.RefCuda(); rc.configure(q)
    q.resampleCurve, q.dispersionCurve, q.windowCurve = rc.curves()
    raw = synth.make_volume(n, a, b, bits, resample=q.resampleCurve, dispersion=q.dispersionCurve)
    h1 = np.ascontiguousarray(raw).copy(); h2 = h1.copy()
    rc.init(h1, h2); rc.process(h1)
    ref = rc.output(0); ref_ml = rc.mean_line()
    rc.cleanup()
    for mode in MODES.values():
        out, ml = run(q, raw, mode, mean_line=ref_ml)
        assert_parity(out, ref, q, max_frac_outside=1e-4, what=f"live reference {shape} mode {mode}", atol_abs=4e-6 * float(np.abs(ref_ml).max()))
    # our own FPN determination against the reference's: tests/test_gpu_reference_full.py


def test_full_size_properties():
    """BASELINE size 1024 x 512 x 256: modes agree, flip is an exact permutation, run-to-run determinism, display extraction"""
    import torch
    n, a, b = 1024, 512, 256
    q = benchmark_params(n, a, b); q.fixedPatternNoiseRemoval = False; q.update_all_curves()
    small = synth.make_volume(n, a, 8, 12, resample=q.resampleCurve, dispersion=q.dispersionCurve)
    raw = np.ascontiguousarray(np.tile(small, (b // 8, 1, 1)))
    outs = {}
    for name, mode in MODES.items():
        outs[name], _ = run(q, raw, mode)
    assert_parity(outs["split"], outs["fused"], q, max_frac_outside=1e-5, what="split vs fused")
    assert_parity(outs["cufft"], outs["fused"], q, max_frac_outside=1e-5, what="cufft vs fused")
    again, _ = run(q, raw, _lib.FFT_FUSED)
    assert np.array_equal(again, outs["fused"]), "fused kernel is not deterministic"
    # periodic input -> periodic output (tile of 8 B-scans)
    assert np.array_equal(outs["fused"][:8], outs["fused"][8:16])
    ref8, _, _ = orc.process(copy.deepcopy(q).__class__(**{**q.__dict__, "bscansPerBuffer": 8}), small)
    assert_parity(outs["fused"][:8], ref8, q, max_frac_outside=1e-5, what="full size vs oracle on the unique tile")
    qf = copy.deepcopy(q); qf.bscanFlip = True
    fl, _ = run(qf, raw, _lib.FFT_FUSED)
    assert np.array_equal(fl[0::2], outs["fused"][0::2, ::-1]) and np.array_equal(fl[1::2], outs["fused"][1::2])
    # display extraction and output conversion straight from the volume in HBM
    p = OctPipeline(); qq = copy.deepcopy(q)
    assert p.initializeCuda(None, None, qq)
    p.octCudaPipeline(raw); p.sync()
    vol = p.copy_output(0)
    dB = torch.empty(n // 2 * a, dtype=torch.float32, device="cuda"); dE = torch.empty(a * b, dtype=torch.float32, device="cuda")
    for frame, nf, fn in ((5, 1, 0), (250, 16, 0), (3, 4, 1), (999, 1, 0)):
        p.changeDisplayedBscanFrame(frame, nf, fn, dB); p.changeDisplayedEnFaceFrame(frame, nf, fn, dE); p.sync()
        torch.cuda.synchronize()
        assert np.allclose(dB.cpu().numpy(), orc.bscan_frame(vol, n // 2, a, b, frame, nf, fn), rtol=2e-6, atol=1e-7)
        assert np.allclose(dE.cpu().numpy(), orc.enface_frame(vol, n // 2, a, b, frame, nf, fn), rtol=2e-6, atol=1e-7)
    conv = torch.empty(vol.size, dtype=torch.int16, device="cuda")
    p.float_to_output(0, conv); p.sync(); torch.cuda.synchronize()
    assert np.array_equal(conv.cpu().numpy().view(np.uint16).reshape(vol.shape), orc.float_to_output(vol, 12))
    p.cleanupCuda()


def test_full_size_config4_properties():
    """BASELINE config 4 at its full size -- 2048 x 1024 x 512, 16-bit (2 GiB raw, 2 GiB processed), full chain with FPN, B-scan
    flip and sinusoidal scan correction: the buffer is a tile of 8 unique B-scans, so the output is periodic, its first tile is
    bit-identical to processing the 8 B-scans alone (same FPN line, same kernels) and matches the oracle."""
    n, a, b = 2048, 1024, 512
    q = benchmark_params(n, a, b, 16); q.bscanFlip = True; q.sinusoidalScanCorrection = True; q.update_all_curves()
    small = synth.make_volume(n, a, 8, 16, resample=q.resampleCurve, dispersion=q.dispersionCurve)
    raw = np.ascontiguousarray(np.tile(small, (b // 8, 1, 1)))
    assert raw.nbytes == 2 << 30
    out, ml = run(q, raw, _lib.FFT_FUSED)
    del raw
    assert out.shape == (b, a, n // 2) and np.isfinite(out).all()
    # the reference's sinusoidal kernel leaves the very last line of a BUFFER untouched (cuda_code.cu:499: index < samples - width):
    # that line is the one place where a tile at the end of the buffer legitimately differs from a tile in its middle
    for k in (1, 17, 62):
        assert np.array_equal(out[:8], out[8 * k:8 * k + 8]), f"tile {k} differs from tile 0"
    last = out[8 * 63:8 * 64]
    assert np.array_equal(out[:7], last[:7]) and np.array_equal(out[7, :-1], last[7, :-1]), "last tile differs before its last line"
    assert not np.array_equal(out[7, -1], last[7, -1]), "the last line of the buffer is expected to stay uncorrected"
    qs = copy.deepcopy(q); qs.bscansPerBuffer = 8
    out8, ml8 = run(qs, small, _lib.FFT_FUSED)
    assert np.array_equal(ml8, ml), "FPN line must come from the first B-scan only (cuda_code.cu:1520-1522)"
    assert np.array_equal(out8, last), "full-size launch (end of the buffer) vs the unique tile processed alone"
    ref8, _, _ = orc.process(qs, small, mean_line=ml.astype(np.float64), determine_fpn=False)
    assert_parity(last, ref8, q, max_frac_outside=1e-5, what="config 4 full size vs oracle on the unique tile",
                  atol_abs=4e-6 * float(np.abs(ml).max()))


@pytest.mark.parametrize("mode", list(MODES))
def test_adversarial_lines(mode):
    n = 1024
    q = benchmark_params(n, 5, 1); q.fixedPatternNoiseRemoval = False; q.update_all_curves()
    raw = synth.adversarial_lines(n, 12).reshape(1, 5, n)
    ref, _, _ = orc.process(q, raw)
    out, _ = run(q, raw, MODES[mode])
    assert np.all(np.isneginf(out[0, 0])) and np.all(np.isneginf(ref[0, 0]))          # all-zero spectrum: log10(0), no clamp
    from tests.util import amplitude
    floor = 4e-7 * float(amplitude(ref[:, 1:], q).max())           # all energy sits in one or two bins: round-off of the peak
    assert_parity(out[:, 1:], ref[:, 1:], q, atol_abs=floor, max_frac_outside=1e-3, what=f"adversarial {mode}")


@pytest.mark.parametrize("bits,n", [(8, 1024), (32, 1024), (12, 1664), (8, 100), (14, 4096)])
def test_other_containers_and_line_lengths(bits, n):
    """u8 / u32 containers and non power-of-two lines (default settings.ini has width 1664) through AUTO mode"""
    q = benchmark_params(n, 8, 2, bits); q.fixedPatternNoiseRemoval = False; q.update_all_curves()
    raw = synth.make_volume(n, 8, 2, min(bits, 20), resample=q.resampleCurve, dispersion=q.dispersionCurve).astype(synth.container_dtype(bits))
    ref, _, _ = orc.process(q, raw)
    out, _ = run(q, raw, _lib.FFT_AUTO)
    assert_parity(out, ref, q, max_frac_outside=1e-4, what=f"bits={bits} N={n}")
    if bits == 32:
        q.bitshift = True
        ref, _, _ = orc.process(q, raw); out, _ = run(q, raw, _lib.FFT_AUTO)
        assert_parity(out, ref, q, max_frac_outside=1e-4, what="u32 bitshift")


@pytest.mark.parametrize("n,a,b", [(100, 7, 3), (1024, 1, 4), (1024, 2, 3), (1024, 3, 2), (1664, 5, 2), (2048, 64, 1)])
def test_sinusoidal_correction_geometries(n, a, b):
    """the sinusoidal kernel (cuda_code.cu:491-514) on line lengths that are not a multiple of four (scalar path), on tiny B-scans
    where the reference's flat addressing takes the second source line from the NEXT B-scan (A = 1, 2, 3: n + 1 = A), with the
    background removal folded in, and the untouched last line of the buffer"""
    q = benchmark_params(n, a, b); q.fixedPatternNoiseRemoval = False; q.sinusoidalScanCorrection = True; q.bscanFlip = True
    q.postProcessBackgroundRemoval = (a % 2 == 1); q.postProcessBackgroundWeight = 0.6; q.postProcessBackgroundOffset = 0.01
    q.update_all_curves()
    bg = (0.1 + 0.05 * np.cos(np.arange(n // 2) / 9.0)).astype(np.float32) if q.postProcessBackgroundRemoval else None
    raw = synth.make_volume(n, a, b, 12, resample=q.resampleCurve, dispersion=q.dispersionCurve)
    ref, _, _ = orc.process(q, raw, pp_background=bg)
    out, _ = run(q, raw, _lib.FFT_AUTO, pp_background=bg)
    assert_parity(out, ref, q, max_frac_outside=1e-4, saturated=bool(q.postProcessBackgroundRemoval), what=f"sinusoidal N={n} A={a} B={b}")
    plain = copy.deepcopy(q); plain.sinusoidalScanCorrection = False
    uncorrected, _ = run(plain, raw, _lib.FFT_AUTO, pp_background=bg)
    assert np.array_equal(out[-1, -1], uncorrected[-1, -1]), "the last line of the buffer stays uncorrected (cuda_code.cu:499)"


def test_fpn_determination_modes_and_slabs():
    n, a, b = 1024, 36, 2
    q = benchmark_params(n, a, b); q.buffersPerVolume = 2; q.bscansForNoiseDetermination = 2; q.update_all_curves()
    raw1 = synth.make_volume(n, a, b, 12, resample=q.resampleCurve, dispersion=q.dispersionCurve)
    raw2 = synth.make_volume(n, a, b, 12, resample=q.resampleCurve, dispersion=q.dispersionCurve, b_offset=7)
    p = OctPipeline(); qq = copy.deepcopy(q)
    assert p.initializeCuda(None, None, qq)
    p.octCudaPipeline(raw1); p.sync()
    assert p._lib.octb200_current_buffer_nr(p.handle) == 0          # starts at buffersPerVolume-1 then +1 (cuda_code.cu:1146,1531)
    ml1 = p.fpn_mean_line()
    p.octCudaPipeline(raw2); p.sync()
    assert p._lib.octb200_current_buffer_nr(p.handle) == 1
    assert np.array_equal(p.fpn_mean_line(), ml1)                    # determined once (cuda_code.cu:1521)
    slab0, slab1 = p.copy_output(0), p.copy_output(1)
    r1, _, _ = orc.process(q, raw1, mean_line=ml1.astype(np.float64), determine_fpn=False)
    r2, _, _ = orc.process(q, raw2, mean_line=ml1.astype(np.float64), determine_fpn=False)
    floor = 4e-6 * float(np.abs(ml1).max())                          # cancellation X - M: round-off of the subtracted line
    assert_parity(slab0, r1, q, atol_abs=floor, max_frac_outside=1e-4); assert_parity(slab1, r2, q, atol_abs=floor, max_frac_outside=1e-4)
    qq.redetermineFixedPatternNoise = True
    p.octCudaPipeline(raw2); p.sync()
    ml2 = p.fpn_mean_line()
    assert not np.array_equal(ml2, ml1)
    qq.continuousFixedPatternNoiseDetermination = True
    p.octCudaPipeline(raw1); p.sync()
    assert np.allclose(p.fpn_mean_line()[: n // 2], ml1[: n // 2], rtol=1e-5, atol=1e-3)
    p.cleanupCuda()


def test_postprocess_background_recording_and_removal():
    n, a, b = 1024, 16, 2
    q = benchmark_params(n, a, b); q.fixedPatternNoiseRemoval = False; q.postProcessBackgroundRemoval = True
    q.postProcessBackgroundWeight = 0.5; q.postProcessBackgroundOffset = 0.02; q.update_all_curves()
    raw = synth.make_volume(n, a, b, 12, resample=q.resampleCurve, dispersion=q.dispersionCurve)
    qn = copy.deepcopy(q); qn.postProcessBackgroundRemoval = False
    base, _, _ = orc.process(qn, raw)
    bg = orc.postprocess_background(base, n // 2, a)
    fired = []
    p = OctPipeline(); qq = copy.deepcopy(q)
    assert p.initializeCuda(None, None, qq)
    p.set_callbacks(background=lambda ptr: fired.append(ptr))
    qq.postProcessBackgroundRecordingRequested = True
    p.octCudaPipeline(raw); p.sync()
    assert len(fired) == 1                                            # Gpu2HostNotifier::backgroundSignalCallback (cuda_code.cu:655)
    assert np.allclose(p.postprocess_background(), bg, rtol=1e-4, atol=2e-5)
    out1 = p.copy_output(0)
    p.octCudaPipeline(raw); p.sync()                                 # second call: removal folded into the main kernel
    out2 = p.copy_output(0)
    p.cleanupCuda()
    ref, _, _ = orc.process(q, raw, pp_background=bg)
    assert_parity(out1, ref, q, saturated=True, max_frac_outside=1e-4); assert_parity(out2, ref, q, saturated=True, max_frac_outside=1e-4)


def test_streaming_to_host_and_callbacks():
    import torch
    n, a, b = 1024, 32, 4
    q = benchmark_params(n, a, b); q.fixedPatternNoiseRemoval = False; q.streamToHost = True; q.saveAs32bitFloat = True; q.update_all_curves()
    raw = synth.make_volume(n, a, b, 12, resample=q.resampleCurve, dispersion=q.dispersionCurve)
    s = [np.zeros((b, a, n // 2), np.uint16) for _ in range(2)]
    f = [np.zeros((b, a, n // 2), np.float32) for _ in range(2)]
    got = {"conv": [], "float": []}
    p = OctPipeline(); qq = copy.deepcopy(q)
    h1 = np.ascontiguousarray(raw).copy(); h2 = h1.copy()
    assert p.initializeCuda(h1, h2, qq)                               # pins the plugin's two buffers (cuda_code.cu:1135-1136)
    p.cuda_registerStreamingBuffers(s[0], s[1], s[0].nbytes)
    p.cuda_registerFloatStreamingBuffers(f[0], f[1], f[0].nbytes)
    p.set_callbacks(streaming=lambda ptr: got["conv"].append(ptr), float_streaming=lambda ptr: got["float"].append(ptr))
    for i in range(3):
        p.octCudaPipeline(h1 if i % 2 == 0 else h2)
    p.sync()
    vol = p.copy_output(0)
    assert len(got["conv"]) == 3 and len(got["float"]) == 3
    assert got["conv"][-1] in (s[0].ctypes.data, s[1].ctypes.data)
    last = s[0] if got["conv"][-1] == s[0].ctypes.data else s[1]
    assert np.array_equal(last, orc.float_to_output(vol, 12))         # floatToOutput (cuda_code.cu:943-967)
    lastf = f[0] if got["float"][-1] == f[0].ctypes.data else f[1]
    assert np.array_equal(lastf, vol)
    qq.streamingBuffersToSkip = 1
    n0 = len(got["conv"])
    for i in range(4):
        p.octCudaPipeline(h1)
    p.sync()
    assert len(got["conv"]) - n0 == 2                                 # every second buffer (cuda_code.cu:1358)
    p.cuda_unregisterStreamingBuffers(); p.cuda_unregisterFloatStreamingBuffers()
    p.cleanupCuda()


def test_host_and_device_paths_agree_and_null_reprocesses():
    import torch
    n, a, b = 2048, 16, 2
    q = benchmark_params(n, a, b, 16); q.update_all_curves()
    raw = synth.make_volume(n, a, b, 16, resample=q.resampleCurve, dispersion=q.dispersionCurve)
    p = OctPipeline(); assert p.initializeCuda(None, None, copy.deepcopy(q))
    p.octCudaPipeline(raw); p.sync(); o1 = p.copy_output(0)
    p.octCudaPipeline(None); p.sync(); o2 = p.copy_output(0)          # NULL: re-process the device copy (cuda_code.cu:1400)
    d = torch.from_numpy(raw.view(np.int16)).cuda()
    p.process_device(d); p.sync(); o3 = p.copy_output(0)
    bound = torch.zeros(o1.size, dtype=torch.float32, device="cuda")
    p.bind_output(bound); p.process_device(d); p.sync(); torch.cuda.synchronize()
    assert np.array_equal(o1, o2) and np.array_equal(o1, o3) and np.array_equal(bound.cpu().numpy().reshape(o1.shape), o1)
    p.cleanupCuda()


def test_error_behaviour():
    q = benchmark_params(1024, 8, 1)
    p = OctPipeline(fft_mode=_lib.FFT_SPLIT)
    q8 = copy.deepcopy(q); q8.samplesPerLine = 1664
    assert not p.initializeCuda(None, None, q8)                       # SPLIT exists for 1024 / 2048 only
    assert "unsupported" in p._create_error
    p = OctPipeline(fft_mode=_lib.FFT_FUSED)
    q8 = copy.deepcopy(q); q8.samplesPerLine = 1006                   # 2 * 503: no fused kernel for this length, and no silent fallback
    assert not p.initializeCuda(None, None, q8)
    assert "unsupported" in p._create_error
    p = OctPipeline()
    L = p._lib
    cfg = _lib.Config(1024, 8, 1, 1, 12, -1, 0, 0, 0); h = C.c_void_p()
    assert L.octb200_create(C.byref(cfg), C.byref(h)) == 0
    prm = q.to_c()
    assert L.octb200_set_params(h, C.byref(prm)) == 0
    raw = np.zeros(8 * 1024, np.uint16)
    assert L.octb200_process_host(h, raw.ctypes.data) == _lib.ERR_NOT_READY      # resampling on, no curve uploaded
    assert b"resample" in L.octb200_last_error(h)
    bad = np.zeros(100, np.float32)
    assert L.octb200_set_resample_curve(h, bad.ctypes.data, 100) == _lib.ERR_INVALID
    assert L.octb200_process_host(h, None) == _lib.ERR_NOT_READY                  # nothing uploaded yet
    assert L.octb200_destroy(h) == 0


@pytest.mark.parametrize("n,a,b", [(1024, 8, 3), (100, 70, 2), (2048, 130, 1), (1024, 64, 2)])
def test_volume_u8_layout(n, a, b):
    """updateDisplayedVolume (cuda_code.cu:915-941) as a tiled transpose: full and partial 64 x 64 tiles, A not a multiple of four
    (byte-wise stores), depth not a multiple of 64"""
    import torch
    q = benchmark_params(n, a, b); q.fixedPatternNoiseRemoval = False; q.update_all_curves()
    raw = synth.make_volume(n, a, b, 12, resample=q.resampleCurve, dispersion=q.dispersionCurve)
    p = OctPipeline(); assert p.initializeCuda(None, None, copy.deepcopy(q))
    p.octCudaPipeline(raw); p.sync(); vol = p.copy_output(0)
    assert vol.min() >= 0.0 and vol.max() < 1.0                        # inside [0,1): the conversion is a plain truncation
    tex = torch.full((n // 2 * b * a + 64,), 0xAB, dtype=torch.uint8, device="cuda")
    p.volume_u8(0, tex); p.sync(); torch.cuda.synchronize()
    host = tex.cpu().numpy()
    assert (host[n // 2 * b * a:] == 0xAB).all(), "wrote past the texture"
    t = host[: n // 2 * b * a].reshape(n // 2, b, a)                   # [z][B-scan][A-scan], z flipped (cuda_code.cu:935)
    expect = (vol.astype(np.float64) * 255.0).astype(np.uint8)
    assert np.array_equal(t, expect.transpose(2, 0, 1)[::-1])
    p.cleanupCuda()


def test_volume_u8_slab_of_a_multi_buffer_volume():
    """buffersPerVolume = 2: the second buffer lands in B-scans [B, 2B) of the texture (cuda_code.cu:931 bufferNr * bscansPerBuffer)"""
    import torch
    n, a, b = 1024, 12, 2
    q = benchmark_params(n, a, b); q.fixedPatternNoiseRemoval = False; q.buffersPerVolume = 2; q.update_all_curves()
    raws = [synth.make_volume(n, a, b, 12, resample=q.resampleCurve, dispersion=q.dispersionCurve, b_offset=4 * i) for i in range(2)]
    p = OctPipeline(); assert p.initializeCuda(None, None, copy.deepcopy(q))
    tex = torch.zeros(n // 2 * 2 * b * a, dtype=torch.uint8, device="cuda")
    vols = {}
    for raw in raws:
        p.octCudaPipeline(raw); p.sync()
        nr = p.current_buffer_nr()
        vols[nr] = p.copy_output(nr)
        p.volume_u8(nr, tex)
    p.sync(); torch.cuda.synchronize()
    assert sorted(vols) == [0, 1]
    t = tex.cpu().numpy().reshape(n // 2, 2 * b, a)
    for nr, vol in vols.items():
        expect = (vol.astype(np.float64) * 255.0).astype(np.uint8).transpose(2, 0, 1)[::-1]
        assert np.array_equal(t[:, nr * b:(nr + 1) * b], expect), f"slab {nr}"
    p.cleanupCuda()
