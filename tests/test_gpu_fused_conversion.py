"""floatToOutput (cuda_code.cu:943-967) folded into the fused kernel's epilogue: the u16 buffer streamed to the host must be
bit-identical to (a) the oracle's floatToOutput of the float volume the same call produced, (b) the stand-alone conversion pass
(OCTB200_FLAG_SEPARATE_CONVERSION, the reference's own order of kernels, cuda_code.cu:1366) -- over interpolators, FPN, flip,
log / linear scaling that over- and undershoots [0, 1], background removal, 10/12/16-bit containers, N = 1024 / 2048, packed input,
and chains where the slab is NOT final after the main kernel (sinusoidal correction: the kernel that writes the final slab
converts, in every FFT mode; u8 containers and SPLIT / CUFFT chains without it keep the separate pass)."""
import copy

import numpy as np
import pytest

from octproz_b200 import OctPipeline, _lib, benchmark_params, synth
from octproz_b200.packing import pack12
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def stream(q, raw, flags=0, packing=_lib.PACK_CONTAINER, calls=1):
    """process `raw` `calls` times with stream-to-host on; returns (float volume, last streamed container buffer, launches per call)"""
    n, a, b = int(q.samplesPerLine), int(q.ascansPerBscan), int(q.bscansPerBuffer)
    qq = copy.deepcopy(q); qq.streamToHost = True
    s = [np.zeros((b, a, n // 2), np.uint16) for _ in range(2)]
    got = []
    p = OctPipeline(fft_mode=_lib.FFT_FUSED, input_packing=packing, flags=flags)
    assert p.initializeCuda(None, None, qq), getattr(p, "_create_error", "")
    p.cuda_registerStreamingBuffers(s[0], s[1], s[0].nbytes)
    p.set_callbacks(streaming=lambda ptr: got.append(ptr))
    buf = np.ascontiguousarray(raw)
    p.octCudaPipeline(buf); p.sync()                   # first call may carry the FPN determination launches
    l0 = p.launch_count()
    for _ in range(calls):
        p.octCudaPipeline(buf)
    p.sync()
    per_call = (p.launch_count() - l0) / calls
    vol = p.copy_output(0)
    assert len(got) == calls + 1
    last = s[0] if got[-1] == s[0].ctypes.data else s[1]
    out = last.copy()
    p.cuda_unregisterStreamingBuffers(); p.cleanupCuda()
    return vol, out, per_call


CASES = {
    "benchmark": {},
    "nofpn_flip": dict(fixedPatternNoiseRemoval=False, bscanFlip=True),
    "linear_interp": dict(resamplingInterpolation=0),
    "lanczos": dict(resamplingInterpolation=2),
    "noresample_rolling": dict(resampling=False, backgroundRemoval=True, rollingAverageWindowSize=16),
    # scaling chosen so that a good part of the output leaves [0, 1] on both sides: the saturation of cuda_code.cu:946 matters
    "log_overshoot": dict(signalGrayscaleMin=20.0, signalGrayscaleMax=60.0, signalMultiplicator=1.3),
    "linscale": dict(signalLogScaling=False, signalGrayscaleMin=0.0, signalGrayscaleMax=40.0),
    "ppbg": dict(postProcessBackgroundRemoval=True, postProcessBackgroundWeight=0.7, postProcessBackgroundOffset=0.01),
}


@pytest.mark.parametrize("shape", [(1024, 12), (2048, 16), (1024, 10)], ids=["1024x12bit", "2048x16bit", "1024x10bit"])
@pytest.mark.parametrize("name", list(CASES))
def test_fused_conversion_bit_identical(shape, name):
    n, bits = shape
    q = benchmark_params(n, 24, 3, bits)
    for k, v in CASES[name].items():
        setattr(q, k, v)
    q.update_all_curves()
    if name == "ppbg":
        q.loadPostProcessingBackground(0.05 + 0.02 * np.cos(np.arange(n // 2) / 17.0))
    raw = synth.make_volume(n, 24, 3, bits, resample=q.resampleCurve, dispersion=q.dispersionCurve)
    vol, conv, l_fused = stream(q, raw)
    vol_s, conv_s, l_sep = stream(q, raw, flags=_lib.FLAG_SEPARATE_CONVERSION)
    assert np.array_equal(vol, vol_s), "the float volume must not depend on where the conversion runs"
    want = orc.float_to_output(vol, bits)
    assert np.array_equal(conv_s, want), "stand-alone floatToOutput vs oracle"
    assert np.array_equal(conv, want), "fused floatToOutput vs oracle"
    assert l_sep - l_fused == 1, (l_fused, l_sep)       # the conversion pass is really gone
    assert want.min() < want.max()
    if name == "log_overshoot":
        top = (1 << (10 if bits <= 10 else 12 if bits <= 12 else 16)) - 1
        assert (want == top).any(), "this case is meant to exercise the saturation of cuda_code.cu:946"


def test_fused_conversion_packed_input_and_buffer_alternation():
    n = 1024
    q = benchmark_params(n, 16, 2); q.update_all_curves()
    raw = synth.make_volume(n, 16, 2, 12, resample=q.resampleCurve, dispersion=q.dispersionCurve)
    vol, conv, _ = stream(q, raw)
    vol_p, conv_p, l_p = stream(q, pack12(raw), packing=_lib.PACK_12P)
    assert np.array_equal(vol_p, vol) and np.array_equal(conv_p, conv) and l_p == 1
    # alternating device buffers over several calls (cuda_code.cu:1360): every call converts into the other buffer
    _, conv3, l3 = stream(q, raw, calls=3)
    assert np.array_equal(conv3, conv) and l3 == 1


def stream_mode(q, raw, mode, flags=0):
    """like stream() for any FFT mode: (float volume, streamed containers of the second call, launches of the second call)"""
    n, a, b = int(q.samplesPerLine), int(q.ascansPerBscan), int(q.bscansPerBuffer)
    qq = copy.deepcopy(q); qq.streamToHost = True
    dt = np.uint8 if q.bitDepth <= 8 else np.uint16
    s = [np.zeros((b, a, n // 2), dt) for _ in range(2)]
    got = []
    p = OctPipeline(fft_mode=mode, flags=flags)
    assert p.initializeCuda(None, None, qq), getattr(p, "_create_error", "")
    p.cuda_registerStreamingBuffers(s[0], s[1], s[0].nbytes)
    p.set_callbacks(streaming=lambda ptr: got.append(ptr))
    buf = np.ascontiguousarray(raw)
    p.octCudaPipeline(buf); p.sync()
    l0 = p.launch_count()
    p.octCudaPipeline(buf); p.sync()
    launches = p.launch_count() - l0
    vol = p.copy_output(0)
    out = (s[0] if got[-1] == s[0].ctypes.data else s[1]).copy()
    p.cuda_unregisterStreamingBuffers(); p.cleanupCuda()
    return vol, out, launches


@pytest.mark.parametrize("mode,launches", [(_lib.FFT_FUSED, 2), (_lib.FFT_SPLIT, 3), (_lib.FFT_CUFFT, 4)], ids=["fused", "split", "cufft"])
def test_conversion_folded_into_the_sinusoidal_kernel(mode, launches):
    """sinusoidal scan correction rewrites the slab after the main kernel (cuda_code.cu:1552): the kernel that writes the FINAL slab
    also writes the converted line, whatever ran the FFT -- no separate floatToOutput pass, same bits"""
    n = 1024
    q = benchmark_params(n, 32, 2); q.sinusoidalScanCorrection = True; q.bscanFlip = True; q.update_all_curves()
    raw = synth.make_volume(n, 32, 2, 12, resample=q.resampleCurve, dispersion=q.dispersionCurve)
    vol, conv, l = stream_mode(q, raw, mode)
    vol_s, conv_s, l_s = stream_mode(q, raw, mode, flags=_lib.FLAG_SEPARATE_CONVERSION)
    assert np.array_equal(vol, vol_s) and np.array_equal(conv, conv_s)
    assert np.array_equal(conv, orc.float_to_output(vol, 12))
    assert l == launches and l_s == launches + 1, (l, l_s)


def test_separate_pass_kept_where_nothing_can_fold_it():
    """SPLIT / CUFFT chains without a sinusoidal pass, and u8 containers, keep floatToOutput as its own kernel"""
    n = 1024
    q = benchmark_params(n, 16, 2); q.update_all_curves()
    raw = synth.make_volume(n, 16, 2, 12, resample=q.resampleCurve, dispersion=q.dispersionCurve)
    for mode, launches in ((_lib.FFT_SPLIT, 3), (_lib.FFT_CUFFT, 4)):
        vol, conv, l = stream_mode(q, raw, mode)
        assert np.array_equal(conv, orc.float_to_output(vol, 12)) and l == launches, (mode, l)
    q8 = benchmark_params(n, 16, 2, 8); q8.sinusoidalScanCorrection = True; q8.update_all_curves()
    raw8 = synth.make_volume(n, 16, 2, 8, resample=q8.resampleCurve, dispersion=q8.dispersionCurve)
    vol, conv, l = stream_mode(q8, raw8, _lib.FFT_AUTO)
    assert conv.dtype == np.uint8 and np.array_equal(conv, orc.float_to_output(vol, 8))


def test_special_values_convert_like_the_reference():
    """all-zero lines give log(0) = -inf, which saturates to 0 (cuda_code.cu:946 __saturatef); full-scale lines stay finite"""
    n = 1024
    q = benchmark_params(n, 8, 1); q.fixedPatternNoiseRemoval = False; q.update_all_curves()
    raw = np.zeros((1, 8, n), np.uint16); raw[0, 1] = 4095; raw[0, 2, 5] = 4095; raw[0, 3, ::2] = 4095
    vol, conv, _ = stream(q, raw)
    assert np.isneginf(vol[0, 0]).all()
    assert np.array_equal(conv, orc.float_to_output(vol, 12))
    assert (conv[0, 0] == 0).all()
