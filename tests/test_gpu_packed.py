"""12-bit packed input (OCTB200_PACK_12P): the same samples delivered packed must give bit-identical output to the container
path -- directly in the fused kernel's slot conversion (4-tap / plain stages) and through the unpack pre-pass (Lanczos, rolling
mean, SPLIT and CUFFT chains); host and device entry points; error behaviour."""
import copy

import numpy as np
import pytest

from octproz_b200 import OctPipeline, _lib, benchmark_params, synth
from octproz_b200.packing import pack12

pytestmark = pytest.mark.gpu


def run(q, raw, packing, mode, device_resident=False):
    p = OctPipeline(fft_mode=mode, input_packing=packing)
    assert p.initializeCuda(None, None, copy.deepcopy(q)), getattr(p, "_create_error", "")
    buf = np.ascontiguousarray(raw)
    launches0 = p.launch_count()
    if device_resident:
        import torch
        d = torch.from_numpy(buf.view(np.uint8).reshape(-1)).cuda()
        p.process_device(d)
    else:
        p.octCudaPipeline(buf)
    p.sync()
    out, launches = p.copy_output(0), p.launch_count() - launches0
    p.cleanupCuda()
    return out, launches


CASES = {"benchmark": {}, "nofpn": dict(fixedPatternNoiseRemoval=False), "linear": dict(resamplingInterpolation=0),
         "noresample": dict(resampling=False), "flip": dict(bscanFlip=True), "lanczos": dict(resamplingInterpolation=2),
         "rolling": dict(backgroundRemoval=True, rollingAverageWindowSize=16), "linscale": dict(signalLogScaling=False, signalGrayscaleMax=400.0)}


@pytest.mark.parametrize("n", [1024, 2048])
@pytest.mark.parametrize("name", list(CASES))
def test_packed_equals_containers_fused(n, name):
    q = benchmark_params(n, 24, 3)
    for k, v in CASES[name].items():
        setattr(q, k, v)
    q.update_all_curves()
    raw = synth.make_volume(n, 24, 3, 12, resample=q.resampleCurve, dispersion=q.dispersionCurve)
    want, l0 = run(q, raw, _lib.PACK_CONTAINER, _lib.FFT_FUSED)
    got, l1 = run(q, pack12(raw), _lib.PACK_12P, _lib.FFT_FUSED)
    assert np.array_equal(got, want), name
    # direct path: no extra launch; Lanczos / rolling mean: one unpack kernel in front of the chain
    assert l1 - l0 == (1 if name in ("lanczos", "rolling") else 0), (name, l0, l1)
    got_d, _ = run(q, pack12(raw), _lib.PACK_12P, _lib.FFT_FUSED, device_resident=True)
    assert np.array_equal(got_d, want)


@pytest.mark.parametrize("mode", [_lib.FFT_SPLIT, _lib.FFT_CUFFT])
def test_packed_equals_containers_other_modes(mode):
    q = benchmark_params(1024, 16, 2); q.update_all_curves()
    raw = synth.make_volume(1024, 16, 2, 12, resample=q.resampleCurve, dispersion=q.dispersionCurve)
    want, _ = run(q, raw, _lib.PACK_CONTAINER, mode)
    got, _ = run(q, pack12(raw), _lib.PACK_12P, mode)
    assert np.array_equal(got, want)


def test_packed_adversarial_lines_and_errors():
    n = 1024
    q = benchmark_params(n, 8, 1); q.fixedPatternNoiseRemoval = False; q.update_all_curves()
    raw = np.zeros((1, 8, n), np.uint16)
    raw[0, 1] = 4095; raw[0, 2, 17] = 4095; raw[0, 3, ::2] = 4095; raw[0, 4] = np.arange(n) * 4 % 4096
    raw[0, 5] = np.random.default_rng(1).integers(0, 4096, n); raw[0, 6, -1] = 4095; raw[0, 7, 0] = 1
    want, _ = run(q, raw, _lib.PACK_CONTAINER, _lib.FFT_FUSED)
    got, _ = run(q, pack12(raw), _lib.PACK_12P, _lib.FFT_FUSED)
    assert np.array_equal(np.nan_to_num(got, neginf=-1e30), np.nan_to_num(want, neginf=-1e30))
    # packing needs 12-bit data
    q16 = benchmark_params(n, 8, 1, 16)
    p = OctPipeline(input_packing=_lib.PACK_12P)
    assert not p.initializeCuda(None, None, q16) and "bitDepth 12" in p._create_error
