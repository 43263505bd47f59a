"""The two things the fused kernel's epilogue can do besides writing the float line -- the converted u16 line for the stream-to-host
path (CONV kernels) and the en-face capture for the peer gather -- in the SAME launch: what every rank of the multi-GPU end-to-end
loop runs (bench.py e2e leg at N > 1).  One rank is enough to exercise the kernel variant; the peer stores themselves are covered
by tests/test_gpu_multi.py.  (Sorted last on purpose: it combines features the earlier files test one by one.)"""
import copy

import numpy as np
import pytest

from octproz_b200 import OctPipeline, _lib, benchmark_params, synth
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,bits,kw", [(1024, 12, {}), (1024, 12, dict(bscanFlip=True, fixedPatternNoiseRemoval=False)), (2048, 16, {})])
def test_streaming_and_automatic_gather_in_one_launch(n, bits, kw):
    import torch
    from tests.test_gpu_multi import _window_tensor
    a, b = 40, 4
    q = benchmark_params(n, a, b, bits)
    for k, v in kw.items():
        setattr(q, k, v)
    q.streamToHost = True
    q.update_all_curves()
    raw = np.ascontiguousarray(synth.make_volume(n, a, b, bits, resample=q.resampleCurve, dispersion=q.dispersionCurve))
    s = [np.zeros((b, a, n // 2), np.uint16) for _ in range(2)]
    got_cb = []
    p = OctPipeline(fft_mode=_lib.FFT_FUSED)
    assert p.initializeCuda(None, None, copy.deepcopy(q)), getattr(p, "_create_error", "")
    p.cuda_registerStreamingBuffers(s[0], s[1], s[0].nbytes)
    p.set_callbacks(streaming=lambda ptr: got_cb.append(ptr))
    p.enface_gather_connect(p.enface_gather_init(0, 1, a * b, 0))
    p.octCudaPipeline(raw); p.sync()                      # FPN determination happens here
    dev = torch.device("cuda", 0)
    for frame in (17, n // 2 - 1, 300):
        p.enface_gather_auto(True, frame, 1, 0)
        l0 = p.launch_count()
        p.octCudaPipeline(raw)
        launches = p.launch_count() - l0
        ptr = p.enface_gather_wait(); p.sync()
        gathered = _window_tensor(ptr, a * b, torch, dev).clone()
        want = torch.empty(a * b, dtype=torch.float32, device=dev)
        p.changeDisplayedEnFaceFrame(frame, 1, 0, want); p.sync()
        vol = p.copy_output(0)
        last = s[0] if got_cb[-1] == s[0].ctypes.data else s[1]
        # compute + converted output + en-face capture + peer stores + publish: one kernel; + the consumer kernel behind the gather
        assert launches == 2, launches
        assert torch.equal(gathered, want), frame
        assert np.array_equal(last, orc.float_to_output(vol, bits)), frame
    p.enface_gather_close()
    p.cuda_unregisterStreamingBuffers()
    p.cleanupCuda()


@pytest.mark.skipif(not orc.have_ref("libref_cuda.so"), reason="oracle/_ref/libref_cuda.so not built")
@pytest.mark.parametrize("bits,kw", [(12, {}), (16, dict(signalGrayscaleMin=20.0, signalGrayscaleMax=70.0, signalMultiplicator=1.4)), (10, {})])
def test_float_to_output_restatement_pinned_to_the_reference_kernel(bits, kw):
    """the oracle's floatToOutput (and with it the fused conversion, which is tested bit-identical to it) against the reference's OWN
    kernel: the reference streams its converted buffer to the host (cuda_code.cu:1357-1372); applying the restatement to the float
    volume the same run produced must give exactly those containers -- including values saturated at both ends"""
    import ctypes as C
    n, a, b = 1024, 24, 2
    q = benchmark_params(n, a, b, bits); q.fixedPatternNoiseRemoval = False; q.streamToHost = True
    for k, v in kw.items():
        setattr(q, k, v)
    rc = orc.RefCuda(); rc.configure(q)
    q.resampleCurve, q.dispersionCurve, q.windowCurve = rc.curves()
    raw = synth.make_volume(n, a, b, bits, resample=q.resampleCurve, dispersion=q.dispersionCurve)
    h1 = np.ascontiguousarray(raw).copy(); h2 = h1.copy()
    rc.init(h1, h2)
    s1 = np.zeros((b, a, n // 2), np.uint16); s2 = np.zeros_like(s1)
    rc.L.refcuda_register_streaming(s1.ctypes.data, s2.ctypes.data, C.c_size_t(s1.nbytes))
    rc.process(h1)
    rc.L.refcuda_sync()
    vol = rc.output(0)
    want = orc.float_to_output(vol, bits)
    ok = np.array_equal(s1, want) or np.array_equal(s2, want)
    rc.L.refcuda_unregister_streaming()
    rc.cleanup()
    assert ok, "oracle floatToOutput differs from the reference's kernel on the reference's own volume"
    assert want.min() < want.max()


def test_cpp_replay_matches_the_python_mirror(tmp_path):
    """the C++ host loop and the Python mirror drive the same library: same number of launches per buffer, same output"""
    import json
    import subprocess

    from octproz_b200.acquisition import write_raw_file
    from tests.test_host_mirror import build
    n, a, b = 1024, 32, 4
    q = benchmark_params(n, a, b); q.update_all_curves()
    vol = synth.make_volume(n, a, 2 * b, 12, resample=q.resampleCurve, dispersion=q.dispersionCurve)
    path = str(tmp_path / "two_buffers.raw")
    write_raw_file(path, vol)
    exe = str(tmp_path / "replay")
    build("examples/replay_main.cpp", exe)
    r = subprocess.run([exe, path, str(n), str(a), str(b), "12", "6"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    res = json.loads(r.stdout.strip().splitlines()[-1])
    assert res["processed_buffers"] == 6 and res["ascans_per_s"] > 0
    # the two file buffers alternate; which one is processed first depends on the start-up race between the acquisition thread and
    # the buffer blocking of Processing::slot_start (the same in the reference, processing.cpp:124-134): 0,1,0,1,0,1 or 1,0,1,0,1,0.
    # The FPN line comes from the first processed buffer, the reported output is the last processed one.
    halves = (np.ascontiguousarray(vol[:b]), np.ascontiguousarray(vol[b:]))
    candidates = []
    for first in (0, 1):
        p = OctPipeline(); assert p.initializeCuda(None, None, copy.deepcopy(q))
        p.octCudaPipeline(halves[first]); p.sync()
        p.octCudaPipeline(halves[1 - first]); p.sync()
        candidates.append(float(p.copy_output(0).astype(np.float64).sum()))
        p.cleanupCuda()
    assert any(abs(res["output_sum"] - want) <= 1e-6 * abs(want) + 1e-3 for want in candidates), (res["output_sum"], candidates)
    assert res["launches"] >= 6


def test_cpp_dispersion_engine_reproduces_the_reference_search(tmp_path):
    """examples/estimate_dispersion_main.cpp (the C++ DispersionEstimationEngine of include/octb200_host.hpp over
    octb200_dispersion_sweep) on the golden frame: the same best d2 / d3 / d1 as the search that ran every trial through the
    reference's CPU path and metric code (tests/golden/estimator.npz), with two sweeps + the two plotted A-scans = 3 sweep calls"""
    import json
    import os
    import subprocess

    from tests.test_host_mirror import build
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "estimator.npz"))
    raw = np.ascontiguousarray(g["raw"]).astype(np.uint16)
    lines, n = raw.shape
    path = str(tmp_path / "frame.raw")
    raw.tofile(path)
    exe = str(tmp_path / "estimate")
    build("examples/estimate_dispersion_main.cpp", exe)
    r = subprocess.run([exe, path, str(n), str(lines), "12", "-160", "0", "-40", "40", "16", str(lines)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    res = json.loads(r.stdout.strip().splitlines()[-1])
    want_d2, want_d3, want_d1 = (float(x) for x in g["search_log0"])
    assert (res["bestD2"], res["bestD3"]) == (want_d2, want_d3), (res, g["search_log0"])
    assert abs(res["calculatedD1"] - want_d1) < 1e-9
    assert res["bestMetricValueD2"] > 0 and res["bestMetricValueD3"] >= res["bestMetricValueD2"] * 0.999
