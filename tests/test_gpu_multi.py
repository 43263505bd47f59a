"""En-face gather over peer memory (include/octb200.h: octb200_enface_gather_*).
  * one GPU: the gather kernel with world = 1 must reproduce octb200_enface_frame bit for bit (same arithmetic, same order);
  * two or more GPUs (skipped on a single-GPU box): torchrun workers compare the peer gather with extraction + NCCL
    all_gather and with the un-sharded oracle (tests/multi_gpu_worker.py)."""
import copy
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from octproz_b200 import OctPipeline, _lib, benchmark_params, synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _window_tensor(ptr, n, torch, dev):
    class W:
        __cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2}
    return torch.as_tensor(W(), device=dev)


def test_single_rank_gather_equals_enface_frame():
    import torch
    n, a, b = 1024, 32, 6
    q = benchmark_params(n, a, b); q.fixedPatternNoiseRemoval = False; q.update_all_curves()
    raw = synth.make_volume(n, a, b, 12, resample=q.resampleCurve, dispersion=q.dispersionCurve)
    p = OctPipeline(fft_mode=_lib.FFT_FUSED)
    assert p.initializeCuda(None, None, copy.deepcopy(q)), getattr(p, "_create_error", "")
    p.octCudaPipeline(np.ascontiguousarray(raw)); p.sync()
    dev = torch.device("cuda", 0)
    handle = p.enface_gather_init(0, 1, a * b, 0)
    assert len(handle) == 64
    p.enface_gather_connect(handle)
    for (frame, nf, fn) in ((0, 1, 0), (17, 1, 0), (100, 7, 0), (500, 30, 0), (300, 4, 1), (9999, 1, 0)):
        want = torch.empty(a * b, dtype=torch.float32, device=dev)
        p.changeDisplayedEnFaceFrame(frame, nf, fn, want)
        p.enface_gather(frame, nf, fn)
        ptr = p.enface_gather_wait(); p.sync()
        got = _window_tensor(ptr, a * b, torch, dev).clone()
        assert torch.equal(got, want), (frame, nf, fn)
    # every gather is consumed behind itself (wait for all slabs, copy into the private display frame, acknowledge): the frame handed
    # out is always the latest gather; a healthy run has no device-side time-outs
    for f in range(8):
        p.enface_gather(f, 1, 0)
        ptr = p.enface_gather_wait(); p.sync()
        p.changeDisplayedEnFaceFrame(f, 1, 0, want); p.sync()
        assert torch.equal(_window_tensor(ptr, a * b, torch, dev).clone(), want), f
    st = p.enface_gather_status()
    assert st["ack_timeouts"] == 0 and st["arrival_timeouts"] == 0 and st["sequence"] == 14, st
    p.enface_gather_close()
    with pytest.raises(Exception):
        p.enface_gather(10, 1, 0)          # not connected any more: loud failure, no fallback
    # shard geometry is validated
    with pytest.raises(Exception):
        p.enface_gather_init(0, 1, a * b - 1, 0)
    p.cleanupCuda()


@pytest.mark.parametrize("n,kw", [(1024, {}), (1024, dict(bscanFlip=True)), (2048, dict(bscanFlip=True)),
                                  (1024, dict(bscanFlip=True, sinusoidalScanCorrection=True)),
                                  (1024, dict(fixedPatternNoiseRemoval=True)), (1024, dict(resamplingInterpolation=2))])
def test_automatic_gather_fused_into_the_epilogue(n, kw):
    """octb200_enface_gather_auto: every process call gathers the frame -- inside the fused kernel when the slab is final
    after it, by the appended stand-alone kernel otherwise (sinusoidal correction); both must equal octb200_enface_frame."""
    import torch
    a, b = 48, 5
    q = benchmark_params(n, a, b); q.fixedPatternNoiseRemoval = False
    for k, v in kw.items():
        setattr(q, k, v)
    q.update_all_curves()
    raw = np.ascontiguousarray(synth.make_volume(n, a, b, 12, resample=q.resampleCurve, dispersion=q.dispersionCurve))
    p = OctPipeline(fft_mode=_lib.FFT_FUSED)
    assert p.initializeCuda(None, None, copy.deepcopy(q)), getattr(p, "_create_error", "")
    dev = torch.device("cuda", 0)
    p.enface_gather_connect(p.enface_gather_init(0, 1, a * b, 0))
    launches, cases = [], []
    for (frame, nf, fn) in ((17, 1, 0), (n // 2 - 1, 1, 0), (40, 5, 0), (n // 2 - 3, 9, 0), (100, 40, 0), (60, 6, 1), (0, 70, 1)):
        p.enface_gather_auto(True, frame, nf, fn)
        l0 = p.launch_count()
        p.octCudaPipeline(raw)
        launches.append(p.launch_count() - l0); cases.append(nf)
        ptr = p.enface_gather_wait(); p.sync()
        got = _window_tensor(ptr, a * b, torch, dev).clone()
        want = torch.empty(a * b, dtype=torch.float32, device=dev)
        p.changeDisplayedEnFaceFrame(frame, nf, fn, want); p.sync()
        if nf == 1 or fn == 1:
            assert torch.equal(got, want), (frame, nf, fn)
        else:       # averaging: tree sum in the kernel vs sequential sum in the frame kernel
            assert torch.allclose(got, want, rtol=2e-6, atol=1e-6), (frame, nf, fn, float((got - want).abs().max()))
    p.enface_gather_auto(False)
    l0 = p.launch_count(); p.octCudaPipeline(raw); p.sync()
    base = p.launch_count() - l0
    sinus = bool(kw.get("sinusoidalScanCorrection"))
    # one displayed depth frame: fused, no extra launch for the gather; multi-frame average / MIP or a later pass over the slab
    # (sinusoidal correction): the stand-alone gather kernel is appended to the chain
    # (+ 1: the consumer kernel behind every gather)
    assert all(l == base + 1 + (1 if (sinus or nf > 1) else 0) for l, nf in list(zip(launches, cases))[1:]), (launches, cases, base)
    p.enface_gather_close()
    p.cleanupCuda()


def test_two_ranks_peer_gather_matches_nccl_and_oracle():
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    world = 2 if ngpu < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29611", os.path.join(ROOT, "tests", "multi_gpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("MULTI_GPU_RESULT ")]
    assert line, r.stdout[-2000:]
    res = json.loads(line[0][len("MULTI_GPU_RESULT "):])
    assert res["world"] == world and res["max_abs_err_vs_oracle"] < 1e-3
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"enface_gather_n{world}.json"), "w"))
