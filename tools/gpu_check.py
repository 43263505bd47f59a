"""Development harness run on the B200 box through gpurun (not part of the product, not a test):
   python tools/gpu_check.py parity   -> every mode x parameter variant against the fp64 oracle
   python tools/gpu_check.py refcuda  -> against the reference's unmodified cuda_code.cu (oracle/_ref/libref_cuda.so)
   python tools/gpu_check.py timing   -> device-resident / end-to-end timings of all modes and of the reference CUDA build
Results are printed and written to gpurun_out/*.json."""
from __future__ import annotations

import copy
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from octproz_b200 import OctPipeline, _lib, benchmark_params, synth  # noqa: E402
from oracle import oracle as orc  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
os.makedirs(OUT, exist_ok=True)
MODES = {"fused": _lib.FFT_FUSED, "split": _lib.FFT_SPLIT, "cufft": _lib.FFT_CUFFT}


def variants(n):
    v = {}
    b = benchmark_params(n, 64, 8)
    v["benchmark"] = b
    q = copy.deepcopy(b); q.fixedPatternNoiseRemoval = False; v["nofpn"] = q
    q = copy.deepcopy(b); q.fixedPatternNoiseRemoval = False; q.resamplingInterpolation = 0; v["linear"] = q
    q = copy.deepcopy(b); q.fixedPatternNoiseRemoval = False; q.resamplingInterpolation = 2; v["lanczos"] = q
    q = copy.deepcopy(b); q.fixedPatternNoiseRemoval = False; q.resampling = False; v["noresample"] = q
    q = copy.deepcopy(b); q.fixedPatternNoiseRemoval = False; q.windowing = False; q.dispersionCompensation = False; v["klin_only"] = q
    q = copy.deepcopy(b); q.fixedPatternNoiseRemoval = False; q.backgroundRemoval = True; q.rollingAverageWindowSize = 64; v["rolling64"] = q
    q = copy.deepcopy(b); q.fixedPatternNoiseRemoval = False; q.backgroundRemoval = True; q.rollingAverageWindowSize = 8; q.resamplingInterpolation = 2; v["rolling8_lanczos"] = q
    q = copy.deepcopy(b); q.fixedPatternNoiseRemoval = False; q.bscanFlip = True; q.sinusoidalScanCorrection = True; v["flip_sinus"] = q
    q = copy.deepcopy(b); q.fixedPatternNoiseRemoval = False; q.bscanFlip = True; v["flip"] = q
    q = copy.deepcopy(b); q.fixedPatternNoiseRemoval = False; q.signalLogScaling = False; q.signalGrayscaleMin = 0.0; q.signalGrayscaleMax = 400.0; v["linscale"] = q
    q = copy.deepcopy(b); q.fixedPatternNoiseRemoval = False; q.bitshift = True; q.bitDepth = 16; v["bitshift16"] = q
    return v


def stats(a, b):
    d = np.abs(a.astype(np.float64) - b.astype(np.float64))
    fin = np.isfinite(d)
    d = d[fin]
    return {"max": float(d.max()), "p9999": float(np.quantile(d, 0.9999)), "p99": float(np.quantile(d, 0.99)),
            "median": float(np.median(d)), "frac_gt_1e-4": float((d > 1e-4).mean()), "nonfinite": int((~fin).sum())}


def run_mine(q, raw, mode, mean_line=None):
    p = OctPipeline(fft_mode=mode)
    h = np.ascontiguousarray(raw)
    assert p.initializeCuda(None, None, q), getattr(p, "_create_error", "")
    if mean_line is not None:
        p.set_fpn_mean_line(mean_line)
    p.octCudaPipeline(h)
    p.sync()
    out = p.copy_output(0)
    ml = p.fpn_mean_line()
    p.cleanupCuda()
    return out, ml


def cmd_parity():
    res = {}
    for n in (1024, 2048):
        for name, q in variants(n).items():
            q = copy.deepcopy(q); q.update_all_curves()
            raw = synth.make_volume(n, q.ascansPerBscan, q.bscansPerBuffer, q.bitDepth, resample=q.resampleCurve, dispersion=q.dispersionCurve)
            ref, ml, _ = orc.process(q, raw)
            for mname, mode in MODES.items():
                key = f"N{n}/{name}/{mname}"
                try:
                    out, mml = run_mine(copy.deepcopy(q), raw, mode)
                    s = stats(out, ref)
                    if q.fixedPatternNoiseRemoval:
                        h = n // 2
                        s["meanline_max_diff"] = float(np.abs(mml[:h] - ml[:h]).max())
                        out2, _ = run_mine(copy.deepcopy(q), raw, mode, mean_line=ml.astype(np.float32))
                        s["with_oracle_meanline"] = stats(out2, ref)
                    res[key] = s
                    print(key, json.dumps(s), flush=True)
                except Exception as e:  # noqa: BLE001
                    res[key] = {"error": str(e)}
                    print(key, "ERROR", e, flush=True)
    json.dump(res, open(os.path.join(OUT, "parity_oracle.json"), "w"), indent=1)


def cmd_refcuda():
    res = {}
    rc = orc.RefCuda()
    for n in (1024, 2048):
        for name, q in variants(n).items():
            q = copy.deepcopy(q)
            rc.configure(q)
            q.resampleCurve, q.dispersionCurve, q.windowCurve = rc.curves()   # identical LUTs for both sides
            raw = synth.make_volume(n, q.ascansPerBscan, q.bscansPerBuffer, q.bitDepth, resample=q.resampleCurve, dispersion=q.dispersionCurve)
            h1 = np.ascontiguousarray(raw).copy(); h2 = h1.copy()
            rc.init(h1, h2)
            rc.process(h1)
            ref = rc.output(0)
            rc.cleanup()
            orc64, _, _ = orc.process(q, raw)
            sref = stats(ref, orc64)
            print(f"N{n}/{name}/REFCUDA-vs-oracle", json.dumps(sref), flush=True)
            res[f"N{n}/{name}/refcuda_vs_oracle"] = sref
            for mname, mode in MODES.items():
                key = f"N{n}/{name}/{mname}"
                try:
                    out, _ = run_mine(copy.deepcopy(q), raw, mode)
                    s = stats(out, ref)
                    res[key] = s
                    print(key, "vs REFCUDA", json.dumps(s), flush=True)
                except Exception as e:  # noqa: BLE001
                    res[key] = {"error": str(e)}
                    print(key, "ERROR", e, flush=True)
    json.dump(res, open(os.path.join(OUT, "parity_refcuda.json"), "w"), indent=1)


def cmd_timing():
    import torch
    res = {}
    for (n, a, b, bits) in ((1024, 512, 256, 12), (2048, 1024, 128, 16)):
        q = benchmark_params(n, a, b, bits); q.update_all_curves()
        t0 = time.time()
        small = synth.make_volume(n, a, 8, bits, resample=q.resampleCurve, dispersion=q.dispersionCurve)
        raw = np.ascontiguousarray(np.tile(small, (b // 8, 1, 1)))
        h1 = torch.from_numpy(raw).pin_memory(); h2 = torch.from_numpy(raw.copy()).pin_memory()
        print("synth", time.time() - t0, raw.shape, flush=True)
        ascans = a * b
        for mname, mode in MODES.items():
            p = OctPipeline(fft_mode=mode)
            assert p.initializeCuda(None, None, copy.deepcopy(q)), getattr(p, "_create_error", "")
            p.octCudaPipeline(h1.numpy()); p.sync()
            for _ in range(3):
                p.process_device(None)
            p.sync()
            iters = 20
            p.event_record(0)
            for _ in range(iters):
                p.process_device(None)
            p.event_record(1)
            ms = p.event_elapsed_ms(0, 1) / iters
            kms = p.time_kernel(p._lib.octb200_output_device_ptr(p.handle, 0) and torch.from_numpy(raw).cuda(), 10) if False else None
            # end to end: pinned host -> H2D -> kernels
            for _ in range(2):
                p.octCudaPipeline(h1.numpy()); p.octCudaPipeline(h2.numpy())
            p.sync()
            t0 = time.perf_counter()
            for i in range(10):
                p.octCudaPipeline((h1 if i % 2 == 0 else h2).numpy())
            p.sync()
            e2e = (time.perf_counter() - t0) / 10
            key = f"N{n}x{a}x{b}/{mname}"
            res[key] = {"resident_ms": ms, "resident_MHz": ascans / ms / 1e3, "e2e_ms": e2e * 1e3, "e2e_MHz": ascans / e2e / 1e6,
                        "algorithmic_GBs": ascans * n * 4 / ms / 1e6}
            print(key, json.dumps(res[key]), flush=True)
            p.cleanupCuda()
        # reference CUDA build, same box
        try:
            rc = orc.RefCuda(); rc.configure(q)
            a1 = np.ascontiguousarray(raw).copy(); a2 = a1.copy()
            rc.init(a1, a2)
            rc.process(a1); rc.L.refcuda_sync()
            tres = rc.time(None, None, 20, 3) / 20
            te2e = rc.time(a1, a2, 10, 2) / 10
            rc.cleanup()
            key = f"N{n}x{a}x{b}/refcuda"
            res[key] = {"resident_ms": tres * 1e3, "resident_MHz": ascans / tres / 1e6, "e2e_ms": te2e * 1e3, "e2e_MHz": ascans / te2e / 1e6}
            print(key, json.dumps(res[key]), flush=True)
        except Exception as e:  # noqa: BLE001
            print("refcuda timing failed", e, flush=True)
        del h1, h2
    json.dump(res, open(os.path.join(OUT, "timing.json"), "w"), indent=1)


def cmd_quick():
    for n in (1024, 2048):
        q = benchmark_params(n, 16, 2); q.update_all_curves()
        raw = synth.make_volume(n, 16, 2, 12, resample=q.resampleCurve, dispersion=q.dispersionCurve)
        ref, ml, _ = orc.process(q, raw)
        for mname, mode in MODES.items():
            out, mml = run_mine(copy.deepcopy(q), raw, mode)
            print("quick", n, mname, json.dumps(stats(out, ref)), flush=True)


def cmd_perf():
    """resident timing of the fused kernel + whole chain for both workloads (use OCTB200_LIB to pick an A/B build)"""
    import torch
    tag = os.environ.get("OCTB200_TAG", "default")
    res = {}
    for (n, a, b, bits) in ((1024, 512, 256, 12), (2048, 1024, 128, 16)):
        q = benchmark_params(n, a, b, bits); q.update_all_curves()
        small = synth.make_volume(n, a, 8, bits, resample=q.resampleCurve, dispersion=q.dispersionCurve)
        raw = np.ascontiguousarray(np.tile(small, (b // 8, 1, 1)))
        d = [torch.from_numpy(raw.view(np.int16)).cuda(), torch.from_numpy(raw.view(np.int16).copy()).cuda()]
        p = OctPipeline(fft_mode=_lib.FFT_FUSED)
        assert p.initializeCuda(None, None, copy.deepcopy(q)), getattr(p, "_create_error", "")
        if os.environ.get("OCTB200_AUTOGATHER"):
            p.enface_gather_connect(p.enface_gather_init(0, 1, a * b, 0))
            p.enface_gather_auto(True, 100, int(os.environ["OCTB200_AUTOGATHER"]), 0)
        p.process_device(d[0]); p.sync()
        for i in range(5):
            p.process_device(d[i & 1])
        p.sync()
        iters = 50
        p.event_record(0)
        for i in range(iters):
            p.process_device(d[i & 1])
        p.event_record(1)
        ms = p.event_elapsed_ms(0, 1) / iters
        kms = p.time_kernel(d[1], 30)
        out = p.copy_output(0)
        ref, _, _ = orc.process(copy.deepcopy(q).__class__(**{**q.__dict__, "bscansPerBuffer": 8}), small, mean_line=p.fpn_mean_line().astype(np.float64), determine_fpn=False)
        st = stats(out[:8], ref)
        key = f"{tag}/N{n}"
        res[key] = {"chain_ms": ms, "kernel_ms": kms, "MHz": a * b / ms / 1e3, "alg_GBs": a * b * n * 4 / kms / 1e6, "p9999_vs_oracle": st["p9999"], "gt1e-4": st["frac_gt_1e-4"]}
        print(key, json.dumps(res[key]), flush=True)
        p.cleanupCuda()
    json.dump(res, open(os.path.join(OUT, f"perf_{tag}.json"), "w"), indent=1)


def cmd_estimator():
    """dispersion estimator: GPU sweep (two launches per coefficient) vs the reference's CPU path re-run per trial"""
    import time as _t
    from octproz_b200.dispersion_estimator import DispersionEstimationEngine, DispersionEstimatorParameters, PEAK_VALUE
    from oracle import estimator_oracle as eo
    n, lines, trials = 1024, 512, 100
    q = benchmark_params(n, lines, 1); q.update_all_curves()
    raw = synth.make_volume(n, lines, 1, 12, resample=q.resampleCurve, dispersion=-q.dispersionCurve).reshape(lines, n)
    p = OctPipeline(fft_mode=_lib.FFT_FUSED)
    assert p.initializeCuda(None, None, q)
    res = {}
    for center in (10, 100):
        eng = DispersionEstimationEngine(p)
        prm = DispersionEstimatorParameters(numberOfCenterAscans=center, useLinearAscans=True, numberOfAscanSamplesToIgnore=15, sharpnessMetric=PEAK_VALUE,
                                            d2start=-150.0, d2end=-50.0, d3start=-20.0, d3end=20.0, numberOfDispersionSamples=trials)
        eng.setParams(prm)
        eng.startDispersionEstimation(raw, 12, n, lines)
        t0 = _t.perf_counter(); r = eng.startDispersionEstimation(raw, 12, n, lines); t_gpu = _t.perf_counter() - t0
        off, cnt = eo.center_lines(lines, center)
        block = raw[off:off + cnt]
        rc = orc.RefCpu()

        def ref_trials(pairs):
            out = []
            for d2, d3 in pairs:
                qq = copy.copy(q); qq.d2, qq.d3 = float(d2), float(d3); qq.signalLogScaling = False; qq.ascansPerBscan = cnt
                out.append(eo.ref_metric(rc.process(qq, block, threads=1), n // 2, eo.PEAK_VALUE, 0.0, 15))
            return out
        t0 = _t.perf_counter()
        w = eo.estimate(ref_trials, dict(numberOfDispersionSamples=trials, d2start=-150.0, d2end=-50.0, d3start=-20.0, d3end=20.0))
        t_cpu = _t.perf_counter() - t0
        res[f"center{center}"] = {"gpu_s": t_gpu, "reference_cpu_s": t_cpu, "speedup": t_cpu / t_gpu, "best_gpu": [r["bestD2"], r["bestD3"]],
                                  "best_reference": [w["bestD2"], w["bestD3"]], "trials_per_coefficient": trials}
        print("estimator", center, json.dumps(res[f"center{center}"]), flush=True)
    json.dump(res, open(os.path.join(OUT, "estimator.json"), "w"), indent=1)
    p.cleanupCuda()


def cmd_randomref():
    """the seeded random configurations of tests/random_configs.py against the LIVE reference CUDA build (oracle/_ref/libref_cuda.so)
    on the same raw buffers: per configuration and mode, the parity metric of tests/util.py at the strict 1e-4 tolerance"""
    from tests.random_configs import describe, random_chain_config
    from tests.util import parity_report
    seed = 0x0C7B200 + 7
    rc = orc.RefCuda()
    res = []
    for n in (1024, 2048):
        for i in range(18):
            rng = np.random.default_rng([seed, n, i])
            q, extras = random_chain_config(rng, n)
            rc.configure(q)
            q.resampleCurve, q.dispersionCurve, q.windowCurve = rc.curves()
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
            row = {"n": n, "i": i, "config": describe(q)}
            for mname, mode in MODES.items():
                qq = copy.deepcopy(q)
                p = OctPipeline(fft_mode=mode)
                assert p.initializeCuda(None, None, qq), getattr(p, "_create_error", "")
                if ml is not None:
                    p.set_fpn_mean_line(ml)
                if extras["pp_background"] is not None:
                    qq.loadPostProcessingBackground(extras["pp_background"])
                p.octCudaPipeline(h1); p.sync()
                out = p.copy_output(0)
                p.cleanupCuda()
                row[mname] = parity_report(out, ref, q, saturated=bool(q.postProcessBackgroundRemoval), atol_abs=floor)
            res.append(row)
            print("randomref", n, i, {m: (round(row[m]["max_ratio"], 2), row[m]["frac_outside"]) for m in MODES}, flush=True)
    json.dump(res, open(os.path.join(OUT, "randomref.json"), "w"), indent=1)
    worst = max(max(r[m]["frac_outside"] for m in MODES) for r in res)
    print("randomref worst frac_outside", worst, flush=True)


def cmd_display():
    """device time of the display / aux kernels at the BASELINE size (1024 x 512 x 256): 3-D volume texture, en-face and B-scan frames,
    stand-alone floatToOutput, and the post-FFT passes of the 2048 x 1024 x 128 chain with sinusoidal correction"""
    import torch
    res = {}
    n, a, b = 1024, 512, 256
    q = benchmark_params(n, a, b); q.update_all_curves()
    small = synth.make_volume(n, a, 8, 12, resample=q.resampleCurve, dispersion=q.dispersionCurve)
    raw = np.ascontiguousarray(np.tile(small, (b // 8, 1, 1)))
    d = torch.from_numpy(raw.view(np.int16)).cuda()
    p = OctPipeline(fft_mode=_lib.FFT_FUSED)
    assert p.initializeCuda(None, None, copy.deepcopy(q))
    p.process_device(d); p.sync()
    tex = torch.zeros(n // 2 * a * b, dtype=torch.uint8, device="cuda")
    conv = torch.zeros(n // 2 * a * b, dtype=torch.int16, device="cuda")
    dB = torch.empty(n // 2 * a, dtype=torch.float32, device="cuda"); dE = torch.empty(a * b, dtype=torch.float32, device="cuda")

    def timed(name, fn, iters=20):
        fn(); p.sync()
        p.event_record(0)
        for _ in range(iters):
            fn()
        p.event_record(1)
        res[name] = p.event_elapsed_ms(0, 1) / iters * 1e3
        print("display", name, f"{res[name]:.1f} us", flush=True)
    timed("volume_u8 (64 MiB texture)", lambda: p.volume_u8(0, tex))
    timed("float_to_output stand-alone (u16)", lambda: p.float_to_output(0, conv))
    timed("enface_frame 1 frame", lambda: p.changeDisplayedEnFaceFrame(100, 1, 0, dE))
    timed("enface_frame 16-frame average", lambda: p.changeDisplayedEnFaceFrame(100, 16, 0, dE))
    timed("bscan_frame 1 frame", lambda: p.changeDisplayedBscanFrame(100, 1, 0, dB))
    p.cleanupCuda()
    json.dump(res, open(os.path.join(OUT, "display_kernels_us.json"), "w"), indent=1)


if __name__ == "__main__":
    {"estimator": cmd_estimator, "perf": cmd_perf, "quick": cmd_quick, "parity": cmd_parity, "refcuda": cmd_refcuda, "timing": cmd_timing,
     "randomref": cmd_randomref, "display": cmd_display}[sys.argv[1]]()
