import sys, json, copy, numpy as np, torch
sys.path.insert(0, '.')
from octproz_b200 import OctPipeline, _lib, benchmark_params, synth
res = {}
for (n, a, b, bits) in ((1664, 512, 256, 12), (512, 512, 256, 12), (4096, 512, 64, 12), (1024, 512, 256, 8), (1024, 512, 256, 32), (2048, 1024, 128, 8)):
    q = benchmark_params(n, a, b, bits); q.update_all_curves()
    small = synth.make_volume(n, a, 8, min(bits, 12), resample=q.resampleCurve, dispersion=q.dispersionCurve).astype(synth.container_dtype(bits))
    raw = np.ascontiguousarray(np.tile(small, (b // 8, 1, 1)))
    d = torch.from_numpy(raw.view(np.uint8)).cuda()
    for mname, mode in (("fused", _lib.FFT_FUSED), ("cufft", _lib.FFT_CUFFT)):
        p = OctPipeline(fft_mode=mode)
        assert p.initializeCuda(None, None, copy.deepcopy(q)), getattr(p, "_create_error", "")
        p.process_device(d); p.sync()
        for _ in range(3): p.process_device(d)
        p.sync()
        p.event_record(0)
        for _ in range(20): p.process_device(d)
        p.event_record(1)
        ms = p.event_elapsed_ms(0, 1) / 20
        res[f"{n}x{a}x{b}-{bits}bit/{mname}"] = {"ms": ms, "MHz": a * b / ms / 1e3, "Gsample_s": n * a * b / ms / 1e6}
        print(f"GENPERF {n}x{a}x{b}-{bits}bit {mname}: {ms:.3f} ms  {a*b/ms/1e3:.1f} MHz  {n*a*b/ms/1e6:.1f} Gsample/s", flush=True)
        p.cleanupCuda()
json.dump(res, open("gpurun_out/generic_perf.json", "w"), indent=1)
