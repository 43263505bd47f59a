"""Writes golden raw -> B-scan vectors from the REFERENCE'S UNMODIFIED cuda_code.cu
(oracle/_ref/libref_cuda.so: nvcc --use_fast_math -arch=sm_100, built by oracle/Makefile where /root/reference exists).
Needs a GPU, so it is run on the B200 box through gpurun:
    gpurun -- 'python tests/golden/make_golden_refcuda.py gpurun_out/golden'
and the resulting refcuda_*.npz are committed under tests/golden/.  Inputs are the seeded synthetic buffers of
octproz_b200/synth.py; the fixtures store the raw input, the three host LUTs the reference generated, the
reference's output and (for FPN cases) the mean line it determined."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from octproz_b200 import synth  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from tests.golden.cases import chain_cases  # noqa: E402

dst = sys.argv[1] if len(sys.argv) > 1 else HERE
os.makedirs(dst, exist_ok=True)
rc = orc.RefCuda()
for n in (1024, 2048):
    for name, q in chain_cases(n).items():
        rc.configure(q)
        q.resampleCurve, q.dispersionCurve, q.windowCurve = rc.curves()
        raw = synth.make_volume(n, q.ascansPerBscan, q.bscansPerBuffer, q.bitDepth, resample=q.resampleCurve, dispersion=q.dispersionCurve)
        h1 = np.ascontiguousarray(raw).copy(); h2 = h1.copy()
        rc.init(h1, h2)
        extra = {}
        if q.postProcessBackgroundRemoval:
            bg = (0.2 + 0.1 * np.cos(np.arange(n // 2) / 40.0)).astype(np.float32)
            rc.L.refcuda_set_postprocess_background(bg.ctypes.data, n // 2)
            extra["pp_background"] = bg
        rc.process(h1)
        out = rc.output(0)
        if q.fixedPatternNoiseRemoval:
            extra["mean_line"] = rc.mean_line()
        rc.cleanup()
        np.savez_compressed(os.path.join(dst, f"refcuda_N{n}_{name}.npz"), raw=raw, resample=q.resampleCurve,
                            dispersion=q.dispersionCurve, window=q.windowCurve, out=out, **extra)
        print("wrote", n, name, out.shape, float(out.min()), float(out.max()), flush=True)
