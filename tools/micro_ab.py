"""Smallest possible same-box A/B of two builds of liboctb200.so (no torch): device time of the chain re-run on the resident raw
buffer (octCudaPipeline(NULL), cuda_code.cu:1400) and a hash of the output, so that a faster variant is also shown bit-identical.
   OCTB200_LIB=<variant .so> python tools/micro_ab.py <tag> [1024|2048]"""
import hashlib
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from octproz_b200 import OctPipeline, _lib, benchmark_params, synth  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else "default"
t0 = time.time()
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
a, b, bits = (512, 256, 12) if n == 1024 else (1024, 128, 16)
q = benchmark_params(n, a, b, bits); q.update_all_curves()
small = synth.make_volume(n, a, 4, bits, resample=q.resampleCurve, dispersion=q.dispersionCurve)
raw = np.ascontiguousarray(np.tile(small, (b // 4, 1, 1)))
p = OctPipeline(fft_mode=_lib.FFT_FUSED, flags=int(os.environ.get("OCTB200_FLAGS", "0")))      # 2 = no programmatic dependent launch
assert p.initializeCuda(None, None, q), getattr(p, "_create_error", "")
p.octCudaPipeline(raw); p.sync()
if os.environ.get("OCTB200_AUTOGATHER"):          # single-rank en-face gather fused into the kernel + its consumer kernel, every buffer
    p.enface_gather_connect(p.enface_gather_init(0, 1, a * b, 0))
    p.enface_gather_auto(True, 100, 1, 0)
for _ in range(5):
    p.octCudaPipeline(None)
p.sync()
iters = 100
p.event_record(0)
for _ in range(iters):
    p.octCudaPipeline(None)
p.event_record(1)
ms = p.event_elapsed_ms(0, 1) / iters
out = p.copy_output(0)
res = {"tag": tag, "n": n, "ms_per_volume": ms, "MHz": a * b / ms / 1e3, "sha1_first_8_bscans": hashlib.sha1(out[:8].tobytes()).hexdigest(),
       "sha1_last_bscan": hashlib.sha1(out[-1].tobytes()).hexdigest(), "wall_s": round(time.time() - t0, 1)}
print("MICRO_AB", json.dumps(res), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open(f"gpurun_out/micro_ab_{tag}_N{n}.json", "w"))
p.cleanupCuda()
