"""Small launch sequences for ncu captures (development helper, run on the GPU box):
   ncu --set full --clock-control none --import-source on -k regex:<kernel> -s 2 -c 1 -o gpurun_out/<name> python tools/ncu_targets.py <target>
targets: fused1024 | fused2048 | generic1664 | cufft1024 | u8_1024"""
import copy
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from octproz_b200 import OctPipeline, _lib, benchmark_params, synth  # noqa: E402

T = {"fused1024": (1024, 512, 256, 12, _lib.FFT_FUSED), "fused2048": (2048, 1024, 128, 16, _lib.FFT_FUSED),
     "generic1664": (1664, 512, 256, 12, _lib.FFT_FUSED), "cufft1024": (1024, 512, 256, 12, _lib.FFT_CUFFT),
     "u8_1024": (1024, 512, 256, 8, _lib.FFT_FUSED)}
n, a, b, bits, mode = T[sys.argv[1]]
q = benchmark_params(n, a, b, bits); q.update_all_curves()
small = synth.make_volume(n, a, 8, min(bits, 12), resample=q.resampleCurve, dispersion=q.dispersionCurve).astype(synth.container_dtype(bits))
raw = np.ascontiguousarray(np.tile(small, (b // 8, 1, 1)))
p = OctPipeline(fft_mode=mode, flags=_lib.FLAG_NO_DEPENDENT_LAUNCH)      # plain launches: ncu serialises kernels anyway
assert p.initializeCuda(None, None, copy.deepcopy(q)), getattr(p, "_create_error", "")
p.octCudaPipeline(raw); p.sync()
for _ in range(4):
    p.octCudaPipeline(None)
p.sync()
p.cleanupCuda()
print("done", sys.argv[1])
