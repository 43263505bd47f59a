"""The reference's UNMODIFIED CUDA build (oracle/_ref/libref_cuda.so: cuda_code.cu compiled in place for sm_100 with --use_fast_math)
timed on this GPU on the benchmark workload -- BASELINE.md's "number to beat" on the same box.  Runs in a process of its own because
the reference keeps file-scope global state and exit()s on CUDA errors.  bench.py calls it for its `ref_cuda` key:
    python tools/ref_cuda_bench.py <workload> <steps>
Device-resident: octCudaPipeline(NULL) re-processes the buffer already in d_inputBuffer (cuda_code.cu:1400 skips the H2D copy).
End to end: octCudaPipeline(host buffer) with streaming of the converted output to the host on (floatToOutput + D2H), two alternating
registered host buffers -- the same work our `e2e` leg times."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle import oracle as orc  # noqa: E402  (baseline leg only)


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else bench.DEFAULT_WORKLOAD
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    n, a, b, bits = bench.WORKLOADS[name]
    q = bench.workload_params(name)
    rc = orc.RefCuda()
    rc.configure(q)
    raw = [bench.make_raw(q, seed_offset=0), bench.make_raw(q, seed_offset=8)]
    rc.init(raw[0], raw[1])
    rc.process(raw[0]); rc.L.refcuda_sync()
    resident = rc.time(None, None, steps, 3) / steps
    # end to end with the stream-to-host path on
    q.streamToHost = True
    rc.configure(q)
    conv = (n // 2) * a * b * 2
    s1 = np.zeros(conv, np.uint8); s2 = np.zeros(conv, np.uint8)
    rc.L.refcuda_register_streaming(s1.ctypes.data, s2.ctypes.data, conv)
    e2e = rc.time(raw[0], raw[1], steps, 3) / steps
    rc.L.refcuda_unregister_streaming = getattr(rc.L, "refcuda_unregister_streaming")
    rc.L.refcuda_unregister_streaming()
    rc.cleanup()
    ascans = a * b
    print(json.dumps({"impl": "reference CUDA build (unmodified cuda_code.cu, sm_100, --use_fast_math, cuFFT), same GPU, same synthetic buffers",
                      "workload": name, "steps": steps, "value": ascans / resident / 1e6, "unit": bench.UNIT, "ms_per_step": resident * 1e3,
                      "e2e": {"value": ascans / e2e / 1e6, "unit": bench.UNIT, "ms_per_step": e2e * 1e3, "h2d_bytes_per_step": int(raw[0].nbytes),
                              "d2h_bytes_per_step": conv, "checksum": int(s1[:4096].view(np.uint16).sum()) + int(s2[:4096].view(np.uint16).sum())},
                      "timer": "host wall clock between cudaDeviceSynchronize calls (oracle/ref_drivers/ref_cuda.cpp refcuda_time)"}), flush=True)


if __name__ == "__main__":
    main()
