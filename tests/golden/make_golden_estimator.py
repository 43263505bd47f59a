"""Golden vectors of the dispersion-estimator path, produced by the REFERENCE's own code compiled in place
(oracle/_ref/libref_cpu.so = octprocessor/processor.tpp, oracle/_ref/libref_metric.so = ascanmetriccalculator.cpp).
Container only (needs /root/reference to build those libraries):  python tests/golden/make_golden_estimator.py
Writes tests/golden/estimator.npz."""
import copy
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from octproz_b200 import benchmark_params, synth  # noqa: E402
from oracle import estimator_oracle as eo  # noqa: E402
from oracle import oracle as orc  # noqa: E402

N, LINES = 1024, 6
EST = dict(numberOfDispersionSamples=16, d2start=-160.0, d2end=0.0, d3start=-40.0, d3end=40.0, autoCalcD1=True)


def main():
    q = benchmark_params(N, LINES, 1); q.update_all_curves()
    # the sample dispersion has the sign that puts the compensated (sharp) image into the kept lower half of the inverse transform
    raw = synth.make_volume(N, LINES, 1, 12, resample=q.resampleCurve, dispersion=-q.dispersionCurve).reshape(LINES, N)
    rc = orc.RefCpu()

    def ref_process(d2, d3, log):
        qq = copy.copy(q); qq.d2, qq.d3 = float(d2), float(d3); qq.signalLogScaling = bool(log)
        return rc.process(qq, raw, threads=1).reshape(LINES, N // 2)

    trials = np.array([(-96.625, -0.375), (0.0, 0.0), (-140.0, 20.0), (-60.0, -30.0)], np.float32)
    out = {"raw": raw, "trials": trials, "c": np.array([q.c0, q.c1, q.c2, q.c3], np.float32), "d01": np.array([q.d0, q.d1], np.float32),
           "log": np.array([q.signalGrayscaleMin, q.signalGrayscaleMax, q.signalMultiplicator, q.signalAddend], np.float32)}
    for log in (0, 1):
        a = np.stack([ref_process(d2, d3, log) for d2, d3 in trials])
        out[f"ascans_log{log}"] = a
        thr = 0.4 if log else 40.0
        out[f"thr_log{log}"] = np.float32(thr)
        out[f"metrics_log{log}"] = np.array([[eo.ref_metric(a[k], N // 2, m, thr, ig) for k in range(len(trials))]
                                             for m in range(4) for ig in (0, 15)], np.float32).reshape(4, 2, len(trials))
    # the whole search, every trial through the reference's CPU path and the reference's metric
    for log in (0, 1):
        thr = float(out[f"thr_log{log}"])
        res = eo.estimate(lambda pairs: [eo.ref_metric(ref_process(d2, d3, log), N // 2, eo.PEAK_VALUE, thr, 15) for d2, d3 in pairs], EST)
        out[f"search_log{log}"] = np.array([res["bestD2"], res["bestD3"], res["calculatedD1"]], np.float64)
        out[f"search_metricD2_log{log}"] = np.array(res["metricD2"], np.float32)
        out[f"search_metricD3_log{log}"] = np.array(res["metricD3"], np.float32)
        print("log", log, "best", res["bestD2"], res["bestD3"])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "estimator.npz"), **out)
    print("wrote estimator.npz")


if __name__ == "__main__":
    main()
