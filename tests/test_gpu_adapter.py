"""Drop-in check: the reference's own entry-point names (kernels.h:63-84: initializeCuda / octCudaPipeline /
cleanupCuda / cuda_registerStreamingBuffers ...) implemented by integration/octproz_kernels_adapter.cpp on top of
liboctb200.so, driven by the same headless harness that drives the reference's cuda_code.cu, with the reference's own
OctAlgorithmParameters / Polynomial / WindowFunction code generating the curves (oracle/_ref/libadapter_api.so)."""
import copy
import ctypes as C
import glob
import os

import numpy as np
import pytest

from oracle import oracle as orc
from tests.golden.cases import chain_cases
from tests.util import assert_parity

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not orc.have_ref("libadapter_api.so"), reason="oracle/_ref/libadapter_api.so not built")]
GOLD_DIR = os.path.join(os.path.dirname(__file__), "golden")
GOLDEN = sorted(glob.glob(os.path.join(GOLD_DIR, "refcuda_*.npz")))


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[8:-4] for p in GOLDEN])
def test_reference_api_on_our_library_matches_reference_cuda(path):
    name = os.path.basename(path)[8:-4]
    n = int(name.split("_")[0][1:]); case = name.split("_", 1)[1]
    g = np.load(path)
    q = copy.deepcopy(chain_cases(n)[case])
    ad = orc.RefCuda("libadapter_api.so")
    ad.configure(q)
    r, d, w = ad.curves()
    assert np.array_equal(r, g["resample"]) and np.array_equal(d, g["dispersion"]) and np.array_equal(w, g["window"])
    h1 = np.ascontiguousarray(g["raw"]).copy(); h2 = h1.copy()
    ad.init(h1, h2)
    if "pp_background" in g.files:
        bg = np.ascontiguousarray(g["pp_background"])
        ad.L.refcuda_set_postprocess_background(bg.ctypes.data, n // 2)
    ad.process(h1)
    out = ad.output(0)
    if "mean_line" in g.files:
        # our own FPN determination (the adapter determines the line itself): every bin picks the reference's segment or one that is
        # indistinguishable from it within the reference's own fp32 error bound (tests/test_gpu_reference_full.py classify_fpn_bins);
        # element-wise output parity over the bins with the same segment
        from tests.test_gpu_reference_full import classify_fpn_bins
        ml = ad.mean_line()
        h = n // 2
        stats = np.empty((9, h, 4), np.float32); seg = C.c_int()
        assert ad.L.octb200_adapter_get_fpn_segment_stats(stats.ctypes.data_as(C.c_void_p), h, C.byref(seg)) == 0
        c = classify_fpn_bins(stats, int(seg.value), ml, g["mean_line"])
        assert c["identification_error"] < 5e-4 and np.all(c["gap_over_bound"] <= 1.0), (name, c["differ"], c["gap_over_bound"], c["identification_error"])
        sel = c["same"]
        assert sel.mean() >= 0.9, (name, sel.mean())
        out, gold = out[..., sel], g["out"][..., sel]
    else:
        gold = g["out"]
    ad.cleanup()
    floor = 4e-6 * float(np.abs(g["mean_line"]).max()) if "mean_line" in g.files else 0.0
    assert_parity(out, gold, q, saturated=bool(q.postProcessBackgroundRemoval), max_frac_outside=1e-4, atol_abs=floor, what=name)


def test_streaming_callbacks_through_the_adapter():
    n = 1024
    q = copy.deepcopy(chain_cases(n)["benchmark_nofpn"]); q.streamToHost = True
    g = np.load(os.path.join(GOLD_DIR, "refcuda_N1024_benchmark_nofpn.npz"))
    ad = orc.RefCuda("libadapter_api.so")
    ad.configure(q)
    h1 = np.ascontiguousarray(g["raw"]).copy(); h2 = h1.copy()
    ad.init(h1, h2)
    s1 = np.zeros(g["out"].shape, np.uint16); s2 = np.zeros_like(s1)
    ad.L.refcuda_register_streaming(s1.ctypes.data, s2.ctypes.data, s1.nbytes)
    c0 = (C.c_int(), C.c_int(), C.c_int()); ad.L.refcuda_callback_counts(*[C.byref(c) for c in c0])
    ad.process(h1); ad.process(h2)
    ad.L.refcuda_sync()
    ad.L.octb200_adapter_sync()
    c1 = (C.c_int(), C.c_int(), C.c_int()); ad.L.refcuda_callback_counts(*[C.byref(c) for c in c1])
    assert c1[0].value - c0[0].value == 2                      # Gpu2HostNotifier::dh2StreamingCallback fired once per buffer
    out = ad.output(0)
    assert np.array_equal(s1, orc.float_to_output(out, 12)) or np.array_equal(s2, orc.float_to_output(out, 12))
    ad.L.refcuda_unregister_streaming()
    ad.cleanup()
