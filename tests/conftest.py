import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session", autouse=True)
def built_libraries():
    """make sure the in-tree libraries exist (the driver runs build() first; this covers a bare `pytest`)"""
    lib = os.path.join(ROOT, "octproz_b200", "liboctb200.so")
    orc = os.path.join(ROOT, "oracle", "liboct_oracle.so")
    if not os.path.exists(lib):
        subprocess.check_call(["make", "-s", "-j", "8", "-C", os.path.join(ROOT, "octproz_b200", "csrc")])
    if not os.path.exists(orc):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "all"])
    yield


def pytest_collection_modifyitems(config, items):
    """a bare `pytest` on a machine without a GPU skips the GPU-marked tests instead of failing them"""
    if has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device (GPU tests run with -m gpu on the B200 box)")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def pytest_sessionfinish(session, exitstatus):
    """every parity comparison of the session (tests/util.py REPORT) as one artefact: per case the worst tolerance ratio, the fraction
    outside, and the strict relative amplitude error over the bins above the median"""
    try:
        from tests import util
        if util.REPORT:
            import json
            out = os.path.join(ROOT, "gpurun_out")
            os.makedirs(out, exist_ok=True)
            extra = getattr(util, "EXTRA_REPORT", {})
            worst = {"max_frac_outside": max(r["frac_outside"] for r in util.REPORT),
                     "max_rel_above_median": max((r["max_rel_above_median"] or 0.0) for r in util.REPORT), "comparisons": len(util.REPORT)}
            json.dump({"tolerance": "|amp_a - amp_b| <= rtol * amp_b + atol_frac * median(amp_b) + atol_abs (tests/util.py)", "summary": worst,
                       "fpn_determination": extra.get("fpn_determination"), "cases": util.REPORT}, open(os.path.join(out, "parity_report.json"), "w"), indent=1)
    except Exception as e:  # noqa: BLE001
        print("parity report not written:", e)


def has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:  # noqa: BLE001
        return False
