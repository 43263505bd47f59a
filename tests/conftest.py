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


def has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:  # noqa: BLE001
        return False
