"""include/octb200_host.hpp -- the host side above the C ABI in the reference's own language (C++17, no Qt): AcquisitionBuffer,
AcquisitionSystem, VirtualOCTSystem, OctPipeline (the kernels.h names), Processing::slot_start.  The handshake is exercised on
the CPU against a stand-in pipeline (tests/host/host_mirror_test.cpp, also clean under -fsanitize=thread); the example program
examples/replay_main.cpp must build against the library and fail loudly without a GPU."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "octproz_b200")


def build(src, exe, extra=()):
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-Werror", "-pthread", *extra, "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, src), "-o", exe, "-L" + LIBDIR, "-loctb200", "-Wl,-rpath," + LIBDIR])


def test_handshake_against_a_stand_in_pipeline(tmp_path):
    exe = str(tmp_path / "host_mirror_test")
    build("tests/host/host_mirror_test.cpp", exe)
    r = subprocess.run([exe, str(tmp_path / "replay.raw"), str(tmp_path / "window.f32")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert "host mirror ok" in r.stdout
    # the C++ and the Python mirror of the estimator path's window agree (to the last ulp or one: cosf vs numpy's float32 cos)
    import numpy as np
    from octproz_b200.dispersion_estimator import cpu_path_window
    w = np.fromfile(str(tmp_path / "window.f32"), np.float32)
    assert w.shape == (1024,) and np.allclose(w, cpu_path_window(1024), rtol=0, atol=2e-7)


@pytest.mark.skipif(shutil.which("g++") is None, reason="no g++")
def test_handshake_is_race_free_under_thread_sanitizer(tmp_path):
    exe = str(tmp_path / "host_mirror_tsan")
    try:
        build("tests/host/host_mirror_test.cpp", exe, extra=("-fsanitize=thread", "-g", "-O1"))
    except subprocess.CalledProcessError:
        pytest.skip("ThreadSanitizer runtime not available")
    r = subprocess.run([exe, str(tmp_path / "replay.raw")], capture_output=True, text=True, timeout=300)
    if "FATAL: ThreadSanitizer" in r.stderr and "unexpected memory mapping" in r.stderr:
        pytest.skip("ThreadSanitizer cannot run in this container")
    assert r.returncode == 0 and "WARNING: ThreadSanitizer" not in r.stderr, r.stderr[-2000:]


def test_example_builds_and_fails_loudly_without_a_gpu(tmp_path):
    from tests.conftest import has_gpu
    exe = str(tmp_path / "replay")
    build("examples/replay_main.cpp", exe)
    raw = tmp_path / "tiny.raw"
    raw.write_bytes(bytes(2 * 1024 * 4 * 2 * 2))
    r = subprocess.run([exe, str(raw), "1024", "4", "2", "12", "3"], capture_output=True, text=True, timeout=120)
    if has_gpu():
        assert r.returncode == 0, r.stderr
    else:
        assert r.returncode == 3 and "no CUDA device" in r.stderr       # never computes on the host
