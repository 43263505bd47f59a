"""The multi-GPU en-face gather protocol (octproz_b200/csrc/oct_device.cuh GatherDev) as a CPU model with
std::atomic release / acquire in place of the system-scope PTX (tests/host/gather_protocol_model.cpp): dead-lock freedom and frame
integrity at 8 and 16 ranks -- world sizes the GPU suite (1 to 4 ranks) does not reach -- with one rank running ahead and one
dawdling, the happens-before chain of the plain frame memory under ThreadSanitizer, and a negative control (no acknowledgement wait,
consumers off the compute stream: torn frames must be DETECTED).  TEST-ONLY; the device code is exercised by tests/test_gpu_multi.py."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "host", "gather_protocol_model.cpp")


def build(exe, extra=()):
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-Werror", "-pthread", *extra, SRC, "-o", exe])


@pytest.fixture(scope="module")
def model(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("gather_model") / "gather_protocol_model")
    build(exe)
    return exe


@pytest.mark.parametrize("world,steps,mode", [(8, 2000, "inorder"), (16, 800, "inorder"), (2, 3000, "inorder"), (8, 2000, "async")])
def test_no_deadlock_and_only_complete_frames(model, world, steps, mode):
    r = subprocess.run([model, str(world), str(steps), mode], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "torn_frames 0 timeouts 0" in r.stdout


def test_negative_control_without_flow_control_tears_frames(model):
    """consumers off the compute stream and no acknowledgement wait: a producer two steps ahead overwrites what a slow consumer reads"""
    r = subprocess.run([model, "8", "2000", "async-noflow"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 2, r.stdout + r.stderr


@pytest.mark.parametrize("mode", ["inorder", "async"])
def test_frame_memory_is_ordered_by_the_flags_under_thread_sanitizer(tmp_path, mode):
    exe = str(tmp_path / "gather_protocol_model_tsan")
    try:
        build(exe, extra=("-fsanitize=thread", "-g", "-O1"))
    except subprocess.CalledProcessError:
        pytest.skip("ThreadSanitizer runtime not available")
    r = subprocess.run([exe, "8", "300", mode], capture_output=True, text=True, timeout=600)
    if "FATAL: ThreadSanitizer" in r.stderr and "unexpected memory mapping" in r.stderr:
        pytest.skip("ThreadSanitizer cannot run in this container")
    assert r.returncode == 0 and "WARNING: ThreadSanitizer" not in r.stderr, r.stderr[-2000:]
