"""bench.py's contract pieces that need no GPU: the reference arm (the reference's CPU path on the host cores) and the `config` object
both arms share."""
import argparse
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_both_arms_build_the_identical_config_object():
    sys.path.insert(0, ROOT)
    import bench
    args = argparse.Namespace(workload=bench.DEFAULT_WORKLOAD, mode="fused")
    a, b = bench.make_config(args, 1), bench.make_config(args, 1)
    assert a == b and a["samples_per_ascan"] == 1024 and a["ascans_per_bscan"] == 512 and a["bscans_per_buffer"] == 256 and a["bit_depth"] == 12
    assert "l2" in a and "workload" in a and "model" not in a
    assert bench.make_config(args, 8)["parallelism"] == "b-scan sharding x8"


def test_reference_arm_ignores_omp_num_threads_and_times_the_full_buffer():
    """torch.distributed.run exports OMP_NUM_THREADS=1; round 1's reference arm obeyed it and ran the CPU path on one thread"""
    from oracle import oracle as orc
    if not orc.have_ref("libref_cpu.so"):
        pytest.skip("oracle/_ref/libref_cpu.so not built")
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--cpu-bscans", "16"],
                       capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.strip().splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    cores = len(os.sched_getaffinity(0))
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["unit"].startswith("MHz")
    assert d["cpu_baseline"]["kind"] == "reference" and (d["cpu_baseline"]["cores"] > 1 or cores == 1), d["cpu_baseline"]
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["gpu_launches"] == 0
    assert d["config"]["bscans_per_buffer"] == 256 and "16 B-scans" in d["cpu_baseline"]["sample"]


def test_non_zero_ranks_of_the_reference_arm_exit_without_work():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert r.returncode == 0 and r.stdout.strip() == ""
