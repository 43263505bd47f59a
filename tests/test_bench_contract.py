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


def test_committed_bench_lines_carry_every_key_of_the_contract():
    """the lines the last GPU visit produced (profiles/r02z_bench_{ours,ref}.json): shape of the contract, both arms on one config"""
    ours = json.load(open(os.path.join(ROOT, "profiles", "r02z_bench_ours.json")))
    ref = json.load(open(os.path.join(ROOT, "profiles", "r02z_bench_ref.json")))
    for d in (ours, ref):
        for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                  "dtype", "data", "config", "e2e", "gpu_launches", "cpu_baseline"):
            assert k in d, k
        assert d["vs_baseline"] is None and d["data"] == "synthetic" and d["scaling"] == "weak" and "model" not in d["config"]
        assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"])
        assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"])
    assert ours["config"] == ref["config"] and ours["metric"] == ref["metric"] and ours["unit"] == ref["unit"]
    assert ref["impl"] == "reference" and ref["e2e"]["value"] == ref["value"] and ref["gpu_launches"] == 0
    r = ours["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and r["bound"] == "hbm" and r["unit"] == "GB/s"
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    # achieved = algorithmic bytes per launch / the kernel's own launch time; the step of the resident region IS that one launch
    assert abs(r["achieved"] - r["algorithmic_bytes_per_launch"] / (r["kernel_ms"] * 1e-3) / 1e9) < 1e-6 * r["achieved"]
    assert r["algorithmic_bytes_per_launch"] == 4 * 1024 * 512 * 256 and r["traffic"] is not None
    assert ours["gpu_launches"] == ours["steps"] and abs(ours["ms_per_step"] - r["kernel_ms"]) < 0.05 * r["kernel_ms"]
    # value is the whole-job rate of the timed region, e2e moves host buffers both ways inside its region
    assert abs(ours["value"] - 512 * 256 / (ours["ms_per_step"] * 1e3)) < 1e-6 * ours["value"]
    assert ours["e2e"]["h2d_bytes_per_step"] == 2 * 1024 * 512 * 256 and ours["e2e"]["d2h_bytes_per_step"] == 2 * 512 * 512 * 256
    assert ours["e2e"]["value"] < ours["value"]
    c = ours["clocks"]
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(c) and not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
