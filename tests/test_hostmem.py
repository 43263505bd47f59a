"""host-buffer placement helper (octproz_b200/hostmem.py): cpulist parsing, sysfs lookup against a fake tree, affinity context"""
import os

import pytest

from octproz_b200 import hostmem


def test_parse_cpulist():
    assert hostmem.parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert hostmem.parse_cpulist("5") == {5}
    assert hostmem.parse_cpulist("\n") == set()


def test_gpu_local_cpus_from_sysfs(tmp_path, monkeypatch):
    bdf = "0000:1b:00.0"
    d = tmp_path / bdf
    d.mkdir()
    (d / "local_cpulist").write_text("0-1,4\n")
    (d / "numa_node").write_text("1\n")
    monkeypatch.setattr(hostmem, "pci_address", lambda device: bdf)
    assert hostmem.gpu_local_cpus(0, sysfs=str(tmp_path)) == ({0, 1, 4}, 1)
    monkeypatch.setattr(hostmem, "pci_address", lambda device: "0000:ff:00.0")
    assert hostmem.gpu_local_cpus(0, sysfs=str(tmp_path)) == (None, None)
    monkeypatch.setattr(hostmem, "pci_address", lambda device: None)
    assert hostmem.gpu_local_cpus(0, sysfs=str(tmp_path)) == (None, None)


@pytest.mark.skipif(not hasattr(os, "sched_setaffinity"), reason="no sched_setaffinity")
def test_local_affinity_restores():
    before = os.sched_getaffinity(0)
    one = {min(before)}
    with hostmem.local_affinity(0, cpus=one, node=0) as info:
        assert os.sched_getaffinity(0) == one
        assert info["cpus"] == 1 and info["numa_node"] == 0 and info["applied"] == (one != before)
    assert os.sched_getaffinity(0) == before
    # CPUs outside the current mask / nothing known: no-op
    with hostmem.local_affinity(0, cpus={10 ** 6}) as info:
        assert not info["applied"] and os.sched_getaffinity(0) == before
    with hostmem.local_affinity(0, cpus=set()) as info:
        assert not info["applied"]
