"""Host-side placement for the H2D / D2H legs of the path (the reference pins the plugin's acquisition buffers and its streaming
buffers where they happen to be, cuda_code.cu:661,1135-1136).

On a two-socket host a pinned buffer on the far NUMA node costs PCIe bandwidth (every DMA crosses the socket interconnect).  Linux
places pages on the node of the thread that first touches them (and cudaHostAlloc / cudaHostRegister touch them from the calling
thread), so running the allocating thread on the GPU's local CPUs is enough: `gpu_local_cpus` reads them from sysfs
(/sys/bus/pci/devices/<bdf>/local_cpulist) and `local_affinity` is a context manager around the allocation (and, if wanted, the
acquisition loop).  Everything degrades to a no-op when sysfs, the PCI address or sched_setaffinity are unavailable."""
from __future__ import annotations

import os


def parse_cpulist(text: str) -> set[int]:
    """'0-3,8,10-11' -> {0, 1, 2, 3, 8, 10, 11} (the kernel's cpulist format)"""
    cpus: set[int] = set()
    for part in text.strip().split(","):
        part = part.strip()
        if not part:
            continue
        if "-" in part:
            lo, hi = part.split("-", 1)
            cpus.update(range(int(lo), int(hi) + 1))
        else:
            cpus.add(int(part))
    return cpus


def pci_address(device: int) -> str | None:
    """'dddd:bb:dd.0' of CUDA device `device` (as torch numbers them, i.e. after CUDA_VISIBLE_DEVICES)"""
    try:
        import torch
        pr = torch.cuda.get_device_properties(device)
        return f"{int(pr.pci_domain_id):04x}:{int(pr.pci_bus_id):02x}:{int(pr.pci_device_id):02x}.0"
    except Exception:  # noqa: BLE001
        return None


def gpu_local_cpus(device: int, sysfs: str = "/sys/bus/pci/devices") -> tuple[set[int] | None, int | None]:
    """(CPUs local to the GPU's PCIe root, NUMA node) or (None, None) when unknown"""
    bdf = pci_address(device)
    if bdf is None:
        return None, None
    try:
        cpus = parse_cpulist(open(os.path.join(sysfs, bdf, "local_cpulist")).read())
    except Exception:  # noqa: BLE001
        return None, None
    node = None
    try:
        node = int(open(os.path.join(sysfs, bdf, "numa_node")).read().strip())
    except Exception:  # noqa: BLE001
        pass
    return (cpus or None), node


class local_affinity:
    """with local_affinity(device) as info: ...  -- the calling thread (and threads it starts meanwhile) run on the GPU-local CPUs;
    the previous affinity is restored on exit.  info = {"numa_node", "cpus", "applied"}."""

    def __init__(self, device: int, cpus: set[int] | None = None, node: int | None = None):
        if cpus is None:
            cpus, node = gpu_local_cpus(device)
        self.cpus, self.node, self.prev, self.info = cpus, node, None, {"numa_node": node, "cpus": 0, "applied": False}

    def __enter__(self):
        try:
            prev = os.sched_getaffinity(0)
            want = (self.cpus & prev) if self.cpus else set()
            if want and want != prev:
                os.sched_setaffinity(0, want)
                self.prev = prev
                self.info.update(cpus=len(want), applied=True)
            elif want:
                self.info.update(cpus=len(want))
        except Exception:  # noqa: BLE001
            self.prev = None
        return self.info

    def __exit__(self, *exc):
        if self.prev is not None:
            try:
                os.sched_setaffinity(0, self.prev)
            except Exception:  # noqa: BLE001
                pass
        return False
