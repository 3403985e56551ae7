"""Host-side placement for the host-buffer pipelines: run the process that feeds a GPU on the NUMA node that GPU hangs
off, so that its pinned staging buffers (first touch) and its copy-issuing thread are local to the PCIe root complex.

Every rank of a multi-GPU job moves its own poses over its own PCIe link; what the ranks share is host memory
bandwidth and, on a multi-socket host, the inter-socket fabric.  Binding is a no-op on a single-node host (all GPUs
report NUMA node 0, as the 8-GPU boxes of this pool do): there the shared ceiling is host DRAM itself, measured by
``scripts/experiments/exp_pcie_nrank.py``."""
from __future__ import annotations

import glob
import os


def _cpulist(text: str) -> list[int]:
    cpus: list[int] = []
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.extend(range(int(lo), int(hi or lo) + 1))
    return cpus


def gpu_numa_node(device: int) -> int | None:
    """NUMA node of the PCI function of CUDA device ``device`` (sysfs), or None when the platform does not say."""
    try:
        import torch

        bdf = torch.cuda.get_device_properties(device)
        addr = f"{bdf.pci_domain_id:04x}:{bdf.pci_bus_id:02x}:{bdf.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{addr}/numa_node") as f:
            node = int(f.read().strip())
        return node if node >= 0 else None
    except Exception:
        return None


def bind_to_gpu_numa(device: int) -> dict:
    """Restrict this process to the CPUs of the GPU's NUMA node (``os.sched_setaffinity``); pinned buffers allocated
    afterwards are first-touched there.  Returns what was done, for the benchmark record."""
    nodes = sorted(glob.glob("/sys/devices/system/node/node[0-9]*"))
    info = {"numa_nodes": len(nodes), "gpu_numa_node": gpu_numa_node(device), "bound": False}
    try:
        info["cpus_before"] = len(os.sched_getaffinity(0))
        if len(nodes) > 1 and info["gpu_numa_node"] is not None:
            with open(f"/sys/devices/system/node/node{info['gpu_numa_node']}/cpulist") as f:
                cpus = set(_cpulist(f.read())) & os.sched_getaffinity(0)
            if cpus:
                os.sched_setaffinity(0, cpus)
                info["bound"] = True
                info["cpus_after"] = len(cpus)
    except Exception as e:   # placement is an optimisation, never a requirement
        info["error"] = repr(e)
    return info
