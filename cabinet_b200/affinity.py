"""Pin a rank's host threads (and therefore its first-touch / pinned host buffers) to the NUMA node of its GPU.

One process per GPU feeds its device from pinned host memory; the end-to-end evaluation rate at fp32 input is PCIe-bound
(12.6 MB per 1024^2 image), so at 8 ranks per box the uploads must not cross the socket interconnect.  Call
``bind_to_gpu_numa(local_rank)`` right after ``torch.cuda.set_device`` and BEFORE allocating pinned buffers or spawning
worker threads.  A no-op (with the reason reported) when the topology is not exposed.
"""

from __future__ import annotations

import os
from typing import Any, Dict, Set


def _parse_cpulist(text: str) -> Set[int]:
    cpus: Set[int] = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def gpu_numa_node(device_index: int) -> Dict[str, Any]:
    """-> {"pci": "dddd:bb:dd.0", "node": int (-1 unknown), "cpus": set} from sysfs."""
    import torch

    p = torch.cuda.get_device_properties(device_index)
    bus = f"{getattr(p, 'pci_domain_id', 0):04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
    info: Dict[str, Any] = {"pci": bus, "node": -1, "cpus": set()}
    try:
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            info["node"] = int(f.read().strip())
        if info["node"] >= 0:
            with open(f"/sys/devices/system/node/node{info['node']}/cpulist") as f:
                info["cpus"] = _parse_cpulist(f.read())
    except (OSError, ValueError) as e:
        info["error"] = f"{type(e).__name__}: {e}"
    return info


def bind_to_gpu_numa(device_index: int) -> Dict[str, Any]:
    """Restrict this process to the CPUs of the GPU's NUMA node.  -> report dict (JSON-serialisable)."""
    info = gpu_numa_node(device_index)
    report = {"pci": info["pci"], "node": info["node"], "bound": False}
    if "error" in info:
        report["why"] = info["error"]
        return report
    if info["node"] < 0 or not info["cpus"]:
        report["why"] = "numa node not exposed (single-node box or virtualised topology)"
        return report
    try:
        allowed = os.sched_getaffinity(0)
        target = allowed & info["cpus"]
        if not target:
            report["why"] = "no allowed CPU on the GPU's node"
            return report
        os.sched_setaffinity(0, target)
        report.update(bound=True, cpus=len(target), cpus_before=len(allowed))
    except OSError as e:
        report["why"] = f"sched_setaffinity: {e}"
    return report
