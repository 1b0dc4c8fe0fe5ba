"""NUMA placement for one-process-per-GPU jobs.

On a two-socket HGX node half of the GPUs hang off each socket.  A rank whose threads run on -- and whose pinned /
page-cache memory therefore lives on -- the OTHER socket feeds its GPU across the inter-socket link, which all such
ranks share: measured on 8 x B200, the shard loader's aggregate host-to-device rate saturated at ~45 GB/s (11 M
activations/s end to end at 8 ranks against 28 M for the kernels) however the copies were issued.
`bind_process_to_gpu()` pins the calling process (and every thread it starts later: the loader's I/O and feeder
threads, NCCL's proxy threads) to the CPUs that are local to its GPU, so that first-touch allocation puts staging
buffers, tmpfs shard pages it writes, and torch's pinned tensors on the GPU's own memory node.

saev itself is single-GPU and has no counterpart; torchrun does not do this either.
"""

from __future__ import annotations

import os
import pathlib


def _parse_cpulist(text: str) -> list[int]:
    cpus: list[int] = []
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.extend(range(int(a), int(b) + 1))
        else:
            cpus.append(int(part))
    return cpus


def gpu_local_cpus(device_index: int) -> tuple[list[int], str]:
    """(CPUs local to the GPU, how they were found); an empty list if the topology cannot be read."""
    try:
        import pynvml

        pynvml.nvmlInit()
        try:
            # honour CUDA_VISIBLE_DEVICES: map the CUDA ordinal to the NVML device through the PCI bus id
            import torch

            bus = torch.cuda.get_device_properties(device_index)
            bus_id = f"{bus.pci_domain_id:08x}:{bus.pci_bus_id:02x}:{bus.pci_device_id:02x}.0"
            h = pynvml.nvmlDeviceGetHandleByPciBusId(bus_id.encode())
        except Exception:  # noqa: BLE001
            h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            bus_id = pynvml.nvmlDeviceGetPciInfo(h).busId
            bus_id = bus_id.decode() if isinstance(bus_id, bytes) else bus_id
        n_cpu = os.cpu_count() or 1
        try:
            words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
            cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1]
            if cpus:
                return [c for c in cpus if c < n_cpu], "nvml cpu affinity"
        except Exception:  # noqa: BLE001
            pass
        dev = pathlib.Path("/sys/bus/pci/devices") / bus_id.lower()[-12:]
        node = int((dev / "numa_node").read_text())
        if node >= 0:
            return _parse_cpulist((pathlib.Path("/sys/devices/system/node") / f"node{node}" / "cpulist").read_text()), f"sysfs numa node {node}"
    except Exception:  # noqa: BLE001
        pass
    return [], "unknown"


def bind_process_to_gpu(device_index: int) -> dict:
    """Pin this process to the CPUs local to GPU `device_index` (intersected with the CPUs it may already use).
    Returns {"cpus": n, "how": ...}; leaves the affinity alone (cpus = 0) when the topology cannot be read or the
    intersection is empty."""
    cpus, how = gpu_local_cpus(device_index)
    try:
        allowed = os.sched_getaffinity(0)
        want = sorted(set(cpus) & allowed)
        if want and len(want) < len(allowed):
            os.sched_setaffinity(0, want)
            return {"cpus": len(want), "how": how}
        return {"cpus": 0, "how": how + (" (already local / single node)" if want else " (no usable cpus)")}
    except (AttributeError, OSError) as e:
        return {"cpus": 0, "how": f"{how}; sched_setaffinity failed: {e}"}
