"""Host-side placement for the page-locked buffers a rank uploads from.

With one process per GPU on a two-socket box, page-locked memory lands on the NUMA node of the thread that allocates it
(first touch); a rank whose rows sit on the other socket uploads across the socket interconnect, and with eight ranks uploading at
once that link, not PCIe, bounds the H2D rate.  `bind_to_gpu_numa_node` pins the calling thread (and the threads it creates later: the
handle's submit thread, the pinned allocations' first touch) to the CPUs NVML reports as local to the GPU.  Nothing on the device
side depends on it; a box where NVML has no answer, or whose cpuset leaves no local CPU, is left alone.
"""
import os
from typing import Dict


def numa_node_count() -> int:
    try:
        return len([d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()])
    except OSError:
        return 1


def bind_to_gpu_numa_node(device_index: int) -> Dict[str, object]:
    """Returns {"applied": bool, "cpus": how many CPUs the thread may run on afterwards, "numa_nodes": ..., "why": ...}."""
    info: Dict[str, object] = {"applied": False, "numa_nodes": numa_node_count(), "cpus": len(os.sched_getaffinity(0))}
    if os.environ.get("GCL_NUMA_BIND", "1") == "0":
        info["why"] = "GCL_NUMA_BIND=0"
        return info
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(device_index).uuid)
        h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        local = {64 * w + b for w, m in enumerate(words) for b in range(64) if (int(m) >> b) & 1}
    except Exception as e:      # no NVML, no answer for this device: leave the placement to the OS
        info["why"] = f"no NVML CPU affinity ({type(e).__name__})"
        return info
    allowed = os.sched_getaffinity(0)
    target = local & allowed
    if not target:
        info["why"] = "none of the GPU-local CPUs is in this process's cpuset"
        return info
    if target == allowed:
        info["why"] = "every allowed CPU is already local to the GPU"
        return info
    try:
        os.sched_setaffinity(0, target)
    except OSError as e:
        info["why"] = f"sched_setaffinity refused ({e.errno})"
        return info
    info.update(applied=True, cpus=len(target))
    return info
