"""Host topology: keep a rank's threads and pinned buffers on its GPU's NUMA node.

On an 8-GPU HGX box the GPUs hang off two sockets (``nvidia-smi topo -m``: GPUs
0-3 on NUMA node 0, 4-7 on node 1).  A rank whose threads run on the other
socket allocates its pinned staging buffers there (first touch) and every
host->device copy crosses the socket interconnect — at 8 ranks that halves the
aggregate H2D rate.  ``bind_to_gpu`` pins the calling process to the CPUs NVML
reports as local to the GPU *before* any pinned memory is allocated.

(The reference's resource model is machines x procs, ``exec/slicemachine.go:
26-817``; inside one box the analogous placement decision is which socket a
rank's host threads and staging buffers live on.)
"""
from __future__ import annotations

import os
from typing import List, Optional

_bound: Optional[List[int]] = None


def gpu_cpu_affinity(index: int) -> List[int]:
    """CPUs local to GPU ``index`` according to NVML ([] when unknown)."""
    try:
        import pynvml
    except ImportError:
        return []
    try:
        pynvml.nvmlInit()
        h = None
        try:  # CUDA ordinals follow CUDA_VISIBLE_DEVICES, NVML's do not: match by UUID
            import torch
            uuid = str(torch.cuda.get_device_properties(index).uuid)
            h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
        except Exception:  # noqa: BLE001
            h = None
        if h is None:
            h = pynvml.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 1
        words = (ncpu + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = [w * 64 + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1]
        return [c for c in cpus if c < ncpu]
    except Exception:  # noqa: BLE001 - NVML missing/unsupported: leave affinity alone
        return []


def bind_to_gpu(index: int) -> List[int]:
    """Restrict this process to the CPUs local to GPU ``index`` (intersected with
    the CPUs it may already use).  Returns the CPU list in effect; a no-op when
    NVML has no answer or ``BIGSLICE_NO_NUMA_BIND`` is set."""
    global _bound
    if os.environ.get("BIGSLICE_NO_NUMA_BIND") or not hasattr(os, "sched_setaffinity"):
        return []
    local = gpu_cpu_affinity(index)
    if not local:
        return []
    try:
        allowed = os.sched_getaffinity(0)
        cpus = sorted(set(local) & set(allowed))
        if not cpus:
            return []
        os.sched_setaffinity(0, cpus)
        _bound = cpus
        return cpus
    except OSError:
        return []


def bound_cpus() -> Optional[List[int]]:
    return _bound
