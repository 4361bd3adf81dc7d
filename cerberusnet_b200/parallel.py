"""Multi-GPU plumbing for the hot path.

The op is per-sample (every output element depends on one batch item only: blockIdx.x = n in the
reference, correlation_cuda_kernel.cu:35), so it shards over image pairs / batch items with no
exchange step: one process per GPU, ``torch.distributed`` only for the barrier and for reducing
timings (bench.py) or, in training, DDP's gradient all-reduce of the surrounding model.
"""
from __future__ import annotations

import os
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous [begin, end) slice of ``total`` independent units (image pairs, batch items)
    owned by ``rank``; the first ``total % world`` ranks take one extra unit."""
    if world < 1 or not 0 <= rank < world or total < 0:
        raise ValueError(f"bad shard request total={total} world={world} rank={rank}")
    base, extra = divmod(total, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def max_over_ranks(value: float, device=None) -> float:
    """Device-timed durations are reported as the max over ranks (never wall clock)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device=None) -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def aggregate_throughput(units_this_rank: float, elapsed_ms_this_rank: float, device=None) -> float:
    """Whole-job throughput: units processed by all ranks / max-over-ranks time (units per second)."""
    total = sum_over_ranks(units_this_rank, device)
    ms = max_over_ranks(elapsed_ms_this_rank, device)
    return total / (ms * 1e-3)


def gpu_local_cpus(device_index: int) -> Optional[List[int]]:
    """CPUs of the NUMA node the GPU hangs off (NVML's ideal-affinity mask; the device is looked up by UUID so
    CUDA_VISIBLE_DEVICES re-numbering does not matter).  None when NVML cannot tell."""
    try:
        import pynvml
        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(device_index).uuid)
        if not uuid.startswith("GPU-"):
            uuid = "GPU-" + uuid
        h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode() if hasattr(uuid, "encode") else uuid)
        ncpu = os.cpu_count() or 64
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [64 * w + b for w, m in enumerate(words) for b in range(64) if (int(m) >> b) & 1]
        return cpus or None
    except Exception:  # noqa: BLE001 -- no NVML, container without topology: leave the affinity alone
        return None


def bind_to_gpu_numa(device_index: int, local_rank: int = 0, local_world: int = 1) -> dict:
    """Pin this process to its GPU's NUMA-local cores BEFORE it allocates pinned host buffers, so the staging
    memory is first-touched on the node the PCIe root port belongs to.  With several ranks per node the local
    cores are split between them.  (SCALE_r01: eight ranks all on NUMA node 0 got 17 GB/s per GPU instead of
    71 GB/s.)  Returns what was done, for the bench record."""
    node = gpu_numa_node(device_index)
    info = {"bound": False, "cpus": None, "numa_node": node, "mempolicy_preferred": prefer_numa_memory(node)}
    cpus = gpu_local_cpus(device_index)
    if not cpus:
        return info
    try:
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed:
            return info
        share = split_cpus(allowed, local_rank, local_world, device_index)
        os.sched_setaffinity(0, share)
        info.update(bound=True, cpus=f"{share[0]}-{share[-1]} ({len(share)} cores)", numa_cpus=len(allowed))
    except OSError:
        pass
    return info


def split_cpus(cpus: List[int], local_rank: int, local_world: int, device_index: int = 0) -> List[int]:
    """The slice of a NUMA node's cores one rank keeps: ranks whose GPUs share the node take disjoint,
    equal slices (at least one core each); the slice index is the GPU's position among its node-mates,
    approximated by device_index modulo the number of ranks per node."""
    per_node = max(1, min(local_world, 4))            # up to 4 GPUs per NUMA node on an 8-GPU HGX board
    k = max(1, len(cpus) // per_node)
    i = device_index % per_node
    share = cpus[i * k:(i + 1) * k]
    return share or cpus


def gpu_numa_node(device_index: int) -> Optional[int]:
    """NUMA node of the GPU's PCIe root port (sysfs); None if the platform does not say."""
    try:
        pr = torch.cuda.get_device_properties(device_index)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        return node if node >= 0 else None
    except Exception:  # noqa: BLE001
        return None


def prefer_numa_memory(node: Optional[int]) -> bool:
    """set_mempolicy(MPOL_PREFERRED, {node}) for this thread: pages it first-touches afterwards (the pinned staging
    arenas) come from the GPU's node even when the cpuset keeps the process on another node's cores."""
    if node is None:
        return False
    try:
        import ctypes
        libc = ctypes.CDLL(None, use_errno=True)
        mask = ctypes.c_ulong(1 << node)
        MPOL_PREFERRED, SYS_set_mempolicy = 1, 238   # x86_64
        rc = libc.syscall(SYS_set_mempolicy, MPOL_PREFERRED, ctypes.byref(mask), ctypes.c_ulong(64))
        return rc == 0
    except Exception:  # noqa: BLE001
        return False
