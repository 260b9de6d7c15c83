"""One-process-per-GPU plumbing (torch.distributed) for replicas and trajectory-aligned shards.

The hot path has no collective: every rank samples from its own replica or shard with its own Philox stream
(stream_id = rank).  torch.distributed is used for rendezvous, the timing barrier, and reducing per-rank timings
(max) and unit counts (sum) to one whole-job number.  Backend 'nccl' on GPUs, 'gloo' in the CPU tests.
"""

from __future__ import annotations

import os
from typing import Dict, Tuple

import numpy as np


def env_rank() -> Tuple[int, int, int]:
    return int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1')), int(os.environ.get('LOCAL_RANK', '0'))


def init(backend: str, device=None):
    """Initialise the default process group from the torchrun environment (no-op for a single process)."""
    import torch.distributed as dist

    rank, world, _ = env_rank()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', '29500')
        kwargs = {}
        if backend == 'nccl' and device is not None:
            kwargs['device_id'] = device
        dist.init_process_group(backend, rank=rank, world_size=world, **kwargs)
    return rank, world


def _parse_cpulist(text: str):
    cpus = []
    for part in text.strip().split(','):
        if not part:
            continue
        lo, _, hi = part.partition('-')
        cpus.extend(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa(device_index: int):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off (sysfs), so that the pinned host buffers of the
    host-output path are first-touched on that node and the D2H copies do not cross the inter-socket link.  Returns
    (numa_node, n_cpus) or None when the topology is not exposed (single node, container without sysfs).  Only
    narrows the affinity the process already has."""
    try:
        import torch

        props = torch.cuda.get_device_properties(device_index)
        bdf = f'{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0'
        node = int(open(f'/sys/bus/pci/devices/{bdf}/numa_node').read())
        if node < 0:
            return None
        cpus = set(_parse_cpulist(open(f'/sys/devices/system/node/node{node}/cpulist').read()))
        allowed = cpus & set(os.sched_getaffinity(0))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node, len(allowed)
    except Exception:
        return None


def barrier():
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def reduce_scalar(value: float, op: str, device='cpu') -> float:
    """max / sum of a Python float over all ranks."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op={'max': dist.ReduceOp.MAX, 'sum': dist.ReduceOp.SUM}[op])
    return float(t.item())


def whole_job_throughput(units_this_rank: float, seconds_this_rank: float, device='cpu') -> float:
    """Units all ranks processed / the slowest rank's time (the contract of bench.py's `value`)."""
    total = reduce_scalar(units_this_rank, 'sum', device)
    slowest = reduce_scalar(seconds_this_rank, 'max', device)
    return total / slowest


def shard_for_rank(fields: Dict[str, np.ndarray], rank: int, world: int) -> Dict[str, np.ndarray]:
    """This rank's trajectory-aligned shard (see ogbench_b200.sharding)."""
    from . import sharding

    return sharding.take_shard(fields, rank, world) if world > 1 else fields
