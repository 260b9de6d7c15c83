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
