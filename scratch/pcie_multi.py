"""Raw pinned host<->device copy bandwidth with 1..N GPUs copying at the same time (no sampler involved).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scratch/pcie_multi.py

Every rank allocates its own page-locked buffers (after binding to its GPU's NUMA node when the topology is exposed),
all ranks start together, and each times `reps` plain cudaMemcpyAsync calls of `mb` MB with CUDA events on its own
stream.  Rank 0 prints one JSON line: per-rank and summed GB/s for D2H, H2D and both directions at once.  This is the
independent measurement of the box's ceiling for the `e2e` leg of bench.py (whose D2H copies are 65 MB blocks)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from ogbench_b200 import dist_util  # noqa: E402


def main():
    rank, world, local = dist_util.env_rank()
    torch.cuda.set_device(local)
    numa = None if os.environ.get('OGB_NO_NUMA_BIND') else dist_util.bind_to_gpu_numa(local)
    dist_util.init('nccl', device=torch.device('cuda', local))
    mb = int(os.environ.get('PROBE_MB', '65'))
    reps = int(os.environ.get('PROBE_REPS', '40'))
    n = mb << 20
    dev = torch.empty(n, dtype=torch.uint8, device=f'cuda:{local}')
    dev2 = torch.empty(n, dtype=torch.uint8, device=f'cuda:{local}')
    host_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    host_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    host_in.fill_(1)                       # first touch on this rank's node
    host_out.fill_(0)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        dist_util.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        dist_util.barrier()
        return e0.elapsed_time(e1) * 1e-3

    def d2h():
        host_out.copy_(dev, non_blocking=True)

    def h2d():
        dev2.copy_(host_in, non_blocking=True)

    def both():
        with torch.cuda.stream(s1):
            host_out.copy_(dev, non_blocking=True)
        with torch.cuda.stream(s2):
            dev2.copy_(host_in, non_blocking=True)

    out = {}
    for name, fn, factor in (('d2h', d2h, 1), ('h2d', h2d, 1), ('bidir', both, 2)):
        if name == 'bidir':
            torch.cuda.synchronize()
        t = timed(fn)
        if name == 'bidir':      # events were recorded on the default stream: time the side streams through a sync instead
            import time

            dist_util.barrier()
            t0 = time.perf_counter()
            for _ in range(reps):
                fn()
            torch.cuda.synchronize()
            t = time.perf_counter() - t0
        gbs = factor * reps * n / t / 1e9
        total = dist_util.reduce_scalar(gbs, 'sum', device=f'cuda:{local}')
        lo = -dist_util.reduce_scalar(-gbs, 'max', device=f'cuda:{local}')
        out[name] = {'sum_gbs': round(total, 1), 'slowest_rank_gbs': round(lo, 1)}
    if rank == 0:
        out.update({'n_gpus': world, 'block_mb': mb, 'reps': reps, 'numa_bound': numa, 'host_cpus': os.cpu_count()})
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
