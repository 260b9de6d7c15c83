"""The N>1 host path on CPU: two gloo ranks shard a dataset at trajectory boundaries, draw independent Philox
streams, and reduce timings the way bench.py does.  No GPU involved (the data path has no collective)."""

import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, result_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    from ogbench_b200 import dist_util, sharding
    from oracle import philox_np
    from oracle.replay_oracle import DrawsSource, OracleSampler
    from tests.golden.make_golden import cfg, ragged, toy_fields

    r, w = dist_util.init('gloo')
    assert (r, w) == (rank, world)
    lengths = ragged(21, 41, 2, 70)
    fields = toy_fields(21, lengths, (3,), 2, np.float32)
    fields['observations'][:, 0] = np.arange(len(fields['terminals']))       # global row id
    shard = dist_util.shard_for_rank(fields, rank, world)
    bounds = sharding.shard_bounds(fields['terminals'], world)

    # every rank's shard is a valid dataset and the shards tile the dataset (checked with a collective)
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([len(shard['terminals'])]))
    assert sum(int(s) for s in sizes) == len(fields['terminals'])
    assert int(sizes[rank]) == bounds[rank][1] - bounds[rank][0]

    # shard-local sampling with this rank's Philox stream: all rows (transition, next, goals) stay inside the shard
    config = cfg()
    sampler = OracleSampler(shard, config, 'gc')
    draws, _ = philox_np.philox_draws(seed=99, stream=rank, batch_index=0, batch_size=256, n_choices=len(sampler.valid_table),
                                      goal_sets=[(0, True, 0.99, False), (2, False, 0.99, False)], aug=True, p_aug=0.0)
    batch = sampler.sample(256, source=DrawsSource(draws))
    lo, hi = bounds[rank]
    for key in ('observations', 'next_observations', 'value_goals', 'actor_goals'):
        ids = batch[key][:, 0]
        assert (ids >= lo).all() and (ids < hi).all(), key

    # streams of different ranks differ
    mine = torch.from_numpy(draws.idx_pos.copy())
    gathered = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    assert not torch.equal(gathered[0], gathered[1])

    # bench.py's reduction: whole-job units / slowest rank
    value = dist_util.whole_job_throughput(units_this_rank=1000.0 * (rank + 1), seconds_this_rank=1.0 + rank)
    assert abs(value - (1000.0 + 2000.0) / 2.0) < 1e-9
    dist_util.barrier()
    open(os.path.join(result_dir, f'ok{rank}'), 'w').write('ok')
    dist.destroy_process_group()


def test_two_rank_gloo(tmp_path):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f'ok{r}').exists() for r in range(world))
