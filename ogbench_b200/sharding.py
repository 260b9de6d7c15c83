"""Trajectory-aligned shards for datasets too large (or too many) for one replica.

Each shard is a contiguous run of whole trajectories, so it is a valid dataset on its own
(`terminal_locs[-1] == size - 1` holds per shard) and transition indices, trajectory goals *and random goals* are
shard-local -- the same semantics the reference has when it cycles through a directory of .npz shards
(impls/main.py:81-93,185-199).  No row ever crosses a GPU boundary, so the hot path needs no collective.
"""

from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np


def trajectory_ends(terminals: np.ndarray, valids=None) -> np.ndarray:
    """Last row of every trajectory.  Compact datasets mark the last two rows terminal (ogbench/utils.py:71-73); the
    true end is the row whose successor starts a new trajectory, i.e. a terminal row not followed by a terminal row."""
    t = np.asarray(terminals) > 0
    nxt = np.concatenate([t[1:], [False]])
    ends = np.nonzero(t & ~nxt)[0]
    assert len(ends) > 0 and ends[-1] == len(t) - 1, 'dataset must end on a terminal row (datasets.py:188)'
    return ends


def shard_bounds(terminals: np.ndarray, world_size: int) -> List[Tuple[int, int]]:
    """[start, stop) row ranges, one per rank, each a whole number of trajectories, balanced by row count."""
    ends = trajectory_ends(terminals)
    if world_size > len(ends):
        raise ValueError(f'{world_size} shards requested but the dataset has only {len(ends)} trajectories')
    n = len(terminals)
    bounds, start = [], 0
    for rank in range(world_size):
        if rank == world_size - 1:
            stop = n
        else:
            target = (rank + 1) * n / world_size
            j = int(np.searchsorted(ends, target - 1, side='left'))
            j = min(max(j, len(bounds)), len(ends) - (world_size - rank))  # leave >= 1 trajectory for every later rank
            stop = int(ends[j]) + 1
            if stop <= start:
                stop = int(ends[np.searchsorted(ends, start, side='left')]) + 1
        bounds.append((start, stop))
        start = stop
    return bounds


def take_shard(fields: Dict[str, np.ndarray], rank: int, world_size: int) -> Dict[str, np.ndarray]:
    start, stop = shard_bounds(fields['terminals'], world_size)[rank]
    return {k: v[start:stop] for k, v in fields.items()}
