"""Synthetic datasets of the shapes BASELINE.json names; there is no network for the real ones.

Layout follows ogbench.load_dataset(compact_dataset=True) (ogbench/utils.py:60-73): fixed-length trajectories,
terminals on each trajectory's last two rows, valids = 0 on the last row.  Values: observations standard normal
(float32) or uniform bytes (uint8 pixels), actions uniform(-1, 1).
"""

from __future__ import annotations

import dataclasses
from typing import Dict, Optional

import numpy as np


@dataclasses.dataclass(frozen=True)
class Workload:
    key: str
    name: str
    episodes: int
    steps: int               # rows per episode
    obs_shape: tuple
    obs_dtype: str
    act_dim: int
    kind: str                # 'gc' | 'hgc'
    batch: int
    config: dict
    bytes_per_transition: int  # algorithmic bytes, SURVEY.md 8(d)
    seed: int

    @property
    def rows(self) -> int:
        return self.episodes * self.steps


_GCIQL = dict(
    discount=0.99, value_p_curgoal=0.2, value_p_trajgoal=0.5, value_p_randomgoal=0.3, value_geom_sample=True,
    actor_p_curgoal=0.0, actor_p_trajgoal=1.0, actor_p_randomgoal=0.0, actor_geom_sample=False,
    gc_negative=True, p_aug=0.0, frame_stack=None,
)

WORKLOADS: Dict[str, Workload] = {
    'c1': Workload('c1', 'pointmaze-medium-navigate-v0 shape, GCDataset (GCIVL mix), batch 1024', 1000, 1001, (2,), 'float32', 2,
                   'gc', 1024, dict(_GCIQL), 112, 0),
    'c2': Workload('c2', 'antmaze-large-navigate-v0 shape, GCDataset (GCIQL mix), batch 1024', 1000, 1001, (29,), 'float32', 8,
                   'gc', 1024, dict(_GCIQL), 1024, 1),
    'c3': Workload('c3', 'humanoidmaze-giant-navigate-v0 shape, HGCDataset (HIQL, subgoal_steps=25), batch 1024', 1000, 4001,
                   (69,), 'float32', 21, 'hgc', 1024, dict(_GCIQL, discount=0.995, subgoal_steps=25), 4120, 2),
    'c4': Workload('c4', 'visual-cube-double-play-v0 shape (64x64x3 u8), GCDataset, frame_stack=3, p_aug=0.5, batch 256', 1000, 1001,
                   (64, 64, 3), 'uint8', 5, 'gc', 256, dict(_GCIQL, frame_stack=3, p_aug=0.5), 270408, 3),
    'c5': Workload('c5', 'cube-quadruple-play-100M shape, 1/8 trajectory-aligned shard per GPU, GCDataset, batch 4096', 12500, 1001,
                   (55,), 'float32', 5, 'gc', 4096, dict(_GCIQL), 1832, 4),
    # ---- secondary runs named in SURVEY.md 8(d) (same shapes, the benchmark's other sampler settings) ----
    # HIQL on humanoidmaze-giant with the benchmark's own subgoal_steps=100 (hyperparameters.sh:285)
    'c3b': Workload('c3b', 'humanoidmaze-giant-navigate-v0 shape, HGCDataset (HIQL, subgoal_steps=100), batch 1024', 1000, 4001,
                    (69,), 'float32', 21, 'hgc', 1024, dict(_GCIQL, discount=0.995, subgoal_steps=100), 4120, 2),
    # HIQL on pixels, subgoal_steps=10 (hyperparameters.sh:857): 7 stacked images written, 19 distinct frames read
    'c4b': Workload('c4b', 'visual-cube-double-play-v0 shape (64x64x3 u8), HGCDataset (HIQL, subgoal_steps=10), frame_stack=3, '
                    'p_aug=0.5, batch 256', 1000, 1001, (64, 64, 3), 'uint8', 5, 'hgc', 256,
                    dict(_GCIQL, frame_stack=3, p_aug=0.5, subgoal_steps=10), 7 * 36864 + 9 * 8 + 28 + 19 * 12288 + 28, 3),
    # SHARSA's sampler settings (sharsa.py:410-422) on the 100M-shape shard: 7 distinct row gathers
    'c5b': Workload('c5b', 'cube-quadruple-play-100M shape, 1/8 trajectory-aligned shard per GPU, HGCDataset (SHARSA mix), batch 4096',
                    12500, 1001, (55,), 'float32', 5, 'hgc', 4096,
                    dict(_GCIQL, value_geom_sample=False, actor_p_curgoal=0.0, actor_p_trajgoal=0.5, actor_p_randomgoal=0.5,
                         actor_geom_sample=True, gc_negative=False, discount=0.999, subgoal_steps=25),
                    7 * 220 + 28 + 7 * 220 + 28 + 9 * 8, 4),
}


def compact_flags(episodes: int, steps: int):
    """terminals / valids of a compact dataset with fixed-length trajectories (ogbench/utils.py:71-73)."""
    n = episodes * steps
    terminals = np.zeros(n, dtype=np.float32)
    terminals[steps - 1::steps] = 1.0
    valids = (1.0 - terminals).astype(np.float32)
    shifted = np.concatenate([terminals[1:], [1.0]]).astype(np.float32)
    terminals = np.minimum(terminals + shifted, 1.0).astype(np.float32)
    return terminals, valids


def host_fields(w: Workload, episodes: Optional[int] = None, seed: Optional[int] = None) -> Dict[str, np.ndarray]:
    """numpy fields (optionally fewer episodes, e.g. for the CPU baseline of the pixel workload)."""
    episodes = w.episodes if episodes is None else episodes
    rng = np.random.default_rng(w.seed if seed is None else seed)
    n = episodes * w.steps
    if w.obs_dtype == 'uint8':
        obs = rng.integers(0, 256, size=(n, *w.obs_shape), dtype=np.uint8)
    else:
        obs = rng.standard_normal((n, *w.obs_shape), dtype=np.float32)
    actions = rng.uniform(-1.0, 1.0, size=(n, w.act_dim)).astype(np.float32)
    terminals, valids = compact_flags(episodes, w.steps)
    return dict(observations=obs, actions=actions, terminals=terminals, valids=valids)


def device_fields(w: Workload, device: int = 0, episodes: Optional[int] = None, seed: Optional[int] = None):
    """Same shapes generated directly in HBM with torch (plumbing only: avoids a multi-GB host upload)."""
    import torch

    episodes = w.episodes if episodes is None else episodes
    n = episodes * w.steps
    dev = torch.device('cuda', device)
    gen = torch.Generator(device=dev)
    gen.manual_seed(w.seed if seed is None else seed)
    if w.obs_dtype == 'uint8':
        obs = torch.randint(0, 256, (n, *w.obs_shape), dtype=torch.uint8, device=dev, generator=gen)
    else:
        obs = torch.randn((n, *w.obs_shape), dtype=torch.float32, device=dev, generator=gen)
    actions = torch.rand((n, w.act_dim), dtype=torch.float32, device=dev, generator=gen) * 2 - 1
    terminals, valids = compact_flags(episodes, w.steps)
    return dict(observations=obs, actions=actions, terminals=terminals, valids=valids)
