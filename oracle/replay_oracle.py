"""TEST INFRASTRUCTURE ONLY -- CPU (numpy) restatement of the reference replay sampler.

This file is the *checker* for the CUDA path.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the product package
(``ogbench_b200``) never does and has no CPU fallback.

What it restates: ``GCDataset.sample`` / ``HGCDataset.sample`` and the helpers they call in the reference file
``impls/utils/datasets.py`` (every function below cites the lines it follows).  The algorithm is written here
as free functions over explicit index vectors plus an explicit *draw source*, instead of the reference's classes
over the global ``np.random`` stream, so that

* the exact random draws one ``sample()`` consumed come back as a structured ``Draws`` record -- the input of
  the CUDA sampler's validation mode -- and
* the per-row index algebra is visible and separately testable.

Parity pin: the reference ships no tests and no golden vectors ("parity unpinned" by the reference itself).
The pin is ours: ``tests/golden/make_golden.py`` runs the UNMODIFIED reference file (imported under the stubs in
``oracle/refshim.py``) on seeded toy datasets, records its draws and outputs into ``tests/golden/*.npz``, and
``tests/test_oracle_golden.py`` requires this restatement to reproduce every key bit-for-bit from the recorded
draws.  When ``/root/reference`` is mounted the same comparison also runs live.

Third-party arithmetic on the path (numpy, un-pinned by the reference; 2.3.5 here): legacy ``RandomState``
draws, ``searchsorted``, ``round`` (half-even), ``**``.  They are taken from the same numpy the tests run with.
"""

from __future__ import annotations

import dataclasses
from typing import Any, Dict, List, Optional

import numpy as np

TRL_AGENTS = ('trl', 'latent_trl', 'discrete_latent_trl')


# --------------------------------------------------------------------------------------------------------------
# Draw sources
# --------------------------------------------------------------------------------------------------------------
@dataclasses.dataclass
class GoalDraws:
    """Draws of one ``sample_goals`` call (datasets.py:296-327), already transformed."""

    rand_pos: np.ndarray  # int64[B]   randint(n_valid)  -> position in the valid-row table (row id if no 'valids')
    offset: Optional[np.ndarray] = None  # int64[B]   geometric(1-discount)            (geom_sample)
    dist: Optional[np.ndarray] = None  # float64[B] rand()                            (uniform-in-remainder)
    u_traj: Optional[np.ndarray] = None  # float64[B] rand(); absent when p_curgoal == 1.0
    u_cur: Optional[np.ndarray] = None  # float64[B] rand(); absent when p_curgoal == 1.0


@dataclasses.dataclass
class Draws:
    """Everything random one sample() call consumed, in reference call order."""

    idx_pos: Optional[np.ndarray] = None  # int64[B]; None when the caller passed idxs
    goals: List[GoalDraws] = dataclasses.field(default_factory=list)  # value, [low_value], actor
    trl_midpoints: Optional[np.ndarray] = None  # int64[B] randint(idxs, value_goal_idxs), TRL agents only (:259)
    aug_coin: Optional[float] = None  # scalar rand(); None when p_aug is None or evaluation
    crop: Optional[np.ndarray] = None  # int64[B,2] in [0, 2*padding]; None unless aug_coin < p_aug


class NumpyGlobalSource:
    """Draws from the global legacy np.random stream with the reference's own calls (datasets.py:68,70,309,313,321,325,279,333)."""

    def randint(self, high, size):
        return np.random.randint(high, size=size)

    def randint_box(self, low, high, shape):
        return np.random.randint(low, high, shape)

    def randint_between(self, low, high):
        return np.random.randint(low, high)

    def geometric(self, p, size):
        return np.random.geometric(p=p, size=size)

    def rand(self, size):
        return np.random.rand(size)

    def rand_scalar(self):
        return np.random.rand()


class ReplaySource:
    """Replays a DrawRecorder log; asserts call kinds line up, so a schema mistake fails loudly."""

    def __init__(self, log):
        self.log = list(log)
        self.pos = 0

    def _pop(self, kind):
        assert self.pos < len(self.log), 'oracle asked for more draws than the reference consumed'
        got_kind, value = self.log[self.pos]
        assert got_kind == kind, f'draw #{self.pos}: reference called {got_kind}, oracle expected {kind}'
        self.pos += 1
        return value

    def randint(self, high, size):
        out = self._pop('randint')
        assert out.shape == (size,)
        return out

    def randint_box(self, low, high, shape):
        out = self._pop('randint')
        assert tuple(out.shape) == tuple(shape)
        return out

    def randint_between(self, low, high):
        out = self._pop('randint')
        assert out.shape == np.shape(low)
        return out

    def geometric(self, p, size):
        return self._pop('geometric')

    def rand(self, size):
        out = self._pop('rand')
        assert out.shape == (size,)
        return out

    def rand_scalar(self):
        out = self._pop('rand')
        assert out.shape == ()
        return float(out)

    def exhausted(self):
        return self.pos == len(self.log)


class DrawsSource:
    """Serves a structured ``Draws`` record back in reference call order (used to run the oracle on the draws the
    device's Philox mode made, see oracle/philox_np.py)."""

    def __init__(self, draws: Draws):
        q = []
        if draws.idx_pos is not None:
            q.append(draws.idx_pos)
        for g in draws.goals:
            q.append(g.rand_pos)
            q.append(g.offset if g.offset is not None else g.dist)
            if g.u_traj is not None:
                q.extend([g.u_traj, g.u_cur])
        if draws.trl_midpoints is not None:
            q.append(draws.trl_midpoints)
        if draws.aug_coin is not None:
            q.append(draws.aug_coin)
        if draws.crop is not None:
            q.append(draws.crop)
        self.queue = q
        self.pos = 0

    def _pop(self):
        out = self.queue[self.pos]
        self.pos += 1
        return out

    def randint(self, high, size):
        return self._pop()

    def randint_box(self, low, high, shape):
        return self._pop()

    def randint_between(self, low, high):
        return self._pop()

    def geometric(self, p, size):
        return self._pop()

    def rand(self, size):
        return self._pop()

    def rand_scalar(self):
        return float(self._pop())

    def exhausted(self):
        return self.pos == len(self.queue)


# --------------------------------------------------------------------------------------------------------------
# Building blocks
# --------------------------------------------------------------------------------------------------------------
def dataset_size(fields: Dict[str, np.ndarray]) -> int:
    """datasets.py:11-14 -- the longest leaf."""
    return max(len(v) for v in fields.values())


def valid_row_table(fields) -> Optional[np.ndarray]:
    """datasets.py:62-63 -- rows with valids > 0 (int64), or None when the dataset has no 'valids'."""
    if 'valids' not in fields:
        return None
    return np.nonzero(fields['valids'] > 0)[0]


def trajectory_bounds(terminals: np.ndarray):
    """datasets.py:186-187 -- (terminal_locs, initial_locs)."""
    terminal_locs = np.nonzero(terminals > 0)[0]
    initial_locs = np.concatenate([[0], terminal_locs[:-1] + 1])
    return terminal_locs, initial_locs


def rows_from_positions(valid_table, positions):
    """datasets.py:65-70 -- map a randint draw to a dataset row."""
    return positions if valid_table is None else valid_table[positions]


def final_rows(terminal_locs, idxs):
    """datasets.py:306,505 -- first terminal row at or after each idx (searchsorted side='left')."""
    return terminal_locs[np.searchsorted(terminal_locs, idxs)]


def first_rows(initial_locs, idxs):
    """datasets.py:361 -- last initial row at or before each idx (side='right' minus one)."""
    return initial_locs[np.searchsorted(initial_locs, idxs, side='right') - 1]


def stacked_frames(observations, initial_locs, idxs, frame_stack):
    """datasets.py:359-366 -- frames idx-(fs-1)..idx clamped to the trajectory start, oldest first on the last axis."""
    start = first_rows(initial_locs, idxs)
    frames = [observations[np.maximum(idxs - back, start)] for back in range(frame_stack - 1, -1, -1)]
    return np.concatenate(frames, axis=-1)


def shifted_edge_crop(imgs, crop_yx, padding):
    """datasets.py:17-33 in closed form.

    pad(img, p, mode='edge') then dynamic_slice at (cy, cx, 0) with the original shape is
    out[b, y, x, c] = img[b, clip(y + cy_b - p, 0, H-1), clip(x + cx_b - p, 0, W-1), c]; with cy, cx in
    [0, 2p] the slice start never needs XLA's own clamping.
    """
    imgs = np.asarray(imgs)
    b, h, w, _ = imgs.shape
    crop_yx = np.asarray(crop_yx, dtype=np.int64)
    ys = np.clip(np.arange(h)[None, :] + crop_yx[:, 0:1] - padding, 0, h - 1)  # [B,H]
    xs = np.clip(np.arange(w)[None, :] + crop_yx[:, 1:2] - padding, 0, w - 1)  # [B,W]
    return imgs[np.arange(b)[:, None, None], ys[:, :, None], xs[:, None, :]]


def pick_goals(idxs, final, valid_table, p_cur, p_traj, geom_sample, discount, source, size_if_no_valids):
    """datasets.py:296-327 -- returns (goal_idxs, GoalDraws).  Draw order: randint, geometric|rand, rand, rand."""
    n = len(idxs)
    n_choices = size_if_no_valids if valid_table is None else len(valid_table)
    draws = GoalDraws(rand_pos=np.asarray(source.randint(n_choices, n), dtype=np.int64))
    random_goal = rows_from_positions(valid_table, draws.rand_pos)
    if geom_sample:
        draws.offset = np.asarray(source.geometric(1 - discount, n), dtype=np.int64)
        traj_goal = np.minimum(idxs + draws.offset, final)
    else:
        draws.dist = np.asarray(source.rand(n), dtype=np.float64)
        lo = np.minimum(idxs + 1, final)
        # float64, separate multiply and add, round-half-even (:314-316)
        traj_goal = np.round(lo * draws.dist + final * (1 - draws.dist)).astype(int)
    if p_cur == 1.0:
        return idxs, draws
    draws.u_traj = np.asarray(source.rand(n), dtype=np.float64)
    goal = np.where(draws.u_traj < p_traj / (1.0 - p_cur), traj_goal, random_goal)
    draws.u_cur = np.asarray(source.rand(n), dtype=np.float64)
    goal = np.where(draws.u_cur < p_cur, idxs, goal)
    return goal, draws


def subgoal_step(idxs, final, goal, k):
    """datasets.py:478-491 -- (idxs + s, s) with s = min(k, final-idxs), shortened to the goal when it is nearer."""
    steps = np.minimum(np.full(len(idxs), k), final - idxs)
    diff = goal - idxs
    steps = np.where((0 <= diff) & (diff < steps), diff, steps)
    return idxs + steps, steps


def discounted_step_rewards(discount, steps, k, gc_negative):
    """datasets.py:533-541 / 552-560 -- (masks, rewards), float64."""
    success = (steps < k).astype(float)
    masks = 1.0 - success
    if gc_negative:
        rewards = -(1 - discount**steps) / (1 - discount)
    else:
        rewards = (discount**steps) * success
    return masks, rewards


# --------------------------------------------------------------------------------------------------------------
# The sampler
# --------------------------------------------------------------------------------------------------------------
class OracleSampler:
    """numpy restatement of GCDataset / HGCDataset (kind = 'gc' | 'hgc').

    ``fields`` is the dict handed to ``Dataset.create``; ``config`` supports ``[]`` and ``.get``.
    Frames are always stacked on the fly (the reference's ``preprocess_frame_stack=True`` only caches the same
    values, datasets.py:209-211).
    """

    def __init__(self, fields: Dict[str, np.ndarray], config: Any, kind: str = 'gc'):
        assert 'observations' in fields  # datasets.py:54
        assert kind in ('gc', 'hgc')
        self.fields = fields
        self.config = config
        self.kind = kind
        self.size = dataset_size(fields)
        self.valid_table = valid_row_table(fields)
        self.terminal_locs, self.initial_locs = trajectory_bounds(fields['terminals'])
        assert self.terminal_locs[-1] == self.size - 1  # :188
        assert np.isclose(config['value_p_curgoal'] + config['value_p_trajgoal'] + config['value_p_randomgoal'], 1.0)
        assert np.isclose(config['actor_p_curgoal'] + config['actor_p_trajgoal'] + config['actor_p_randomgoal'], 1.0)
        self.trl = kind == 'gc' and config.get('agent_name') in TRL_AGENTS
        if self.trl:
            # datasets.py:198-204: valid_idxs becomes every non-terminal row.  (With frame_stack AND
            # preprocess_frame_stack=True the reference rebuilds its Dataset at :211 and loses this override again; it
            # then trips its own assert at :256 as soon as a final state is drawn, so that mode is not restated.)
            mask = np.ones(self.size, dtype=bool)
            mask[self.terminal_locs] = False
            self.valid_table = np.nonzero(mask)[0]
        if config['frame_stack'] is not None:
            assert 'next_observations' not in fields  # :208
        self.last_draws: Optional[Draws] = None
        self.last_index_vectors: Dict[str, np.ndarray] = {}

    # -- gathers ----------------------------------------------------------------------------------------------
    def _obs(self, idxs):
        """datasets.py:341-346"""
        fs = self.config['frame_stack']
        if fs is None:
            return self.fields['observations'][idxs]
        return stacked_frames(self.fields['observations'], self.initial_locs, idxs, fs)

    def _goal(self, idxs):
        """datasets.py:348-357"""
        if 'oracle_reps' in self.fields:
            return self.fields['oracle_reps'][idxs]
        return self._obs(idxs)

    def _base(self, idxs):
        """datasets.py:72-83 (+229-231): every field at idxs, next_observations synthesised when absent."""
        batch = {k: self.fields[k][idxs] for k in sorted(self.fields)}
        if 'next_observations' not in batch:
            batch['next_observations'] = self.fields['observations'][np.minimum(idxs + 1, self.size - 1)]
        if self.config['frame_stack'] is not None:
            batch['observations'] = self._obs(idxs)
            batch['next_observations'] = self._obs(idxs + 1)  # un-clamped, quirk 3
        return batch

    def _goals(self, idxs, final, prefix, source, geom=None, discount=None):
        cfg = self.config
        goal, draws = pick_goals(
            idxs,
            final,
            self.valid_table,
            cfg[prefix + '_p_curgoal'],
            cfg[prefix + '_p_trajgoal'],
            cfg[prefix + '_geom_sample'] if geom is None else geom,
            cfg['discount'] if discount is None else discount,
            source,
            self.size,
        )
        self.last_draws.goals.append(draws)
        return goal

    def _maybe_crop(self, batch, keys, evaluation, source, padding=3):
        """datasets.py:278-292 / 621-641 + 329-339: one coin per batch, one (cy,cx) per sample for all keys."""
        cfg = self.config
        if cfg['p_aug'] is None or evaluation:
            return
        coin = source.rand_scalar()
        self.last_draws.aug_coin = coin
        if not coin < cfg['p_aug']:
            return
        n = len(batch[keys[0]])
        crop = np.asarray(source.randint_box(0, 2 * padding + 1, (n, 2)), dtype=np.int64)
        self.last_draws.crop = crop
        done = {}
        for key in keys:
            arr = batch[key]
            if arr.ndim != 4:
                continue
            if id(arr) not in done:  # aliases share one cropped array, values identical to cropping twice
                done[id(arr)] = shifted_edge_crop(arr, crop, padding)
            batch[key] = done[id(arr)]

    # -- entry point -------------------------------------------------------------------------------------------
    def sample(self, batch_size, idxs=None, evaluation=False, source=None):
        source = NumpyGlobalSource() if source is None else source
        self.last_draws = Draws()
        if idxs is None:
            n_choices = self.size if self.valid_table is None else len(self.valid_table)
            pos = np.asarray(source.randint(n_choices, batch_size), dtype=np.int64)
            self.last_draws.idx_pos = pos
            idxs = rows_from_positions(self.valid_table, pos)
        idxs = np.asarray(idxs)
        if self.kind == 'gc':
            return self._sample_gc(idxs, evaluation, source)
        return self._sample_hgc(idxs, evaluation, source)

    def _sample_gc(self, idxs, evaluation, source):
        """datasets.py:213-294 (non-TRL)."""
        cfg = self.config
        batch = self._base(idxs)
        final = final_rows(self.terminal_locs, idxs)
        value_goal = self._goals(idxs, final, 'value', source)
        actor_goal = self._goals(idxs, final, 'actor', source)
        batch['value_goals'] = self._goal(value_goal)
        batch['actor_goals'] = self._goal(actor_goal)
        success = (idxs == value_goal).astype(float)
        batch['masks'] = 1.0 - success
        batch['rewards'] = success - (1.0 if cfg['gc_negative'] else 0.0)
        self.last_index_vectors = dict(idxs=idxs, final=final, value_goal=value_goal, actor_goal=actor_goal)
        aug_keys = ['observations', 'next_observations', 'value_goals', 'actor_goals']
        if self.trl:  # datasets.py:254-276
            assert (idxs != final).all()
            assert (idxs != value_goal).all()
            mid = np.asarray(source.randint_between(idxs, value_goal), dtype=np.int64)
            self.last_draws.trl_midpoints = mid
            batch['value_goal_observations'] = self._obs(value_goal)
            batch['actor_goal_observations'] = self._obs(value_goal)   # value_goal_idxs again, as in the reference (:262)
            batch['value_offsets'] = value_goal - idxs
            batch['value_midpoint_offsets'] = mid - idxs
            batch['value_midpoint_observations'] = self._obs(mid)
            batch['value_midpoint_actions'] = self.fields['actions'][mid]
            batch['next_actions'] = self.fields['actions'][idxs + 1]
            batch['value_midpoint_goals'] = self._goal(mid)
            batch['value_cur_goals'] = self._goal(idxs)
            batch['value_next_goals'] = self._goal(idxs + 1)
            self.last_index_vectors['mid'] = mid
            aug_keys += ['value_goal_observations', 'actor_goal_observations', 'value_midpoint_observations',
                         'value_midpoint_goals', 'value_cur_goals', 'value_next_goals']
        self._maybe_crop(batch, aug_keys, evaluation, source)
        return batch

    def _sample_hgc(self, idxs, evaluation, source):
        """datasets.py:496-643."""
        cfg = self.config
        gamma = cfg['discount']
        neg = cfg['gc_negative']
        batch = self._base(idxs)
        final = final_rows(self.terminal_locs, idxs)

        hv_goal = self._goals(idxs, final, 'value', source)
        k_hi = cfg.get('high_subgoal_steps', cfg['subgoal_steps'])
        k_val = k_hi if cfg.get('value_subgoal_steps') is None else cfg['value_subgoal_steps']
        hv_next, hv_s = subgoal_step(idxs, final, hv_goal, k_val)

        batch['high_value_reps'] = batch['observations']  # alias of the un-augmented array (quirk 4)
        batch['high_value_goals'] = self._goal(hv_goal)
        batch['high_value_actions'] = self._goal(hv_next)
        batch['high_value_next_observations'] = self._obs(hv_next)
        batch['high_value_offsets'] = hv_goal - idxs
        batch['high_value_subgoal_steps'] = hv_s
        batch['high_value_masks'], batch['high_value_rewards'] = discounted_step_rewards(gamma, hv_s, k_val, neg)

        k_lo = cfg.get('low_subgoal_steps', cfg['subgoal_steps'])
        lv_next, lv_s = subgoal_step(idxs, final, hv_goal, k_lo)
        batch['low_value_next_observations'] = self._obs(lv_next)
        batch['low_value_subgoal_steps'] = lv_s
        batch['low_value_masks'], batch['low_value_rewards'] = discounted_step_rewards(gamma, lv_s, k_lo, neg)

        lv_goal = None
        if cfg.get('low_discount') is not None:
            lv_goal = self._goals(idxs, final, 'value', source, geom=True, discount=cfg['low_discount'])
            batch['low_value_goals'] = self._goal(lv_goal)
            s = (idxs == lv_goal).astype(float)
            batch['low_value_masks'] = 1.0 - s
            batch['low_value_rewards'] = s - (1.0 if neg else 0.0)

        s = (idxs == hv_goal).astype(float)
        batch['value_goals'] = batch['high_value_goals']
        batch['masks'] = 1.0 - s
        batch['rewards'] = s - (1.0 if neg else 0.0)

        ha_goal = self._goals(idxs, final, 'actor', source)
        k_act = k_hi if cfg.get('actor_subgoal_steps') is None else cfg['actor_subgoal_steps']
        ha_next, _ = subgoal_step(idxs, final, ha_goal, k_act)
        batch['high_actor_goals'] = self._goal(ha_goal)
        batch['high_actor_actions'] = self._goal(ha_next)
        batch['high_actor_next_observations'] = self._obs(ha_next)
        batch['high_actor_targets'] = batch['high_actor_actions']

        la_goal = np.minimum(idxs + k_act, final)
        batch['low_actor_goals'] = self._goal(la_goal)
        batch['low_actor_goal_observations'] = self._obs(la_goal)
        la_next, _ = subgoal_step(idxs, final, ha_goal, k_lo)
        batch['low_actor_next_observations'] = self._obs(la_next)

        self.last_index_vectors = dict(
            idxs=idxs, final=final, hv_goal=hv_goal, hv_next=hv_next, lv_next=lv_next, ha_goal=ha_goal,
            ha_next=ha_next, la_goal=la_goal, la_next=la_next,
        )
        if lv_goal is not None:
            self.last_index_vectors['lv_goal'] = lv_goal
        self._maybe_crop(
            batch,
            [
                'observations', 'next_observations', 'value_goals', 'high_value_goals', 'high_value_actions',
                'high_value_next_observations', 'low_value_next_observations', 'low_actor_goals',
                'low_actor_goal_observations', 'low_actor_next_observations', 'high_actor_goals',
                'high_actor_actions', 'high_actor_next_observations', 'high_actor_targets',
            ],
            evaluation,
            source,
        )
        return batch


class OracleATCSampler:
    """numpy restatement of ATCDataset (datasets.py:369-464): anchor / positive observation pairs (o_t, o_{t+k})."""

    def __init__(self, fields: Dict[str, np.ndarray], config: Any):
        self.fields = fields
        self.config = config
        self.size = dataset_size(fields)
        self.valid_table = valid_row_table(fields)
        self.terminal_locs, self.initial_locs = trajectory_bounds(fields['terminals'])
        assert self.terminal_locs[-1] == self.size - 1  # :391
        if config['frame_stack'] is not None:
            assert 'next_observations' not in fields  # :396
        self._cache = {}
        self.last_draws: Optional[Draws] = None

    def valid_anchors(self, k):
        """datasets.py:417-436"""
        if k not in self._cache:
            cand = self.valid_table if self.valid_table is not None else np.arange(self.size)
            cand = cand[cand + k < self.size]
            cand = cand[cand + k <= final_rows(self.terminal_locs, cand)]
            if len(cand) == 0:
                raise ValueError(f'No valid ATC indices found for k={k}.')
            self._cache[k] = cand
        return self._cache[k]

    def _obs(self, idxs):
        fs = self.config['frame_stack']
        if fs is None:
            return self.fields['observations'][idxs]
        return stacked_frames(self.fields['observations'], self.initial_locs, idxs, fs)

    def sample(self, batch_size, k, evaluation=False, source=None):
        """datasets.py:401-415; np.random.choice(valid, size=B) consumes randint(0, len(valid), B)."""
        source = NumpyGlobalSource() if source is None else source
        self.last_draws = Draws()
        anchors = self.valid_anchors(k)
        pos = np.asarray(source.randint(len(anchors), batch_size), dtype=np.int64)
        self.last_draws.idx_pos = pos
        idxs = anchors[pos]
        batch = {'observations': self._obs(idxs), 'positive_observations': self._obs(idxs + k)}
        if self.config['p_aug'] is not None and not evaluation:
            coin = source.rand_scalar()
            self.last_draws.aug_coin = coin
            if coin < self.config['p_aug']:
                padding = self.config.get('augment_padding', 4)  # :440
                crop = np.asarray(source.randint_box(0, 2 * padding + 1, (batch_size, 2)), dtype=np.int64)
                self.last_draws.crop = crop
                for key in ('observations', 'positive_observations'):
                    if batch[key].ndim == 4:
                        batch[key] = shifted_edge_crop(batch[key], crop, padding)
        return batch


class OracleReplayBuffer:
    """numpy restatement of ReplayBuffer (datasets.py:86-146): a ring of rows plus Dataset.sample (:65-83)."""

    def __init__(self, transition, size):
        self.buffers = {k: np.zeros((size, *np.array(v).shape), dtype=np.array(v).dtype) for k, v in transition.items()}
        self.max_size = size
        self.size = 0
        self.pointer = 0

    @classmethod
    def from_initial_dataset(cls, init, size):
        rb = cls({k: np.asarray(v)[0] for k, v in init.items()}, size)
        n = dataset_size(init)
        for k, v in init.items():
            rb.buffers[k][:n] = v
        rb.size = rb.pointer = n  # :124
        return rb

    def add_transition(self, transition):
        for k, v in transition.items():
            self.buffers[k][self.pointer] = v
        self.pointer = (self.pointer + 1) % self.max_size  # :141
        self.size = max(self.pointer, self.size)  # :142

    def clear(self):
        self.size = self.pointer = 0

    def sample(self, batch_size, idxs=None):
        if idxs is None:
            idxs = np.random.randint(self.size, size=batch_size)  # :70 (a replay buffer has no 'valids')
        batch = {k: v[idxs] for k, v in self.buffers.items()}
        if 'next_observations' not in batch:
            batch['next_observations'] = self.buffers['observations'][np.minimum(idxs + 1, self.size - 1)]  # :82
        return batch
