"""Bit-exact oracle parity at the BASELINE.json shapes (C1..C5).

The toy-sized sweeps of test_gpu_parity.py / test_gpu_philox.py exercise the algebra; these run it where it is used:
1,000 x 1,001-row segments (C1, C2), 1,000 x 4,001 (C3), a 12,512,500-row shard with 25,000 terminal_locs (C5), and
64x64x3 frames with stacking and crop (C4; 100 episodes = 1.2 GB of frames, the per-sample work of the full 12.3 GB
set).  Each shape is compared with the oracle (oracle/replay_oracle.py, pinned to the unmodified reference by
tests/golden) on the very same host fields, every key `array_equal`, in both draw modes:

  (a) rng='numpy'  -- the global np.random stream under a fixed seed, the reference's own call sequence
                      (datasets.py:213-294, :496-643);
  (b) Philox mode  -- the device's draws rebuilt in numpy (oracle/philox_np.py) and fed to the oracle.

Three training calls plus one evaluation=True call per mode.  C1 is the index-kernel-only path (all rows <= 16 B).
"""

import numpy as np
import pytest

from oracle import philox_np
from oracle.replay_oracle import DrawsSource, OracleSampler
from tests.gpu_util import to_host
from tests.test_gpu_philox import goal_sets_for

pytestmark = pytest.mark.gpu

SHAPES = ['c1', 'c2', 'c3', 'c5', 'c4']


def _fields(key):
    from ogbench_b200 import synthetic

    w = synthetic.WORKLOADS[key]
    episodes = 100 if w.obs_dtype == 'uint8' else w.episodes   # C4: 100 x 1,001 frames of 64x64x3 (1.2 GB)
    return w, synthetic.host_fields(w, episodes=episodes)


def _compare(got, want, rows_ok=None, tag=''):
    assert set(got) == set(want), set(got) ^ set(want)
    for k in want:
        g, x = got[k], want[k]
        assert g.dtype == x.dtype and g.shape == x.shape, (tag, k, g.dtype, x.dtype, g.shape, x.shape)
        if rows_ok is None:
            assert np.array_equal(g, x), (tag, k)
        else:
            assert np.array_equal(g[rows_ok], x[rows_ok]), (tag, k)


@pytest.mark.parametrize('key', SHAPES)
def test_full_size_bit_exact_vs_oracle(key):
    from ogbench_b200 import Dataset, GCDataset, HGCDataset

    w, fields = _fields(key)
    cls = GCDataset if w.kind == 'gc' else HGCDataset
    oracle = OracleSampler(fields, w.config, w.kind)
    ds = Dataset.create(**{k: v for k, v in fields.items()})
    B = w.batch

    # ---- (a) the reference's own np.random sequence ----
    dev = cls(ds, w.config, rng='numpy')
    for call in range(4):
        evaluation = call == 3
        np.random.seed(1000 + call)
        want = oracle.sample(B, evaluation=evaluation)
        state_after = np.random.get_state()[1][:8].copy()
        np.random.seed(1000 + call)
        got = to_host(dev.sample(B, evaluation=evaluation))
        assert np.array_equal(np.random.get_state()[1][:8], state_after)    # the same number of draws was consumed
        _compare(got, want, tag=f'{key} numpy call {call}')
    del dev

    # ---- (b) the on-device Philox draws, replayed through the oracle ----
    seed, stream_id = 0xC0FFEE1234, 3
    dev = cls(ds, w.config, seed=seed, stream_id=stream_id)
    n_choices = len(oracle.valid_table)
    for call in range(4):
        evaluation = call == 3
        got = to_host(dev.sample(B, evaluation=evaluation))
        aug = w.config['p_aug'] is not None and not evaluation
        draws, knife = philox_np.philox_draws(seed, stream_id, call, B, n_choices, goal_sets_for(w.config, w.kind), aug,
                                              w.config['p_aug'] or 0.0)
        src = DrawsSource(draws)
        want = oracle.sample(B, evaluation=evaluation, source=src)
        assert src.exhausted()
        ok = ~knife                      # rows whose geometric quotient sits within 1e-9 of an integer (none in practice)
        assert ok.mean() > 0.999
        _compare(got, want, rows_ok=ok, tag=f'{key} philox call {call}')


@pytest.mark.parametrize('key', ['c1', 'c2', 'c5'])
def test_full_size_given_idxs_and_many(key):
    """Explicit idxs at the extremes of the table (first / last drawable rows, trajectory ends) and a multi-batch
    launch sliced against successive oracle calls on the rebuilt Philox draws."""
    from ogbench_b200 import Dataset, GCDataset

    w, fields = _fields(key)
    oracle = OracleSampler(fields, w.config, w.kind)
    ds = Dataset.create(**fields)
    dev = GCDataset(ds, w.config, rng='numpy')
    T, n = w.steps, w.rows
    idxs = np.array([0, 1, T - 3, T - 2, T, 2 * T - 2, n - T, n - 3, n - 2, n // 2, n // 2 + 1], dtype=np.int64)
    np.random.seed(7)
    want = oracle.sample(len(idxs), idxs=idxs)
    np.random.seed(7)
    got = to_host(dev.sample(len(idxs), idxs=idxs))
    _compare(got, want, tag=f'{key} idxs')

    seed, stream_id, K = 99, 1, 5
    dev = GCDataset(ds, w.config, seed=seed, stream_id=stream_id)
    many = to_host(dev.sample_many(K, w.batch))
    n_choices = len(oracle.valid_table)
    for call in range(K):
        draws, knife = philox_np.philox_draws(seed, stream_id, call, w.batch, n_choices, goal_sets_for(w.config, w.kind), True, 0.0)
        want = oracle.sample(w.batch, source=DrawsSource(draws))
        _compare({k: v[call] for k, v in many.items()}, want, rows_ok=~knife, tag=f'{key} many {call}')


@pytest.mark.parametrize('key,K', [('c2', 160), ('c3', 96), ('c5', 40)])
def test_many_tile_launch_bit_exact_vs_oracle(key, K):
    """A launch with more 32-row tiles than the gather kernel has warps (C2/C5: 2,368 warps, C3: 1,184), so that most tiles
    are handed out by the ticket counter: every batch of the launch against the oracle on the rebuilt Philox draws."""
    from ogbench_b200 import Dataset, GCDataset, HGCDataset

    w, fields = _fields(key)
    cls = GCDataset if w.kind == 'gc' else HGCDataset
    oracle = OracleSampler(fields, w.config, w.kind)
    ds = Dataset.create(**fields)
    seed, stream_id = 424242, 5
    dev = cls(ds, w.config, seed=seed, stream_id=stream_id)
    assert K * w.batch // 32 > 2 * (1184 if key == 'c3' else 2368)
    many = to_host(dev.sample_many(K, w.batch))
    n_choices = len(oracle.valid_table)
    for call in range(K):
        draws, knife = philox_np.philox_draws(seed, stream_id, call, w.batch, n_choices, goal_sets_for(w.config, w.kind), True, 0.0)
        want = oracle.sample(w.batch, source=DrawsSource(draws))
        _compare({k: v[call] for k, v in many.items()}, want, rows_ok=~knife, tag=f'{key} many-tile {call}')
