"""Loader (ogbench/utils.py:14-96) against golden vectors from the reference; shard cycling (impls/main.py:185-199)."""

import os

import numpy as np
import pytest

from ogbench_b200 import loader
from tests.golden.make_golden_loader import CASES

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


@pytest.mark.parametrize('name,raw_kw,load_kw', CASES, ids=[c[0] for c in CASES])
def test_load_dataset_matches_reference(name, raw_kw, load_kw):
    want = np.load(os.path.join(HERE, name + '.npz'))
    raw = os.path.join(HERE, name + '_raw.npz')
    for compact in (False, True):
        for add_info in (False, True):
            got = loader.load_dataset(raw, compact_dataset=compact, add_info=add_info, **load_kw)
            prefix = f'c{int(compact)}i{int(add_info)}/'
            keys = {k[len(prefix):] for k in want.files if k.startswith(prefix)}
            assert set(got) == keys
            for k in keys:
                w = want[prefix + k]
                assert got[k].dtype == w.dtype and got[k].shape == w.shape and np.array_equal(got[k], w), (compact, add_info, k)


def test_add_oracle_reps_matches_reference():
    """add_oracle_reps (ogbench/relabel_utils.py:93-155) against arrays produced by the unmodified reference function."""
    import types

    from tests.golden.make_golden_loader import ORACLE_ENVS, oracle_inputs

    want = np.load(os.path.join(HERE, 'loader_oracle_reps.npz'))
    for name, cubes, buttons in ORACLE_ENVS:
        ds = oracle_inputs()
        loader.add_oracle_reps(name, None, ds, num_cubes=cubes, num_buttons=buttons)
        assert ds['oracle_reps'].dtype == np.float32 and np.array_equal(ds['oracle_reps'], want[name]), name
        env = types.SimpleNamespace(unwrapped=types.SimpleNamespace(_num_cubes=cubes, _num_buttons=buttons))
        ds2 = oracle_inputs()
        loader.add_oracle_reps(name, env, ds2)
        assert np.array_equal(ds2['oracle_reps'], want[name]), name
    with pytest.raises(ValueError):
        loader.add_oracle_reps('powderworld-easy-play-v0', None, oracle_inputs())


def test_make_datasets_matches_reference():
    """make_datasets = the dataset half of ogbench.make_env_and_datasets(dataset_only=True) (ogbench/utils.py:134-235):
    file resolution, dtypes per environment family, oraclerep handling, removal of the info keys."""
    from tests.golden.make_golden_loader import MAKE_CASES, file_stem

    want = np.load(os.path.join(HERE, 'loader_make_datasets.npz'))
    raw_dir = os.path.join(HERE, 'loader_raw')
    for name, _, cubes, add_info in MAKE_CASES:
        got = loader.make_datasets(name, dataset_path=os.path.join(raw_dir, file_stem(name) + '.npz'), compact_dataset=True,
                                   add_info=add_info, num_cubes=cubes)
        for split, ds in zip(('train', 'val'), got):
            prefix = f'{name}/{split}/'
            keys = {k[len(prefix):] for k in want.files if k.startswith(prefix)}
            assert set(ds) == keys, (name, split, set(ds) ^ keys)
            for k in keys:
                w = want[prefix + k]
                assert ds[k].dtype == w.dtype and ds[k].shape == w.shape and np.array_equal(ds[k], w), (name, split, k)
    with pytest.raises(ValueError):                                           # single-task relabelling needs the environment
        loader.make_datasets('cube-single-play-singletask-v0', dataset_path='x.npz')
    with pytest.raises(FileNotFoundError):
        loader.make_datasets('antmaze-large-navigate-v0', dataset_dir=raw_dir + '/missing')


def test_singletask_relabelling_matches_reference():
    """relabel_dataset (ogbench/relabel_utils.py:4-90) against rewards / masks computed by the unmodified reference function
    for the five environment families, and make_datasets for a 'singletask' name (ogbench/utils.py:164-171, 217-220)."""
    from tests.golden.make_golden_loader import SINGLETASK_ENVS, singletask_env, singletask_inputs

    want = np.load(os.path.join(HERE, 'loader_singletask.npz'))
    for name in SINGLETASK_ENVS:
        ds = singletask_inputs(singletask_env(name))
        loader.relabel_dataset(name, singletask_env(name), ds)
        for k in ('rewards', 'masks'):
            w = want[f'{name}/{k}']
            assert ds[k].dtype == w.dtype == np.float32 and np.array_equal(ds[k], w), (name, k)
    with pytest.raises(ValueError):
        loader.relabel_dataset('powderworld-easy-singletask-v0', singletask_env('x'), singletask_inputs(singletask_env('x')))
    env = singletask_env('x')
    env.unwrapped._reward_task_id = None
    with pytest.raises(AssertionError):
        loader.relabel_dataset('antmaze-large-singletask-v0', env, singletask_inputs(env))
    got = loader.make_datasets('antmaze-large-navigate-singletask-task1-v0', dataset_path=os.path.join(HERE, 'loader_raw', 'antmaze-large-navigate-v0.npz'),
                               compact_dataset=True, env=singletask_env('antmaze-large-singletask-task1-v0'))
    for split, ds in zip(('train', 'val'), got):
        keys = {k[len(f'make/{split}/'):] for k in want.files if k.startswith(f'make/{split}/')}
        assert set(ds) == keys and 'rewards' in keys and 'qpos' not in keys, (split, set(ds) ^ keys)
        for k in keys:
            w = want[f'make/{split}/{k}']
            assert ds[k].dtype == w.dtype and np.array_equal(ds[k], w), (split, k)


def test_list_shards(tmp_path):
    for n in ('b.npz', 'a.npz', 'a-val.npz', 'c.txt'):
        (tmp_path / n).write_bytes(b'')
    assert [os.path.basename(p) for p in loader.list_shards(str(tmp_path))] == ['a.npz', 'b.npz']
    with pytest.raises(FileNotFoundError):
        loader.list_shards(str(tmp_path / 'missing'))


def test_shard_cycler_schedule_without_device():
    """Swap schedule of impls/main.py:185-199: at every step i with i % interval == 0 the next shard becomes current."""
    loads = []

    def make(path):
        loads.append(path)
        return path

    cyc = loader.ShardCycler(['s0', 's1', 's2'], make, replace_interval=4)
    seen = [cyc.at_step(i) for i in range(1, 14)]
    cyc.close()
    assert seen == ['s0'] * 3 + ['s1'] * 4 + ['s2'] * 4 + ['s0'] * 2
    assert cyc.swaps == 3 and loads[:4] == ['s0', 's1', 's2', 's0']
    single = loader.ShardCycler(['only'], make, replace_interval=2)
    assert [single.at_step(i) for i in range(1, 6)] == ['only'] * 5


@pytest.mark.gpu
def test_shard_cycler_on_device(tmp_path):
    from tests.golden.make_golden import cfg

    paths = []
    for s in range(3):
        rng = np.random.default_rng(s)
        n_ep, T = 6, 30
        obs = rng.standard_normal((n_ep * T, 4)).astype(np.float32)
        obs[:, 0] = 1000 * s + np.arange(n_ep * T)                     # shard id readable from the batch
        term = np.zeros(n_ep * T, dtype=bool)
        term[T - 1::T] = True
        p = str(tmp_path / f'shard{s}.npz')
        np.savez(p, observations=obs, actions=rng.uniform(-1, 1, (n_ep * T, 2)).astype(np.float32), terminals=term)
        paths.append(p)
    config = cfg()
    cyc = loader.ShardCycler(loader.list_shards(str(tmp_path)), lambda p: loader.load_gc_dataset(p, config, seed=1), replace_interval=5)
    for i in range(1, 16):
        batch = cyc.at_step(i).sample(64)
        shard = (i // 5) % 3
        for key in ('observations', 'next_observations', 'value_goals', 'actor_goals'):   # goals are shard-local
            ids = np.asarray(batch[key])[:, 0]
            assert ((ids >= 1000 * shard) & (ids < 1000 * shard + 180)).all(), (i, key)
    cyc.close()
    assert cyc.swaps == 3


@pytest.mark.gpu
def test_shard_cycler_revisits_draw_new_batches(tmp_path):
    """The reference keeps one np.random stream across shard swaps, so a revisited shard never replays its batches: the
    cycler carries the Philox counter from sampler to sampler (and no two steps of a run share a counter)."""
    from tests.golden.make_golden import cfg

    for s in range(2):
        rng = np.random.default_rng(10 + s)
        n_ep, T = 5, 40
        term = np.zeros(n_ep * T, dtype=bool)
        term[T - 1::T] = True
        np.savez(str(tmp_path / f'shard{s}.npz'), observations=rng.standard_normal((n_ep * T, 3)).astype(np.float32),
                 actions=rng.uniform(-1, 1, (n_ep * T, 2)).astype(np.float32), terminals=term)
    config = cfg()
    cyc = loader.ShardCycler(loader.list_shards(str(tmp_path)), lambda p: loader.load_gc_dataset(p, config, seed=3), replace_interval=4)
    visits, counters = {}, []
    for i in range(1, 17):                       # shards 0,1,0,1 in blocks of four steps (the first block has three)
        sampler = cyc.at_step(i)
        counters.append(sampler.state_dict()['counter'])
        batch = sampler.sample(32)
        visits.setdefault(cyc.index, []).append(np.asarray(batch['observations']).copy())
    cyc.close()
    assert counters == list(range(16))           # one global batch counter for the whole run
    first, second = visits[0][:3], visits[0][3:6]
    assert not any(np.array_equal(a, b) for a in first for b in second)


def test_shard_cycler_carries_sampler_state():
    class Fake:
        def __init__(self, path):
            self.path, self.counter = path, 0

        def state_dict(self):
            return {'counter': self.counter}

        def load_state_dict(self, state):
            self.counter = state['counter']

    cyc = loader.ShardCycler(['a', 'b'], Fake, replace_interval=3, prefetch=False)
    seen = []
    for i in range(1, 10):
        s = cyc.at_step(i)
        seen.append((s.path, s.counter))
        s.counter += 1
    assert seen == [('a', 0), ('a', 1), ('b', 2), ('b', 3), ('b', 4), ('a', 5), ('a', 6), ('a', 7), ('b', 8)]
