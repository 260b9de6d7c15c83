"""Golden vectors for the loader from the unmodified reference `ogbench/utils.py:load_dataset` (imported under stubs
for gymnasium and ogbench.relabel_utils, neither of which the function touches).  -> tests/golden/loader_*.npz"""

import importlib.util
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = os.environ.get('OGB_REFERENCE_ROOT', '/root/reference')
HERE = os.path.dirname(os.path.abspath(__file__))


def raw_file(path, seed, lengths, pixel=False, info=True):
    """A raw OGBench .npz as data_gen_scripts write it (generate_locomaze.py:196-210): bool terminals, one per episode."""
    rng = np.random.default_rng(seed)
    n = int(np.sum(lengths))
    obs = rng.integers(0, 256, (n, 4, 4, 3), dtype=np.uint8) if pixel else rng.standard_normal((n, 3))
    terminals = np.zeros(n, dtype=bool)
    terminals[np.cumsum(lengths) - 1] = True
    payload = dict(observations=obs, actions=rng.uniform(-1, 1, (n, 2)), terminals=terminals)
    if info:
        payload.update(qpos=rng.standard_normal((n, 2)), qvel=rng.standard_normal((n, 2)), button_states=rng.integers(0, 2, (n, 3)))
    np.savez_compressed(path, **payload)


def load_reference_utils():
    for name in ('gymnasium', 'ogbench', 'ogbench.relabel_utils'):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules['ogbench.relabel_utils'].add_oracle_reps = None
    sys.modules['ogbench.relabel_utils'].relabel_dataset = None
    spec = importlib.util.spec_from_file_location('ogb_reference_utils', os.path.join(REF, 'ogbench', 'utils.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


CASES = [
    ('loader_state', dict(seed=1, lengths=[5, 3, 9, 2], pixel=False), dict(ob_dtype=np.float32, action_dtype=np.float32)),
    ('loader_pixel', dict(seed=2, lengths=[4, 6], pixel=True), dict(ob_dtype=np.uint8, action_dtype=np.float32)),
]


ORACLE_ENVS = [  # (env name, num_cubes, num_buttons)
    ('antmaze-large-navigate-oraclerep-v0', None, None), ('antsoccer-arena-navigate-oraclerep-v0', None, None),
    ('cube-double-play-oraclerep-v0', 2, None), ('scene-play-oraclerep-v0', 1, 2), ('puzzle-3x3-play-oraclerep-v0', None, None),
]


def oracle_inputs(seed=5, n=13):
    rng = np.random.default_rng(seed)
    return dict(qpos=rng.standard_normal((n, 40)), button_states=rng.integers(0, 2, (n, 9)).astype(np.int64))


def load_reference_relabel():
    spec = importlib.util.spec_from_file_location('ogb_reference_relabel', os.path.join(REF, 'ogbench', 'relabel_utils.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main_oracle_reps():
    """ogbench/relabel_utils.py:93-155 run unmodified on synthetic qpos / button_states -> loader_oracle_reps.npz"""
    ref = load_reference_relabel()
    payload = {}
    for name, cubes, buttons in ORACLE_ENVS:
        env = types.SimpleNamespace(unwrapped=types.SimpleNamespace(_num_cubes=cubes, _num_buttons=buttons))
        ds = oracle_inputs()
        ref.add_oracle_reps(name, env, ds)
        payload[name] = ds['oracle_reps']
    np.savez_compressed(os.path.join(HERE, 'loader_oracle_reps.npz'), **payload)
    print('loader_oracle_reps', len(payload), 'arrays')


MAKE_CASES = [  # (dataset name, pixel observations, num_cubes, add_info)
    ('antmaze-large-navigate-v0', False, None, False),
    ('visual-cube-single-play-v0', True, None, False),
    ('powderworld-easy-play-v0', True, None, True),
    ('cube-double-play-oraclerep-v0', False, 2, False),
    ('pointmaze-medium-navigate-oraclerep-v0', False, None, True),
]


def make_raw_pair(directory, file_stem, pixel, seed):
    """Raw train / validation files as the data-generation scripts write them (float64 observations/actions, bool terminals)."""
    rng = np.random.default_rng(seed)
    for suffix, lengths in (('', [6, 4, 7]), ('-val', [5, 3])):
        n = int(np.sum(lengths))
        obs = rng.integers(0, 256, (n, 4, 4, 3), dtype=np.uint8) if pixel else rng.standard_normal((n, 3))
        terminals = np.zeros(n, dtype=bool)
        terminals[np.cumsum(lengths) - 1] = True
        np.savez_compressed(os.path.join(directory, f'{file_stem}{suffix}.npz'), observations=obs,
                            actions=rng.uniform(-1, 1, (n, 2)) * (3 if 'powderworld' in file_stem else 1), terminals=terminals,
                            qpos=rng.standard_normal((n, 30)), qvel=rng.standard_normal((n, 4)), button_states=rng.integers(0, 2, (n, 3)))


def load_reference_utils_with_relabel():
    """ogbench/utils.py with the real ogbench/relabel_utils.py behind it (gymnasium stubbed: dataset_only never calls it)."""
    relabel = load_reference_relabel()
    for name in ('gymnasium', 'ogbench'):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules['ogbench.relabel_utils'] = relabel
    spec = importlib.util.spec_from_file_location('ogb_reference_utils2', os.path.join(REF, 'ogbench', 'utils.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def file_stem(name):
    splits = name.split('-')
    return '-'.join(splits[:-2] + splits[-1:]) if 'oraclerep' in splits else name


def main_make_datasets():
    """ogbench.make_env_and_datasets(..., dataset_only=True) run unmodified -> loader_make_datasets.npz (+ the raw files)."""
    ref = load_reference_utils_with_relabel()
    raw_dir = os.path.join(HERE, 'loader_raw')
    os.makedirs(raw_dir, exist_ok=True)
    payload = {}
    for i, (name, pixel, cubes, add_info) in enumerate(MAKE_CASES):
        make_raw_pair(raw_dir, file_stem(name), pixel, 50 + i)
        env = types.SimpleNamespace(unwrapped=types.SimpleNamespace(_num_cubes=cubes, _num_buttons=None))
        train, val = ref.make_env_and_datasets(name, dataset_path=os.path.join(raw_dir, file_stem(name) + '.npz'), compact_dataset=True,
                                               dataset_only=True, cur_env=env, add_info=add_info)
        for split, ds in (('train', train), ('val', val)):
            for k, v in ds.items():
                payload[f'{name}/{split}/{k}'] = v
    np.savez_compressed(os.path.join(HERE, 'loader_make_datasets.npz'), **payload)
    print('loader_make_datasets', len(payload), 'arrays')


SINGLETASK_ENVS = ['antmaze-large-singletask-task1-v0', 'antsoccer-arena-singletask-v0', 'cube-double-singletask-task2-v0',
                   'scene-singletask-task3-v0', 'puzzle-3x3-singletask-v0']


def singletask_env(name):
    """What relabel_dataset reads from a single-task environment (ogbench/relabel_utils.py:13-84), as a plain namespace."""
    rng = np.random.default_rng(len(name))
    task = types.SimpleNamespace(_reward_task_id=1, cur_goal_xy=rng.standard_normal(2), _goal_tol=0.9, _num_cubes=2, _num_buttons=9,
                                 _data=types.SimpleNamespace(mocap_pos=rng.standard_normal((2, 3)) * 0.05),
                                 _target_button_states=rng.integers(0, 2, 9), _target_drawer_pos=0.01, _target_window_pos=-0.02)
    return types.SimpleNamespace(unwrapped=task, reset=lambda: None)


def singletask_inputs(env, seed=9, n=45):
    """qpos / button_states: a third of the rows satisfies every sub-goal of `env`'s task, a third only some, the rest none."""
    rng = np.random.default_rng(seed)
    task = env.unwrapped
    qpos = rng.standard_normal((n, 40))
    buttons = rng.integers(0, 2, (n, 9)).astype(np.int64)
    third = n // 3
    drawer = 14 + task._num_cubes * 7 + task._num_buttons
    for r in range(2 * third):
        full = r < third
        qpos[r, :2] = qpos[r, 15:17] = task.cur_goal_xy + rng.uniform(-0.3, 0.3, 2)
        for c in range(task._num_cubes if full else 1):
            qpos[r, 14 + 7 * c:17 + 7 * c] = task._data.mocap_pos[c] + rng.uniform(-0.015, 0.015, 3)
        if full:
            buttons[r] = task._target_button_states
            qpos[r, drawer], qpos[r, drawer + 1] = task._target_drawer_pos + 0.02, task._target_window_pos - 0.03
    return dict(qpos=qpos, button_states=buttons)


def main_singletask():
    """ogbench/relabel_utils.py:4-90 run unmodified -> loader_singletask.npz; plus one make_env_and_datasets case."""
    ref = load_reference_relabel()
    payload = {}
    for name in SINGLETASK_ENVS:
        ds = singletask_inputs(singletask_env(name))
        ref.relabel_dataset(name, singletask_env(name), ds)
        payload[f'{name}/rewards'], payload[f'{name}/masks'] = ds['rewards'], ds['masks']
        assert 0 < (ds['masks'] == 0).sum() < len(ds['masks']), name      # both outcomes occur
    utils = load_reference_utils_with_relabel()
    name = 'antmaze-large-navigate-singletask-task1-v0'                    # files: antmaze-large-navigate-v0(.npz, -val.npz)
    train, val = utils.make_env_and_datasets(name, dataset_path=os.path.join(HERE, 'loader_raw', 'antmaze-large-navigate-v0.npz'),
                                             compact_dataset=True, dataset_only=True, cur_env=singletask_env('antmaze-large-singletask-task1-v0'))
    for split, ds in (('train', train), ('val', val)):
        for k, v in ds.items():
            payload[f'make/{split}/{k}'] = v
    np.savez_compressed(os.path.join(HERE, 'loader_singletask.npz'), **payload)
    print('loader_singletask', len(payload), 'arrays')


def main():
    main_oracle_reps()
    main_make_datasets()
    main_singletask()
    ref = load_reference_utils()
    for name, raw_kw, load_kw in CASES:
        raw = os.path.join(HERE, name + '_raw.npz')
        raw_file(raw, **raw_kw)
        payload = {}
        for compact in (False, True):
            for add_info in (False, True):
                out = ref.load_dataset(raw, compact_dataset=compact, add_info=add_info, **load_kw)
                for k, v in out.items():
                    payload[f'c{int(compact)}i{int(add_info)}/{k}'] = v
        np.savez_compressed(os.path.join(HERE, name + '.npz'), **payload)
        print(name, len(payload), 'arrays')


if __name__ == '__main__':
    main()
