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


def main():
    main_oracle_reps()
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
