"""Pytree (nested dict) fields.  The reference maps its gathers over arbitrary pytrees (datasets.py:13,54-56,80,344,365-366);
the device sampler keeps one resident field per leaf and returns every derived key nested the same way.

Pinned by golden vectors recorded from the unmodified reference (tests/golden/make_golden_pytree.py) for the cases the
reference can run -- regular datasets with explicit next_observations (its compact path indexes the observations dict
directly, datasets.py:82, and raises).  Compact datasets with pytree observations work here as the natural extension
(next_observations mapped over the leaves) and are checked against flat samplers on the same draws."""

import json
import os

import numpy as np
import pytest

from tests.golden.make_golden import cfg, ragged, toy_fields
from tests.golden.make_golden_pytree import flatten, nest
from tests.golden_util import GOLDEN_DIR
from tests.gpu_util import device_sampler

PYTREE_CASES = ['pytree_gc_regular', 'pytree_hgc_regular_deep', 'pytree_gc_regular_eval']


def load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + '.npz'))
    meta = json.loads(str(z['meta']))
    fields = nest({k[len('field/'):]: z[k] for k in z.files if k.startswith('field/')})
    out = {k[len('out/'):]: z[k] for k in z.files if k.startswith('out/')}
    return meta, fields, out


def host(batch):
    return {k: np.asarray(v) for k, v in flatten(batch).items()}


def test_pytree_fixtures_are_well_formed():
    for name in PYTREE_CASES:
        meta, fields, out = load(name)
        assert isinstance(fields['observations'], dict) and isinstance(fields['next_observations'], dict)
        leaves = set(flatten(fields['observations']))
        assert {k.split('/', 1)[1] for k in out if k.startswith('value_goals/')} == leaves
        assert all(out[f'observations/{leaf}'].shape[0] == meta['B'] for leaf in leaves)


@pytest.mark.gpu
@pytest.mark.parametrize('name', PYTREE_CASES)
@pytest.mark.parametrize('output', ['device', 'numpy'])
def test_pytree_batches_match_the_reference(name, output):
    meta, fields, want = load(name)
    sampler = device_sampler(fields, meta['cfg'], meta['kind'], rng='numpy', output=output)
    np.random.seed(meta['seed'])
    batch = sampler.sample(meta['B'], evaluation=meta['evaluation'])
    assert isinstance(batch['observations'], dict) and isinstance(batch['value_goals'], dict)     # nested like the reference's
    got = host(batch)
    assert set(got) == set(want), set(got) ^ set(want)
    for k in want:
        assert got[k].dtype == want[k].dtype and got[k].shape == want[k].shape and np.array_equal(got[k], want[k]), (name, k)


@pytest.mark.gpu
@pytest.mark.parametrize('kind,over', [('gc', {}), ('hgc', dict(subgoal_steps=5)), ('gc', dict(frame_stack=3, p_aug=None))])
def test_compact_pytree_observations_equal_flat_samplers_on_the_same_draws(kind, over):
    base = toy_fields(51, ragged(51, 40, 4, 70), (6,), 3, np.float32)
    n = len(base['terminals'])
    rng = np.random.default_rng(7)
    feat = rng.standard_normal((n, 2, 3)).astype(np.float32)
    ids = rng.integers(0, 100, size=(n, 1)).astype(np.int32)
    nested = dict(base, observations={'state': base['observations'], 'aux': {'feat': feat, 'ids': ids}})
    config = cfg(**over)
    tree = device_sampler(nested, config, kind, seed=31, lookahead=3)
    flats = {leaf: device_sampler(dict(base, observations=arr), config, kind, seed=31)
             for leaf, arr in (('state', base['observations']), ('aux/feat', feat), ('aux/ids', ids))}
    for step in range(5):                                   # crosses a look-ahead block boundary
        got = host(tree.sample(96))
        for leaf, flat in flats.items():
            ref = host(flat.sample(96))
            for k, v in ref.items():
                key = f'{k}/{leaf}' if f'{k}/{leaf}' in got else k          # obs-shaped keys carry the leaf path
                assert np.array_equal(got[key], v), (step, leaf, k)
        obs_keys = {k.split('/')[0] for k in got if k.endswith('/state')}
        assert 'observations' in obs_keys and 'next_observations' in obs_keys and ('value_goals' in obs_keys)
    assert isinstance(tree.get_observations(np.arange(4)), dict)


@pytest.mark.gpu
def test_pytree_replay_buffer_and_refusals():
    from ogbench_b200 import Dataset, GCDataset, ReplayBuffer

    rng = np.random.default_rng(2)

    def transition(i):
        obs = {'state': np.full(3, i, np.float32), 'goal': np.full(2, -i, np.float32)}
        return dict(observations=obs, next_observations={k: v + 1 for k, v in obs.items()}, actions=rng.uniform(-1, 1, 2).astype(np.float32),
                    rewards=np.float64(i))
    rb = ReplayBuffer.create(transition(0), size=32, output='numpy')
    for i in range(1, 21):
        rb.add_transition(transition(i))
    batch = rb.sample(256)
    assert set(batch['observations']) == {'state', 'goal'}
    r = batch['rewards']
    assert np.array_equal(batch['observations']['state'], np.repeat(r[:, None], 3, 1).astype(np.float32))
    assert np.array_equal(batch['next_observations']['goal'], np.repeat(-r[:, None] + 1, 2, 1).astype(np.float32))
    assert set(np.unique(r)) <= set(float(i) for i in range(1, 21))
    base = toy_fields(3, ragged(3, 10, 4, 20), (4,), 2, np.float32)
    nested = dict(base, observations={'a': base['observations'], 'b': base['observations'][:, :2].copy()})
    with pytest.raises(NotImplementedError):                # the reference's augment() reads batch[key].shape (datasets.py:337)
        GCDataset(Dataset.create(**nested), cfg(p_aug=0.5))
