"""ReplayBuffer (datasets.py:86-146): oracle vs golden vectors from the reference (CPU), device vs golden (GPU)."""

import os

import numpy as np
import pytest

from oracle.replay_oracle import OracleReplayBuffer
from tests.golden.make_golden_replay import CAPACITY, make_transitions, run

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'rb_script.npz')


def load():
    z = np.load(GOLD)
    outs = {}
    for key in z.files:
        if key.startswith('out/'):
            _, i, name = key.split('/')
            outs.setdefault(int(i), {})[name] = z[key]
    init_out = {k[len('init_out/'):]: z[k] for k in z.files if k.startswith('init_out/')}
    return [outs[i] for i in sorted(outs)], init_out, z['init_state']


def check(got, want):
    assert set(got) == set(want)
    for k in want:
        g = np.asarray(got[k])
        assert g.dtype == want[k].dtype and g.shape == want[k].shape and np.array_equal(g, want[k]), k


def test_names_are_not_shadowed():
    assert 'rb_script' in os.path.basename(GOLD)


def test_oracle_replay_buffer_matches_reference():
    want, init_want, init_state = load()
    transitions = make_transitions(200)
    rb = OracleReplayBuffer(transitions[0], CAPACITY)
    got = run(rb, transitions, lambda b, n: b.sample(n))
    assert len(got) == len(want)
    for g, w in zip(got, want):
        check(g, w)
    init = {k: np.stack([t[k] for t in transitions[:20]]) for k in transitions[0]}
    rb2 = OracleReplayBuffer.from_initial_dataset(init, 40)
    for t in transitions[20:50]:
        rb2.add_transition(t)
    np.random.seed(9)
    check(rb2.sample(24), init_want)
    assert [rb2.size, rb2.pointer, rb2.max_size] == list(init_state)


@pytest.mark.gpu
def test_device_replay_buffer_matches_reference():
    from ogbench_b200 import ReplayBuffer

    want, init_want, init_state = load()
    transitions = make_transitions(200)
    rb = ReplayBuffer.create(transitions[0], size=CAPACITY, rng='numpy', output='numpy')
    got = run(rb, transitions, lambda b, n: b.sample(n))
    for g, w in zip(got, want):
        check(g, w)
    init = {k: np.stack([t[k] for t in transitions[:20]]) for k in transitions[0]}
    rb2 = ReplayBuffer.create_from_initial_dataset(init, size=40, rng='numpy')
    for t in transitions[20:50]:
        rb2.add_transition(t)
    np.random.seed(9)
    check(rb2.sample(24), init_want)
    assert [rb2.size, rb2.pointer, rb2.max_size] == list(init_state)
    # the on-device RNG mode only ever draws filled rows
    rb3 = ReplayBuffer.create(transitions[0], size=CAPACITY)
    with pytest.raises(ValueError):
        rb3.sample(4)                                            # empty buffer: the reference's randint(0) raises too
    marked = dict(transitions[0])
    for i in range(7):
        marked['rewards'] = np.float64(i + 1)
        rb3.add_transition(marked)
    out = np.asarray(rb3.sample(512)['rewards'])
    assert set(np.unique(out)) <= set(float(i + 1) for i in range(7)) and len(np.unique(out)) == 7


@pytest.mark.gpu
def test_replay_buffer_without_next_observations_tracks_row_writes():
    """Transitions without 'next_observations': get_subset synthesises observations[min(idx + 1, size - 1)] (datasets.py:78-83)
    with size = the rows filled so far.  A full buffer of <= 16-byte observations serves that key from the record's shadow
    copy of the next row, which every add_transition has to keep coherent (the written row's predecessor, and the last row,
    which shadows itself)."""
    from ogbench_b200 import ReplayBuffer

    rng = np.random.default_rng(4)

    def transition():
        return dict(observations=rng.standard_normal(2).astype(np.float32), actions=rng.uniform(-1, 1, 2).astype(np.float32),
                    terminals=np.float32(0.0), rewards=np.float32(1.0))

    def check(rb, mirror, filled):
        idxs = np.arange(filled)
        out = rb.sample(filled, idxs=idxs)
        assert np.array_equal(out['observations'], mirror[:filled])
        assert np.array_equal(out['next_observations'], mirror[np.minimum(idxs + 1, filled - 1)])

    # a buffer that is still filling up: the clamp follows the fill level (no shadow in play)
    rb = ReplayBuffer.create(transition(), size=16, output='numpy')
    mirror = np.zeros((16, 2), dtype=np.float32)
    for i in range(10):
        t = transition()
        mirror[i] = t['observations']
        rb.add_transition(t)
        check(rb, mirror, i + 1)
    # a dataset created full keeps the shadow copy; row writes through the C-ABI (what add_transition calls) must keep it
    # coherent, including the last row, which shadows itself
    import ctypes as C

    from ogbench_b200 import Dataset, _native

    n = 12
    init = {k: np.stack([transition()[k] for _ in range(n)]) for k in ('observations', 'actions', 'terminals', 'rewards')}
    ds = Dataset.create(**{k: v.copy() for k, v in init.items()})
    ds.output = 'numpy'
    sampler = ds._plain_sampler()
    mirror = init['observations'].copy()
    check(ds, mirror, n)
    names = list(init)
    for step in range(3 * n):
        row = (5 * step + 3) % n                 # hits every row, the first and the last included
        t = transition()
        mirror[row] = t['observations']
        ptrs = (C.c_void_p * len(names))()
        keep = [np.ascontiguousarray(t[k]) for k in names]
        for i, a in enumerate(keep):
            ptrs[i] = a.ctypes.data
        _native.check(_native.lib().ogb_sampler_write_row(sampler.ptr, row, ptrs, len(names)))
        check(ds, mirror, n)
