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
