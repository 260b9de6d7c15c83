"""Golden vectors for ReplayBuffer (datasets.py:86-146) from the unmodified reference: a scripted sequence of
add_transition / sample / clear calls with the sampled batches recorded.  -> tests/golden/rb_script.npz"""

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refshim  # noqa: E402


def make_transitions(n, seed=7):
    rng = np.random.default_rng(seed)
    return [dict(observations=rng.standard_normal(5).astype(np.float32), actions=rng.uniform(-1, 1, 2).astype(np.float32),
                 rewards=np.float64(rng.standard_normal()), terminals=np.float32(rng.integers(2)), masks=np.float32(1.0),
                 next_observations=rng.standard_normal(5).astype(np.float32)) for _ in range(n)]


SCRIPT = [('add', 30), ('sample', 16, 1), ('add', 25), ('sample', 16, 2), ('add', 60), ('sample', 32, 3), ('clear',), ('add', 3), ('sample', 8, 4)]
CAPACITY = 50


def run(buffer, transitions, sample_fn):
    outs, t = [], 0
    for step in SCRIPT:
        if step[0] == 'add':
            for _ in range(step[1]):
                buffer.add_transition(transitions[t])
                t += 1
        elif step[0] == 'clear':
            buffer.clear()
        else:
            np.random.seed(step[2])
            outs.append(sample_fn(buffer, step[1]))
    return outs


def main():
    ref = refshim.load_reference_datasets_module()
    transitions = make_transitions(200)
    rb = ref.ReplayBuffer.create(transitions[0], size=CAPACITY)
    outs = run(rb, transitions, lambda b, n: b.sample(n))
    payload = {'meta': np.array(json.dumps(dict(script=SCRIPT, capacity=CAPACITY, n_transitions=200)))}
    for i, out in enumerate(outs):
        for k, v in out.items():
            payload[f'out/{i}/{k}'] = v
    # create_from_initial_dataset
    init = {k: np.stack([t[k] for t in transitions[:20]]) for k in transitions[0]}
    rb2 = ref.ReplayBuffer.create_from_initial_dataset(init, size=40)
    for t in transitions[20:50]:
        rb2.add_transition(t)
    np.random.seed(9)
    out = rb2.sample(24)
    for k, v in out.items():
        payload[f'init_out/{k}'] = v
    payload['init_state'] = np.array([rb2.size, rb2.pointer, rb2.max_size])
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'rb_script.npz'), **payload)
    print('rb_script.npz written', len(outs), 'samples')


if __name__ == '__main__':
    main()
