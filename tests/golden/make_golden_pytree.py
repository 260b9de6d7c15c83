"""Golden vectors for PYTREE fields (nested dict observations), recorded from the UNMODIFIED reference sampler.

The reference maps its gathers over arbitrary pytrees (impls/utils/datasets.py:13, 54-56, 80, 344, 365-366); no shipped
OGBench dataset has nested observations, so these cases exist to pin the drop-in's behaviour for callers that do.
Same recipe as make_golden.py: the reference file runs under oracle/refshim.py (whose dict tree_map recurses like jax's),
np.random is seeded, outputs are stored flattened with '/' paths.  Run in the build container (needs /root/reference):

    python -m tests.golden.make_golden_pytree
"""

import json
import os

import numpy as np

from oracle import refshim
from tests.golden.make_golden import cfg, ragged, toy_fields

OUT_DIR = os.path.dirname(os.path.abspath(__file__))


def flatten(tree, prefix=''):
    out = {}
    for k in sorted(tree):
        path = f'{prefix}/{k}' if prefix else k
        if isinstance(tree[k], dict):
            out.update(flatten(tree[k], path))
        else:
            out[path] = tree[k]
    return out


def nest(flat):
    out = {}
    for path, v in flat.items():
        node = out
        parts = path.split('/')
        for p in parts[:-1]:
            node = node.setdefault(p, {})
        node[parts[-1]] = v
    return out


def pytree_fields(seed, lengths, compact=True, deep=False):
    base = toy_fields(seed, lengths, (5,), 2, np.float32, compact=compact)
    rng = np.random.default_rng(1000 + seed)
    n = len(base['terminals'])
    extra = rng.standard_normal((n, 3)).astype(np.float32)
    ids = rng.integers(0, 1000, size=(n,)).astype(np.int32)
    obs = {'state': base['observations'], 'task': {'feat': extra, 'id': ids}} if deep else {'state': base['observations'], 'feat': extra}
    fields = dict(base, observations=obs)
    if not compact:
        nxt = {k: np.concatenate([v[1:], v[-1:]]) for k, v in flatten(obs).items()}
        fields['next_observations'] = nest(nxt)
    return fields


def cases():
    # Only REGULAR datasets (explicit next_observations): for a compact dataset the reference synthesises next_observations
    # with a plain index into self._dict['observations'] (datasets.py:82), which raises TypeError for a dict, and frame
    # stacking requires a compact dataset (:208) -- so these are the pytree cases the reference can run at all.
    L = ragged(21, 8, 3, 50)
    yield dict(name='pytree_gc_regular', kind='gc', fields=pytree_fields(1, L, compact=False), cfg=cfg(), B=48, seed=201)
    yield dict(name='pytree_hgc_regular_deep', kind='hgc', fields=pytree_fields(2, L, compact=False, deep=True), cfg=cfg(subgoal_steps=4),
               B=40, seed=202)
    yield dict(name='pytree_gc_regular_eval', kind='gc', fields=pytree_fields(4, L, compact=False), cfg=cfg(p_aug=None), B=24, seed=204,
               evaluation=True)


def copy_tree(tree):
    return {k: copy_tree(v) if isinstance(v, dict) else v.copy() for k, v in tree.items()}


def run_reference(case):
    ref = refshim.load_reference_datasets_module()
    cls = {'gc': ref.GCDataset, 'hgc': ref.HGCDataset}[case['kind']]
    outs = []
    for preprocess in (True, False):
        ds = ref.Dataset.create(**copy_tree(case['fields']))
        sampler = cls(ds, dict(case['cfg']), preprocess_frame_stack=preprocess)
        np.random.seed(case['seed'])
        outs.append(flatten(sampler.sample(case['B'], evaluation=case.get('evaluation', False))))
    assert outs[0].keys() == outs[1].keys()
    for k in outs[0]:
        assert np.array_equal(outs[0][k], outs[1][k]) and outs[0][k].dtype == outs[1][k].dtype, k
    return outs[-1]


def main():
    for case in cases():
        out = run_reference(case)
        payload = {'meta': np.array(json.dumps(dict(name=case['name'], kind=case['kind'], cfg=case['cfg'], B=case['B'], seed=case['seed'],
                                                    evaluation=case.get('evaluation', False), numpy=np.__version__)))}
        for k, v in flatten(case['fields']).items():
            payload['field/' + k] = v
        for k, v in out.items():
            payload['out/' + k] = v
        path = os.path.join(OUT_DIR, case['name'] + '.npz')
        np.savez_compressed(path, **payload)
        print(f"{case['name']:28s} keys={len(out):2d} {os.path.getsize(path) / 1024:7.1f} KiB  {sorted(out)[:6]} ...")


if __name__ == '__main__':
    main()
