"""Load the committed golden fixtures (made by tests/golden/make_golden.py from the unmodified reference)."""

import glob
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def case_names():
    names = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, '*.npz')))
    return [n for n in names if not n.startswith(('rb_', 'loader_', 'pytree_'))]  # those fixtures have their own tests


def load_case(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + '.npz'))
    meta = json.loads(str(z['meta']))
    fields = {k[len('field/'):]: z[k] for k in z.files if k.startswith('field/')}
    out = {k[len('out/'):]: z[k] for k in z.files if k.startswith('out/')}
    log = [(kind, z[f'draw/{i}']) for i, kind in enumerate(meta['draw_kinds'])]
    idxs = z['idxs'] if 'idxs' in z.files else None
    return dict(meta=meta, fields=fields, out=out, log=log, idxs=idxs, cfg=meta['cfg'], kind=meta['kind'],
                B=meta['B'], evaluation=meta['evaluation'], k=meta.get('k'))


def assert_batches_identical(got, want, label=''):
    assert set(got.keys()) == set(want.keys()), f'{label}: key sets differ: {set(got) ^ set(want)}'
    for k in want:
        g, w = np.asarray(got[k]), np.asarray(want[k])
        assert g.dtype == w.dtype, f'{label}{k}: dtype {g.dtype} != {w.dtype}'
        assert g.shape == w.shape, f'{label}{k}: shape {g.shape} != {w.shape}'
        assert np.array_equal(g, w), f'{label}{k}: values differ ({np.sum(g != w)} of {g.size})'
