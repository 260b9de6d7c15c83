"""Pin the oracle: it must reproduce the reference's outputs bit-for-bit from the reference's recorded draws."""

import numpy as np
import pytest

from oracle import refshim
from oracle.replay_oracle import OracleATCSampler, OracleSampler, ReplaySource
from tests.golden_util import assert_batches_identical, case_names, load_case

CASES = case_names()


def oracle_sample(case, source):
    if case['kind'] == 'atc':
        return OracleATCSampler(case['fields'], case['cfg']).sample(case['B'], case['k'], evaluation=case['evaluation'], source=source)
    sampler = OracleSampler(case['fields'], case['cfg'], case['kind'])
    return sampler.sample(case['B'], idxs=case['idxs'], evaluation=case['evaluation'], source=source)


def test_fixtures_present():
    assert len(CASES) >= 20


@pytest.mark.parametrize('name', CASES)
def test_oracle_matches_golden(name):
    case = load_case(name)
    src = ReplaySource(case['log'])
    got = oracle_sample(case, src)
    assert src.exhausted(), 'oracle consumed fewer draws than the reference'
    assert_batches_identical(got, case['out'], label=name + ':')


@pytest.mark.parametrize('name', CASES)
def test_oracle_global_stream_matches_golden(name):
    """Same check through the global np.random stream: seed -> identical batch, no recording involved."""
    case = load_case(name)
    np.random.seed(case['meta']['seed'])
    got = oracle_sample(case, None)
    assert_batches_identical(got, case['out'], label=name + ':')


@pytest.mark.skipif(not refshim.reference_available(), reason='reference tree not mounted')
@pytest.mark.parametrize('kind', ['gc', 'hgc'])
@pytest.mark.parametrize('seed', [0, 1, 2])
def test_oracle_matches_live_reference(kind, seed):
    """Randomised live cross-check against the unmodified reference file (build container only)."""
    from tests.golden.make_golden import cfg, ragged, toy_fields

    ref = refshim.load_reference_datasets_module()
    rng = np.random.default_rng(1000 + seed)
    pixel = bool(seed % 2)
    lengths = ragged(seed, 6, 2, 12 if pixel else 80)
    fields = toy_fields(seed, lengths, (8, 8, 3) if pixel else (4,), 3, np.uint8 if pixel else np.float32)
    config = cfg(
        value_geom_sample=bool(rng.integers(2)), actor_geom_sample=bool(rng.integers(2)),
        actor_p_curgoal=0.1, actor_p_trajgoal=0.6, actor_p_randomgoal=0.3, gc_negative=bool(rng.integers(2)),
        frame_stack=3 if pixel else None, p_aug=0.5, subgoal_steps=int(rng.integers(1, 9)),
    )
    cls = ref.GCDataset if kind == 'gc' else ref.HGCDataset
    theirs = cls(ref.Dataset.create(**{k: v.copy() for k, v in fields.items()}), dict(config), preprocess_frame_stack=False)
    ours = OracleSampler(fields, config, kind)
    for it in range(5):
        np.random.seed(7 * seed + it)
        want = theirs.sample(33)
        np.random.seed(7 * seed + it)
        got = ours.sample(33)
        assert_batches_identical(got, want, label=f'{kind}/{seed}/{it}:')


def _pad_and_slice(img, cy, cx, p):
    """datasets.py:17-27 spelled out with numpy: edge padding, then a slice of the original shape at (cy, cx)."""
    padded = np.pad(img, ((p, p), (p, p), (0, 0)), mode='edge')
    h, w, _ = img.shape
    cy, cx = int(np.clip(cy, 0, padded.shape[0] - h)), int(np.clip(cx, 0, padded.shape[1] - w))    # lax.dynamic_slice clamps
    return padded[cy:cy + h, cx:cx + w]


@pytest.mark.parametrize('shape,dtype', [((8, 8, 3), np.uint8), ((7, 5, 3), np.uint8), ((6, 9, 1), np.float32), ((64, 64, 9), np.uint8)])
@pytest.mark.parametrize('padding', [1, 3, 4])
def test_crop_closed_form_equals_pad_and_slice(shape, dtype, padding):
    """The oracle's closed form of the augmentation (clip(y + cy - p), clip(x + cx - p)) against the literal
    pad(mode='edge') + slice for EVERY shift the reference can draw (randint(0, 2p + 1), datasets.py:333), and -- when the
    reference tree is mounted -- against the reference's own random_crop / batched_random_crop bodies executed on the
    numpy stand-ins of jnp.pad, lax.dynamic_slice and vmap (oracle/refshim.py)."""
    from oracle.replay_oracle import shifted_edge_crop

    rng = np.random.default_rng(padding * 100 + shape[0])
    shifts = np.array([(cy, cx) for cy in range(2 * padding + 1) for cx in range(2 * padding + 1)], dtype=np.int64)
    b = len(shifts)
    imgs = (rng.integers(0, 256, (b, *shape)).astype(dtype) if np.dtype(dtype) == np.uint8
            else rng.standard_normal((b, *shape)).astype(dtype))
    got = shifted_edge_crop(imgs, shifts, padding)
    want = np.stack([_pad_and_slice(imgs[i], shifts[i, 0], shifts[i, 1], padding) for i in range(b)])
    assert got.dtype == want.dtype and np.array_equal(got, want)
    if refshim.reference_available():
        ref = refshim.load_reference_datasets_module()
        crop_froms = np.concatenate([shifts, np.zeros((b, 1), dtype=np.int64)], axis=1)        # (cy, cx, 0), datasets.py:334
        theirs = ref.batched_random_crop(imgs, crop_froms, padding)
        assert ref.batched_random_crop is ref._literal_batched_random_crop
        assert np.asarray(theirs).dtype == got.dtype and np.array_equal(np.asarray(theirs), got)
