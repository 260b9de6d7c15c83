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
