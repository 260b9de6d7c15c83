"""Helpers shared by the GPU parity tests."""

import numpy as np

from oracle.replay_oracle import OracleSampler, ReplaySource


def to_host(batch):
    return {k: np.asarray(v) for k, v in batch.items()}


def device_sampler(fields, cfg, kind, **kw):
    from ogbench_b200 import ATCDataset, Dataset, GCDataset, HGCDataset

    ds = Dataset.create(**{k: v.copy() for k, v in fields.items()})
    cls = {'gc': GCDataset, 'hgc': HGCDataset, 'atc': ATCDataset}[kind]
    if kind == 'atc':
        kw.pop('dedup', None)
    return cls(ds, cfg, **kw)


def oracle_with_draws(fields, cfg, kind, B, idxs=None, evaluation=False, source=None):
    o = OracleSampler(fields, cfg, kind)
    batch = o.sample(B, idxs=idxs, evaluation=evaluation, source=source)
    return o, batch


def draws_from_log(case):
    """Structured draws of a golden case: replay the reference's log through the oracle."""
    src = ReplaySource(case['log'])
    if case['kind'] == 'atc':
        from oracle.replay_oracle import OracleATCSampler

        o = OracleATCSampler(case['fields'], case['cfg'])
        o.sample(case['B'], case['k'], evaluation=case['evaluation'], source=src)
    else:
        o = OracleSampler(case['fields'], case['cfg'], case['kind'])
        o.sample(case['B'], idxs=case['idxs'], evaluation=case['evaluation'], source=src)
    assert src.exhausted()
    return o.last_draws


def device_sample(sampler, case, **kw):
    if case['kind'] == 'atc':
        return sampler.sample(case['B'], case['k'], evaluation=case['evaluation'], **kw)
    return sampler.sample(case['B'], idxs=case['idxs'], evaluation=case['evaluation'], **kw)
