"""Random sampler cases for the differential tests: `run_case` builds a random dataset layout and sampler configuration
from a numpy Generator and compares the device sampler -- rng='numpy' and the on-device Philox mode -- with the oracle.
Used by tests/test_gpu_fuzz.py (fixed seeds, a few dozen cases) and scratch/fuzz.py (as many as asked for)."""

import numpy as np

from tests.golden.make_golden import cfg, toy_fields
from tests.golden_util import assert_batches_identical


def goal_mix(r):
    """(p_cur, p_traj, p_rand), with the degenerate mixes (a probability of zero or one) over-represented"""
    kind = r.integers(0, 6)
    if kind < 5:
        return [(1.0, 0.0, 0.0), (0.0, 1.0, 0.0), (0.0, 0.0, 1.0), (0.2, 0.5, 0.3), (0.0, 0.5, 0.5)][kind]
    a = int(r.integers(1, 8))
    b = int(r.integers(1, 9 - a))
    return (a / 10, b / 10, (10 - a - b) / 10)


def random_case(r):
    """-> dict(kind, fields, config, batch, output, dedup, summary)"""
    pixel = r.random() < 0.15
    kind = ['gc', 'gc', 'hgc'][r.integers(0, 3)]
    n_traj = int(r.integers(2, 12 if pixel else 120))
    lo = int(r.integers(2, 6))
    hi = int(lo + r.integers(0, 20 if pixel else [6, 40, 400][r.integers(0, 3)]))
    lengths = r.integers(lo, hi + 1, size=n_traj)
    if pixel:
        obs_shape = [(64, 64, 3), (32, 48, 3), (8, 8, 3), (20, 12, 4)][r.integers(0, 4)]
        obs_dtype, fs = np.uint8, [None, 2, 3, 4][r.integers(0, 4)]
    else:
        obs_shape = (int(r.integers(1, 90)),) if r.random() < 0.85 else (int(r.integers(1, 5)), int(r.integers(1, 7)))
        obs_dtype = [np.float32, np.float32, np.float64, np.float16, np.uint8, np.int32][r.integers(0, 6)]
        fs = None if r.random() < 0.8 else int(r.integers(2, 4))
    compact = bool(r.random() < 0.8) or fs is not None          # frame stacking asserts there are no next_observations
    fields = toy_fields(int(r.integers(0, 10**6)), lengths, obs_shape, int(r.integers(1, 12)), obs_dtype, compact=compact,
                        oracle_rep_dim=(int(r.integers(1, 9)) if r.random() < 0.2 else None), extra=bool(r.random() < 0.2))
    vm, am = goal_mix(r), goal_mix(r)
    over = dict(value_p_curgoal=vm[0], value_p_trajgoal=vm[1], value_p_randomgoal=vm[2], value_geom_sample=bool(r.integers(0, 2)),
                actor_p_curgoal=am[0], actor_p_trajgoal=am[1], actor_p_randomgoal=am[2], actor_geom_sample=bool(r.integers(0, 2)),
                discount=[0.9, 0.99, 0.995, 0.999][r.integers(0, 4)], gc_negative=bool(r.integers(0, 2)), frame_stack=fs,
                p_aug=[None, 0.0, 0.5, 1.0][r.integers(0, 4)])
    if kind == 'hgc':
        over['subgoal_steps'] = int(r.integers(1, 30))
        if r.random() < 0.3:
            over['low_discount'] = [0.9, 0.95][r.integers(0, 2)]
        if r.random() < 0.3:
            over['low_subgoal_steps'] = int(r.integers(1, 6))
        if r.random() < 0.2:
            over['high_subgoal_steps'] = int(r.integers(1, 40))
    batch = int(r.integers(1, 40)) if pixel else int([1, 7, 32, 33, 257, 1024, 3000][r.integers(0, 7)])
    output = ['device', 'numpy'][r.integers(0, 2)]
    dedup = bool(r.integers(0, 2))
    summary = (f'kind={kind} obs={obs_shape} {np.dtype(obs_dtype).name} fs={fs} compact={compact} B={batch} output={output} dedup={dedup} '
               f'lengths[{lo},{hi}]x{n_traj} cfg={over}')
    return dict(kind=kind, fields=fields, config=cfg(**over), batch=batch, output=output, dedup=dedup, summary=summary)


def run_atc_case(r, oracle_only=False):
    """ATCDataset.sample(batch, k) in rng='numpy' mode against the oracle (datasets.py:369-464)."""
    from oracle.replay_oracle import OracleATCSampler

    pixel = r.random() < 0.6
    lo = int(r.integers(3, 9))
    lengths = r.integers(lo, lo + int(r.integers(1, 30)), size=int(r.integers(2, 14 if pixel else 80)))
    if pixel:
        obs_shape, obs_dtype = [(64, 64, 3), (32, 48, 3), (8, 8, 3), (20, 12, 4)][r.integers(0, 4)], np.uint8
        fs = [None, 2, 3][r.integers(0, 3)]
    else:
        obs_shape, obs_dtype, fs = (int(r.integers(1, 60)),), np.float32, [None, None, 2][r.integers(0, 3)]
    compact = bool(r.random() < 0.8) or fs is not None
    fields = toy_fields(int(r.integers(0, 10**6)), lengths, obs_shape, int(r.integers(1, 6)), obs_dtype, compact=compact)
    config = dict(frame_stack=fs, p_aug=[None, 0.0, 0.5, 1.0][r.integers(0, 4)])
    if r.random() < 0.5:
        config['augment_padding'] = int(r.integers(0, 6))
    k = int(r.integers(1, lo - 1))                                  # every trajectory has anchors for this offset
    B = int(r.integers(1, 40)) if pixel else int([1, 33, 500][r.integers(0, 3)])
    output = ['device', 'numpy'][r.integers(0, 2)]
    summary = f'kind=atc obs={obs_shape} fs={fs} compact={compact} B={B} k={k} output={output} cfg={config}'
    oracle = OracleATCSampler(fields, config)
    try:
        sampler = None
        if not oracle_only:
            from tests.gpu_util import device_sampler, to_host

            sampler = device_sampler(fields, config, 'atc', rng='numpy', output=output)
            assert np.array_equal(sampler.get_valid_atc_idxs(k), oracle.valid_anchors(k)), 'anchor sets differ'
        for it in range(3):
            seed = int(r.integers(0, 2**31))
            np.random.seed(seed)
            want = oracle.sample(B, k, evaluation=(it == 1))
            if sampler is None:
                continue
            np.random.seed(seed)
            got = to_host(sampler.sample(B, k, evaluation=(it == 1)))
            assert_batches_identical(got, want, label=f'atc, call {it}: ')
    except AssertionError as exc:
        raise AssertionError(f'{summary}\n   {exc}') from exc
    return dict(kind='atc', fields=fields, config=config, batch=B, summary=summary)


def run_trl_case(r, oracle_only=False):
    """TRL branch of GCDataset.sample (datasets.py:198-204, 254-276) in validation mode: the device consumes the draws
    the oracle made from np.random, midpoint draw included."""
    from oracle.replay_oracle import OracleSampler

    pixel = r.random() < 0.2
    lengths = r.integers(4, 4 + int(r.integers(1, 60)), size=int(r.integers(2, 10 if pixel else 60)))
    if pixel:
        obs_shape, obs_dtype, fs = [(64, 64, 3), (8, 8, 3)][r.integers(0, 2)], np.uint8, [None, 3][r.integers(0, 2)]
    else:
        obs_shape, obs_dtype, fs = (int(r.integers(1, 70)),), [np.float32, np.float64][r.integers(0, 2)], None
    fields = toy_fields(int(r.integers(0, 10**6)), lengths, obs_shape, int(r.integers(1, 9)), obs_dtype,
                        oracle_rep_dim=(int(r.integers(1, 6)) if r.random() < 0.3 else None))
    am = goal_mix(r)
    config = cfg(agent_name=['trl', 'latent_trl', 'discrete_latent_trl'][r.integers(0, 3)], value_p_curgoal=0.0, value_p_trajgoal=1.0,
                 value_p_randomgoal=0.0, value_geom_sample=bool(r.integers(0, 2)), actor_p_curgoal=am[0], actor_p_trajgoal=am[1],
                 actor_p_randomgoal=am[2], actor_geom_sample=bool(r.integers(0, 2)), discount=[0.99, 0.999][r.integers(0, 2)],
                 gc_negative=bool(r.integers(0, 2)), frame_stack=fs, p_aug=[None, 0.0, 1.0][r.integers(0, 3)])
    B = int(r.integers(1, 30)) if pixel else int([1, 33, 700][r.integers(0, 3)])
    output = ['device', 'numpy'][r.integers(0, 2)]
    summary = f'kind=trl obs={obs_shape} fs={fs} B={B} output={output} cfg={config}'
    oracle = OracleSampler(fields, config, 'gc')
    try:
        sampler = None
        if not oracle_only:
            from tests.gpu_util import device_sampler, to_host

            sampler = device_sampler(fields, config, 'gc', output=output)
        for it in range(2):
            np.random.seed(int(r.integers(0, 2**31)))
            want = oracle.sample(B, evaluation=(it == 1))
            if sampler is None:
                continue
            got = to_host(sampler.sample(B, evaluation=(it == 1), draws=oracle.last_draws))
            assert_batches_identical(got, want, label=f'trl, call {it}: ')
    except AssertionError as exc:
        raise AssertionError(f'{summary}\n   {exc}') from exc
    return dict(kind='trl', fields=fields, config=config, batch=B, summary=summary)


def run_case(r, oracle_only=False):
    """One random case; raises AssertionError (message prefixed by the case summary) on any mismatch."""
    from oracle import philox_np
    from oracle.replay_oracle import DrawsSource, OracleSampler

    which = r.random()
    if which < 0.12:
        return run_atc_case(r, oracle_only)
    if which < 0.22:
        return run_trl_case(r, oracle_only)
    case = random_case(r)
    kind, fields, config, B = case['kind'], case['fields'], case['config'], case['batch']
    oracle = OracleSampler(fields, config, kind)
    rows = oracle.valid_table if oracle.valid_table is not None else np.arange(len(fields['terminals']))
    try:
        sampler = None
        if not oracle_only:
            from tests.gpu_util import device_sampler, to_host

            sampler = device_sampler(fields, config, kind, rng='numpy', output=case['output'], dedup=case['dedup'])
        for it in range(3):                                          # rng='numpy': the reference's own np.random calls
            evaluation = it == 1
            given = rows[r.integers(0, len(rows), size=B)] if (it == 2 and r.random() < 0.5) else None
            seed = int(r.integers(0, 2**31))
            np.random.seed(seed)
            want = oracle.sample(B, idxs=given, evaluation=evaluation)
            if sampler is None:
                continue
            np.random.seed(seed)
            got = to_host(sampler.sample(B, idxs=given, evaluation=evaluation))
            assert_batches_identical(got, want, label=f'numpy mode, call {it}: ')
        if oracle_only:
            return case
        # on-device RNG: the kernel's draws rebuilt in numpy feed the oracle; the identical batch is demanded
        from tests.test_gpu_philox import goal_sets_for

        pseed, pstream = int(r.integers(0, 2**62)), int(r.integers(0, 1000))
        dev = device_sampler(fields, config, kind, seed=pseed, stream_id=pstream, output=case['output'], dedup=case['dedup'])
        for call in range(2):
            evaluation = call == 1
            got = to_host(dev.sample(B, evaluation=evaluation))
            aug = config['p_aug'] is not None and not evaluation
            draws, knife = philox_np.philox_draws(pseed, pstream, call, B, len(rows), goal_sets_for(config, kind), aug, config['p_aug'] or 0.0)
            src = DrawsSource(draws)
            want = oracle.sample(B, evaluation=evaluation, source=src)
            assert src.exhausted()
            assert set(got) == set(want)
            for k in want:
                assert got[k].dtype == want[k].dtype and got[k].shape == want[k].shape, f'philox mode, call {call}, key {k}: dtype/shape'
                assert np.array_equal(got[k][~knife], want[k][~knife]), f'philox mode, call {call}, key {k}'
        # One case in four (vector observations only): a launch with more 32-row tiles than the gather kernels have warps, so
        # that tiles are handed out by the ticket counter -- byte-identical to the same launch with statically strided
        # tiles (debug bit 3), and two of its batches against the oracle.
        if fields['observations'].ndim == 2 and r.random() < 0.25:
            K = int(np.ceil(float(r.integers(90_000, 200_000)) / B))
            static = device_sampler(fields, config, kind, seed=pseed, stream_id=pstream, output=case['output'], dedup=case['dedup'])
            static.load_state_dict(dev.state_dict())
            static._sampler.set_debug(8)
            counter0 = dev.state_dict()['counter']
            many, many_static = to_host(dev.sample_many(K, B)), to_host(static.sample_many(K, B))
            assert set(many) == set(many_static)
            for k in many:
                assert np.array_equal(many[k], many_static[k]), f'ticket-scheduled launch of {K} x {B} rows differs from static tiles, key {k}'
            for b in sorted({0, K - 1, int(r.integers(0, K))}):
                draws, knife = philox_np.philox_draws(pseed, pstream, counter0 + b, B, len(rows), goal_sets_for(config, kind),
                                                      config['p_aug'] is not None, config['p_aug'] or 0.0)
                want = oracle.sample(B, source=DrawsSource(draws))
                for k in want:
                    assert np.array_equal(many[k][b][~knife], want[k][~knife]), f'big launch, batch {b} of {K}, key {k}'
    except AssertionError as exc:
        raise AssertionError(f'{case["summary"]}\n   {exc}') from exc
    return case
