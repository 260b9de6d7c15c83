"""Randomised differential run: device sampler (rng='numpy') against the oracle over random dataset layouts, sampler
settings, classes and batch sizes.  usage: python scratch/fuzz.py [n_cases] [seed]"""
import sys, os, time, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests.golden.make_golden import cfg, toy_fields
from tests.golden_util import assert_batches_identical
from tests.gpu_util import device_sampler, oracle_with_draws, to_host

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 200
master = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)


def mix(r):
    """(p_cur, p_traj, p_rand) with frequent zeros / ones"""
    kind = r.integers(0, 6)
    if kind == 0: return (1.0, 0.0, 0.0)
    if kind == 1: return (0.0, 1.0, 0.0)
    if kind == 2: return (0.0, 0.0, 1.0)
    if kind == 3: return (0.2, 0.5, 0.3)
    if kind == 4: return (0.0, 0.5, 0.5)
    a = r.integers(1, 8); b = r.integers(1, 9 - a)
    return (a / 10, b / 10, (10 - a - b) / 10)


failures = 0
t0 = time.time()
for case in range(n_cases):
    r = np.random.default_rng(master.integers(0, 2**31))
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
    compact = bool(r.random() < 0.8) or fs is not None
    fields = toy_fields(int(r.integers(0, 10**6)), lengths, obs_shape, int(r.integers(1, 12)), obs_dtype, compact=compact,
                        oracle_rep_dim=(int(r.integers(1, 9)) if r.random() < 0.2 else None), extra=bool(r.random() < 0.2))
    vm, am = mix(r), mix(r)
    over = dict(value_p_curgoal=vm[0], value_p_trajgoal=vm[1], value_p_randomgoal=vm[2], value_geom_sample=bool(r.integers(0, 2)),
                actor_p_curgoal=am[0], actor_p_trajgoal=am[1], actor_p_randomgoal=am[2], actor_geom_sample=bool(r.integers(0, 2)),
                discount=[0.9, 0.99, 0.995, 0.999][r.integers(0, 4)], gc_negative=bool(r.integers(0, 2)), frame_stack=fs,
                p_aug=[None, 0.0, 0.5, 1.0][r.integers(0, 4)])
    if kind == 'hgc':
        over['subgoal_steps'] = int(r.integers(1, 30))
        if r.random() < 0.3: over['low_discount'] = [0.9, 0.95][r.integers(0, 2)]
        if r.random() < 0.3: over['low_subgoal_steps'] = int(r.integers(1, 6))
        if r.random() < 0.2: over['high_subgoal_steps'] = int(r.integers(1, 40))
    config = cfg(**over)
    okind = kind
    B = int([1, 7, 32, 33, 257, 1024, 3000][r.integers(0, 7)]) if not pixel else int(r.integers(1, 40))
    output = ['device', 'numpy'][r.integers(0, 2)]
    try:
        oracle_only = bool(os.environ.get('FUZZ_ORACLE_ONLY'))
        dedup = bool(r.integers(0, 2))
        sampler = None if oracle_only else device_sampler(fields, config, okind, rng='numpy', output=output, dedup=dedup)
        for it in range(3):
            evaluation = it == 1
            given = None
            if it == 2 and r.random() < 0.5:
                from oracle.replay_oracle import OracleSampler
                vt = OracleSampler(fields, config, okind).valid_table
                if vt is None:
                    vt = np.arange(len(fields['terminals']))
                given = vt[r.integers(0, len(vt), size=B)]
            seed = int(r.integers(0, 2**31))
            np.random.seed(seed)
            _, want = oracle_with_draws(fields, config, okind, B, idxs=given, evaluation=evaluation)
            if oracle_only:
                continue
            np.random.seed(seed)
            got = to_host(sampler.sample(B, idxs=given, evaluation=evaluation))
            assert_batches_identical(got, want, label=f'case {case}/{it}: ')
        if not oracle_only:
            # on-device RNG: the kernel's draws rebuilt in numpy feed the oracle; identical batch demanded
            from oracle import philox_np
            from oracle.replay_oracle import DrawsSource, OracleSampler
            from tests.test_gpu_philox import goal_sets_for
            pseed, pstream = int(r.integers(0, 2**62)), int(r.integers(0, 1000))
            dev = device_sampler(fields, config, okind, seed=pseed, stream_id=pstream, output=output, dedup=dedup)
            oracle = OracleSampler(fields, config, okind)
            n_choices = len(oracle.valid_table) if oracle.valid_table is not None else len(fields['terminals'])
            for call in range(2):
                evaluation = call == 1
                got = to_host(dev.sample(B, evaluation=evaluation))
                aug = config['p_aug'] is not None and not evaluation
                draws, knife = philox_np.philox_draws(pseed, pstream, call, B, n_choices, goal_sets_for(config, okind), aug, config['p_aug'] or 0.0)
                src = DrawsSource(draws)
                want = oracle.sample(B, evaluation=evaluation, source=src)
                assert src.exhausted()
                assert set(got) == set(want)
                for k in want:
                    assert got[k].dtype == want[k].dtype and got[k].shape == want[k].shape, k
                    assert np.array_equal(got[k][~knife], want[k][~knife]), f'philox mode, call {call}, key {k}'
    except Exception as exc:
        failures += 1
        print(f'FAIL case {case}: kind={kind} pixel={pixel} obs={obs_shape} {np.dtype(obs_dtype).name} fs={fs} compact={compact} B={B} '
              f'lengths[{lo},{hi}]x{n_traj} cfg={over}\n   {type(exc).__name__}: {str(exc)[:300]}')
        if failures <= 3:
            traceback.print_exc(limit=3)
print(f'{n_cases} cases, {failures} failures, {time.time() - t0:.1f} s')
sys.exit(1 if failures else 0)
