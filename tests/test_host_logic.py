"""Host-side logic that needs no GPU: sharding, the host draw schedule, synthetic layouts, Philox restatement."""

import numpy as np
import pytest

from ogbench_b200 import sharding, synthetic
from oracle import philox_np
from oracle.replay_oracle import OracleSampler, trajectory_bounds
from tests.golden.make_golden import ragged, toy_fields


def test_shard_bounds_are_trajectory_aligned_and_cover():
    lengths = ragged(3, 57, 2, 40)
    fields = toy_fields(3, lengths, (4,), 2, np.float32)
    n = len(fields['terminals'])
    for world in (1, 2, 3, 8):
        bounds = sharding.shard_bounds(fields['terminals'], world)
        assert bounds[0][0] == 0 and bounds[-1][1] == n
        ends = set((np.cumsum(lengths) - 1).tolist())
        for (a, b), (c, _) in zip(bounds, bounds[1:] + [(n, n)]):
            assert b == c and b > a
            assert (b - 1) in ends
        sizes = [b - a for a, b in bounds]
        assert max(sizes) - min(sizes) <= 2 * lengths.max()
        for rank in range(world):
            shard = sharding.take_shard(fields, rank, world)
            terminal_locs, _ = trajectory_bounds(shard['terminals'])
            assert terminal_locs[-1] == len(shard['terminals']) - 1  # each shard is a valid dataset (datasets.py:188)


def test_shard_more_ranks_than_trajectories_raises():
    fields = toy_fields(1, np.array([5, 6]), (2,), 2, np.float32)
    with pytest.raises(ValueError):
        sharding.shard_bounds(fields['terminals'], 3)


def test_synthetic_layout_matches_compact_loader():
    w = synthetic.WORKLOADS['c2']
    f = synthetic.host_fields(w, episodes=3)
    t, v = f['terminals'], f['valids']
    assert f['observations'].shape == (3 * 1001, 29) and f['actions'].shape == (3 * 1001, 8)
    assert t.sum() == 6 and v.sum() == 3 * 1000
    assert np.array_equal(np.nonzero(t)[0], [999, 1000, 2000, 2001, 3001, 3002])
    assert np.array_equal(np.nonzero(v == 0)[0], [1000, 2001, 3002])


def test_algorithmic_bytes_formulas():
    """one source row read + one output row written per unique index vector, plus the small fields."""
    def gc_bytes(d, a):
        return (4 * d + a + 8) + (4 * d + a + 8 + 16)
    assert gc_bytes(8, 8) == synthetic.WORKLOADS['c1'].bytes_per_transition
    assert gc_bytes(116, 32) == synthetic.WORKLOADS['c2'].bytes_per_transition
    assert gc_bytes(220, 20) == synthetic.WORKLOADS['c5'].bytes_per_transition
    assert (7 * 276 + 84 + 8) + (7 * 276 + 84 + 8 + 3 * 8 + 6 * 8) == synthetic.WORKLOADS['c3'].bytes_per_transition
    assert (10 * 12288 + 28) + (4 * 36864 + 20 + 8 + 16) == synthetic.WORKLOADS['c4'].bytes_per_transition


def test_philox_known_answer():
    """Random123 known-answer vectors for Philox4x32-10."""
    out = philox_np.philox4x32_10(0, 0, 0, 0, 0, 0)
    assert [int(x) for x in out] == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    f = 0xFFFFFFFF
    out = philox_np.philox4x32_10(f, f, f, f, f, f)
    assert [int(x) for x in out] == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    out = philox_np.philox4x32_10(0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344, 0xA4093822, 0x299F31D0)
    assert [int(x) for x in out] == [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_philox_draws_feed_the_oracle():
    lengths = ragged(5, 20, 2, 50)
    fields = toy_fields(5, lengths, (3,), 2, np.float32)
    from tests.golden.make_golden import cfg
    from oracle.replay_oracle import DrawsSource

    config = cfg()
    o = OracleSampler(fields, config, 'gc')
    n_choices = len(o.valid_table)
    draws, _ = philox_np.philox_draws(7, 3, 11, 64, n_choices, [(0, True, 0.99, False), (2, False, 0.99, False)], True, 0.0)
    src = DrawsSource(draws)
    batch = o.sample(64, source=src)
    assert src.exhausted()
    assert batch['observations'].shape == (64, 3)
    assert draws.idx_pos.min() >= 0 and draws.idx_pos.max() < n_choices
    assert draws.goals[0].offset.min() >= 1
    d2, _ = philox_np.philox_draws(7, 4, 11, 64, n_choices, [(0, True, 0.99, False), (2, False, 0.99, False)], True, 0.0)
    assert not np.array_equal(d2.idx_pos, draws.idx_pos)  # another stream id -> another stream


def test_host_draw_schedule_matches_reference_order():
    """rng='numpy' consumes np.random exactly as the reference does: checked against the golden draw logs without a GPU."""
    from ogbench_b200.datasets import GCDataset, HGCDataset
    from tests.golden_util import case_names, load_case
    from oracle.refshim import DrawRecorder

    for name in case_names():
        if name.startswith(('trl_', 'atc_')):
            continue  # TRL has no host draw schedule (midpoints need device rows); ATC's needs the native anchor count
        case = load_case(name)
        cls = GCDataset if case['kind'] == 'gc' else HGCDataset
        shell = cls.__new__(cls)  # the draw schedule is pure host logic; bypass the device-backed constructor
        shell.config = case['cfg']
        shell._trl = False
        shell._n_choices = int(np.sum(case['fields']['valids'] > 0)) if 'valids' in case['fields'] else len(case['fields']['terminals'])
        np.random.seed(case['meta']['seed'])
        with DrawRecorder() as rec:
            shell._host_draws(case['B'], case['idxs'], case['evaluation'])
        assert [k for k, _ in rec.log] == case['meta']['draw_kinds'], name
        for (_, got), (_, want) in zip(rec.log, case['log']):
            assert np.array_equal(got, want), name


def test_cpulist_parser_and_bench_config_dict():
    import bench
    from ogbench_b200 import dist_util

    assert dist_util._parse_cpulist('0-3,8,10-11\n') == [0, 1, 2, 3, 8, 10, 11]
    assert dist_util._parse_cpulist('') == []
    # both arms of bench.py print this dict: a pure function of the workload key and batches_per_launch
    c2 = bench.config_dict('c2', 1024)
    assert c2 == bench.config_dict('c2', 1024) and c2['key'] == 'c2' and c2['transitions_per_step_per_gpu'] == 1 << 20
    assert c2['placement'] == 'replica per GPU' and 'shard' in bench.config_dict('c5', 256)['placement']
    assert 'L2' in c2['l2']


def test_bench_refuses_stale_traffic_capture(tmp_path, monkeypatch):
    """roofline.traffic comes from the committed ncu capture only while that capture still describes the kernel that ran."""
    import json

    import bench

    (tmp_path / 'profiles').mkdir()
    entry = {'kernel': 'void relabel_gather_kernel<0, 0>(FusedParams)', 'dram_bytes_per_launch': 1e9, 'duration_us_under_ncu': 190.0,
             'transitions_per_launch': 1 << 20}
    (tmp_path / 'profiles' / 'traffic.json').write_text(json.dumps({'c2': entry}))
    monkeypatch.setattr(bench, 'ROOT', str(tmp_path))
    ok, note = bench.committed_traffic('c2', 'relabel_gather_kernel', 0.185, 1 << 20)
    assert ok == 1e9 and 'under ncu' in note
    half, _ = bench.committed_traffic('c2', 'relabel_gather_kernel', 0.0925, 1 << 19)      # scaled to the launch size
    assert half == 5e8
    stale, why = bench.committed_traffic('c2', 'relabel_gather_kernel', 0.150, 1 << 20)     # the kernel got 21 % faster since
    assert stale is None and 'stale' in why
    other, why = bench.committed_traffic('c2', 'gather_rows_async_kernel', 0.185, 1 << 20)
    assert other is None and 'capture is of' in why
    assert bench.committed_traffic('c9', 'x', 1.0, 1)[0] is None


def test_workload_catalogue_is_consistent():
    """Algorithmic bytes per transition of the bench workloads follow from their shapes."""
    from ogbench_b200 import synthetic

    w = synthetic.WORKLOADS
    obs = {k: int(np.prod(v.obs_shape)) * (1 if v.obs_dtype == 'uint8' else 4) for k, v in w.items()}
    act = {k: v.act_dim * 4 for k, v in w.items()}
    for k in ('c1', 'c2', 'c5'):      # GC: 4 row gathers in and out, actions, terminals + valids, masks + rewards
        assert w[k].bytes_per_transition == (4 * obs[k] + act[k] + 8) + (4 * obs[k] + act[k] + 8 + 16), k
    assert w['c3'].bytes_per_transition == (7 * obs['c3'] + act['c3'] + 8) + (7 * obs['c3'] + act['c3'] + 8 + 9 * 8)
    assert w['c4'].bytes_per_transition == (10 * obs['c4'] + act['c4'] + 8) + (4 * 3 * obs['c4'] + act['c4'] + 8 + 16)
    assert w['c5'].rows == 12_512_500 and w['c3'].rows == 4_001_000


def test_prefetcher_order_errors_and_close():
    """Prefetcher hands out dataset.sample() results in call order, passes keyword arguments on, surfaces the worker's
    exception at the consumer, refuses the np.random replay mode and stops cleanly."""
    import time

    from ogbench_b200.prefetch import Prefetcher

    class Fake:
        rng = 'philox'

        def __init__(self, fail_at=None):
            self.calls, self.fail_at = 0, fail_at

        def sample(self, batch_size, evaluation=False):
            if self.calls == self.fail_at:
                raise RuntimeError('boom')
            self.calls += 1
            return {'n': self.calls, 'batch_size': batch_size, 'evaluation': evaluation}

    fake = Fake()
    with Prefetcher(fake, 7, depth=3, evaluation=True) as batches:
        got = [next(batches) for _ in range(20)]
        assert [g['n'] for g in got] == list(range(1, 21))
        assert all(g['batch_size'] == 7 and g['evaluation'] for g in got)
        time.sleep(0.05)
        assert fake.calls <= 20 + 3 + 2          # at most depth queued + one in hand + one launched ahead of it
        worker = batches._thread
    assert not worker.is_alive()
    with pytest.raises(StopIteration):
        next(batches)

    failing = Prefetcher(Fake(fail_at=2), 4)
    assert next(failing)['n'] == 1 and next(failing)['n'] == 2
    with pytest.raises(RuntimeError, match='boom'):
        next(failing)
    failing.close()

    numpy_mode = Fake()
    numpy_mode.rng = 'numpy'
    with pytest.raises(ValueError):
        Prefetcher(numpy_mode, 4)


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the GPU arm) needs no GPU: one JSON line with the
    contract's keys, `impl: reference`, a cpu_baseline describing the run and an e2e that repeats the line's value."""
    import json
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, 'bench.py'), '--impl', 'reference', '--config', 'c1', '--steps', '2', '--warmup', '1'],
                         capture_output=True, text=True, timeout=300, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, out.stdout
    line = json.loads(lines[0])
    for key in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling', 'vs_baseline',
                'dtype', 'data', 'config', 'cpu_baseline', 'e2e'):
        assert key in line, key
    assert line['impl'] == 'reference' and line['metric'] == 'relabeled transitions/sec' and line['unit'] == 'transitions/s'
    assert line['value'] > 0 and line['higher_is_better'] is True
    assert line['cpu_baseline']['kind'] == 'port' and line['cpu_baseline']['cores'] >= 1 and line['cpu_baseline']['value'] == line['value']
    assert line['e2e'] == {'value': line['value'], 'unit': line['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert 'workload' in line['config']


def test_fuzz_generator_runs_on_the_oracle_alone():
    """The random-case generator of the GPU fuzz test yields configurations the oracle accepts (no GPU needed here)."""
    from tests.fuzz_util import run_case

    master = np.random.default_rng(5)
    kinds = set()
    for _ in range(60):
        case = run_case(np.random.default_rng(int(master.integers(0, 2**31))), oracle_only=True)
        kinds.add((case['kind'], case['fields']['observations'].ndim > 2))
    assert {k for k, _ in kinds} == {'gc', 'hgc', 'atc', 'trl'}


def test_pinned_pool_size_classes():
    """Host blocks of output='numpy' are rounded up by at most 12.5 % (eight size classes per power of two)."""
    from ogbench_b200.datasets import _PinnedPool

    for n in (1, 4096, 4097, 5000, 65536, 65537, 545_259_520, (1 << 30) + 1):
        b = _PinnedPool.bucket_of(n)
        assert b >= n and b >= 4096 and (b <= 4096 or b <= n * 1.125 + 1), (n, b)
        assert _PinnedPool.bucket_of(b) == b               # a size class maps to itself: returned blocks are found again
