"""GCDataset(..., lookahead=K) and jax_compat=True.

lookahead: the reference's loop calls `sample(batch_size)` once per step (impls/main.py:202); with lookahead=K those calls
pop successive batches of one K-batch launch.  The bar is the usual one: the sequence of batches must be bit-identical to
the sequence direct calls return, whatever is mixed in between (other batch sizes, evaluation batches, given idxs,
sample_many, checkpoints).

jax_compat: scalar keys as float32 / int32, equal to numpy's astype of the float64 / int64 keys (what `jit` does to the
reference's arrays with x64 off, impls/main.py:204-207)."""

import numpy as np
import pytest

from tests.golden.make_golden import cfg, ragged, toy_fields
from tests.gpu_util import device_sampler, to_host

pytestmark = pytest.mark.gpu


def _same(a, b, label):
    assert set(a) == set(b), label
    for k in a:
        x, y = np.asarray(a[k]), np.asarray(b[k])
        assert x.dtype == y.dtype and x.shape == y.shape and np.array_equal(x, y), (label, k)


@pytest.mark.parametrize('kind,over', [('gc', {}), ('hgc', dict(subgoal_steps=4)), ('gc', dict(frame_stack=2, p_aug=0.5))])
@pytest.mark.parametrize('output', ['device', 'numpy'])
def test_lookahead_returns_the_direct_call_sequence(kind, over, output):
    pixel = 'frame_stack' in over
    fields = toy_fields(31, ragged(31, 12 if pixel else 80, 3, 40), (8, 8, 3) if pixel else (11,), 3, np.uint8 if pixel else np.float32)
    config = cfg(**over)
    direct = device_sampler(fields, config, kind, seed=21, stream_id=2, output=output)
    ahead = device_sampler(fields, config, kind, seed=21, stream_id=2, output=output, lookahead=5)
    rng = np.random.default_rng(0)
    n = len(fields['terminals'])
    valid = np.nonzero(fields['valids'] > 0)[0]
    script = ([('s', 64, False)] * 7 + [('s', 64, True)] + [('s', 64, False)] * 4 + [('s', 17, False)] * 6 + [('idxs', rng.choice(valid, 9))]
              + [('s', 17, False)] * 2 + [('many', 3, 20)] + [('s', 64, False)] * 11 + [('goals', rng.choice(valid, 5))] + [('s', 64, False)] * 3)
    for step, op in enumerate(script):
        assert direct.state_dict() == ahead.state_dict(), step
        if op[0] == 's':
            _same(to_host(direct.sample(op[1], evaluation=op[2])), to_host(ahead.sample(op[1], evaluation=op[2])), (step, op))
        elif op[0] == 'idxs':
            _same(to_host(direct.sample(0, idxs=op[1])), to_host(ahead.sample(0, idxs=op[1])), (step, 'idxs'))
        elif op[0] == 'many':
            _same(to_host(direct.sample_many(op[1], op[2])), to_host(ahead.sample_many(op[1], op[2])), (step, 'many'))
        else:
            assert np.array_equal(direct.sample_goals(op[1], 0.2, 0.5, 0.3, True), ahead.sample_goals(op[1], 0.2, 0.5, 0.3, True))
    assert direct.state_dict() == ahead.state_dict()
    assert n > 0


def test_lookahead_checkpoint_resumes_mid_block():
    fields = toy_fields(5, ragged(5, 50, 4, 60), (6,), 2, np.float32)
    config = cfg()
    a = device_sampler(fields, config, 'gc', seed=4, lookahead=8)
    for _ in range(3):
        a.sample(32)
    state = a.state_dict()
    assert state == {'counter': 3}                      # three batches handed out, although eight were drawn
    want = [to_host(a.sample(32)) for _ in range(7)]    # crosses into the next block
    b = device_sampler(fields, config, 'gc', seed=4, lookahead=8)
    b.load_state_dict(state)
    for i, w in enumerate(want):
        _same(to_host(b.sample(32)), w, i)
    a.load_state_dict(state)                            # ... and rewinding the same sampler drops what it drew ahead
    _same(to_host(a.sample(32)), want[0], 'rewound')


def test_lookahead_slices_hand_off_through_dlpack():
    """Every popped batch is a [B, ...] tensor of its own for a DLPack consumer, alive after the sampler has moved on."""
    import torch

    fields = toy_fields(9, ragged(9, 60, 5, 70), (29,), 8, np.float32)
    config = cfg()
    direct = device_sampler(fields, config, 'gc', seed=6)
    ahead = device_sampler(fields, config, 'gc', seed=6, lookahead=4)
    kept = []
    for step in range(10):                               # 2.5 blocks: batches of recycled blocks must stay intact
        batch = ahead.sample(128)
        tensors = {k: torch.from_dlpack(v) for k, v in batch.items()}
        assert tensors['observations'].shape == (128, 29) and tensors['masks'].shape == (128,)
        assert tensors['masks'].dtype == torch.float64 and tensors['actions'].dtype == torch.float32
        kept.append((tensors, to_host(direct.sample(128))))
        del batch
    torch.cuda.synchronize()
    for step, (tensors, want) in enumerate(kept):
        for k, w in want.items():
            assert np.array_equal(tensors[k].cpu().numpy(), w), (step, k)


@pytest.mark.parametrize('kind,over', [('gc', {}), ('hgc', dict(subgoal_steps=5, gc_negative=False)), ('hgc', dict(subgoal_steps=3, low_discount=0.9)),
                                       ('gc', dict(agent_name='trl', value_p_curgoal=0.0, value_p_trajgoal=1.0, value_p_randomgoal=0.0))])
def test_jax_compat_scalars_are_the_narrowed_reference_scalars(kind, over):
    import torch

    fields = toy_fields(13, ragged(13, 70, 6, 50), (7,), 3, np.float32)
    config = cfg(**over)
    wide = device_sampler(fields, config, kind, seed=8)
    narrow = device_sampler(fields, config, kind, seed=8, jax_compat=True)
    for B, K in ((200, 1), (64, 3)):
        w, x = wide.sample_many(K, B), narrow.sample_many(K, B)
        assert set(w) == set(x)
        n_scalar = 0
        for k in w:
            a, b = np.asarray(w[k]), np.asarray(x[k])
            if a.dtype == np.float64:
                assert b.dtype == np.float32 and np.array_equal(b, a.astype(np.float32)), k
                assert torch.from_dlpack(x[k]).dtype == torch.float32
                n_scalar += 1
            elif a.dtype == np.int64:
                assert b.dtype == np.int32 and np.array_equal(b, a.astype(np.int32)), k
                assert torch.from_dlpack(x[k]).dtype == torch.int32
                n_scalar += 1
            else:
                assert b.dtype == a.dtype and np.array_equal(a, b), k
        assert n_scalar >= 2
    # host output and the look-ahead keep the narrow dtypes
    host = device_sampler(fields, config, kind, seed=8, jax_compat=True, output='numpy', lookahead=3)
    batch = host.sample(50)
    assert batch['masks'].dtype == np.float32 and batch['rewards'].dtype == np.float32


@pytest.mark.parametrize('output', ['device', 'numpy'])
def test_prefetcher_pipelines_given_idxs_and_multi_batch_items(output):
    """Prefetcher over sample_async: items are launched one ahead of the one being handed over; with an `idxs` iterable
    and num_batches every item is sample_many(K, B, idxs=...) of the next index array, and the iterator ends with it."""
    from ogbench_b200 import Prefetcher

    fields = toy_fields(41, ragged(41, 90, 4, 50), (13,), 4, np.float32)
    config = cfg()
    valid = np.nonzero(fields['valids'] > 0)[0]
    rng = np.random.default_rng(5)
    K, B = 3, 40
    index_arrays = [rng.choice(valid, K * B) for _ in range(7)]
    direct = device_sampler(fields, config, 'gc', seed=12, output=output)
    ahead = device_sampler(fields, config, 'gc', seed=12, output=output)
    got = []
    with Prefetcher(ahead, B, depth=2, num_batches=K, idxs=iter(index_arrays)) as batches:
        for item in batches:
            got.append(to_host(item))
    assert len(got) == len(index_arrays)
    for step, idxs in enumerate(index_arrays):
        _same(got[step], to_host(direct.sample_many(K, B, idxs=idxs)), step)
    # a pending batch is a plain two-phase call as well
    a, b = ahead.sample_async(B), ahead.sample_async(B, evaluation=True)
    _same(to_host(a.result()), to_host(direct.sample(B)), 'async 0')
    _same(to_host(b.result()), to_host(direct.sample(B, evaluation=True)), 'async 1')
    assert a.result() is a.result()


def test_begun_host_copies_queue_and_report():
    """output='numpy' + sample_async: the device-to-host copy is begun at launch time (ogb_batch_copy_to_host_begin) and
    ended by result().  Several copies begun before any is ended return the batches direct calls return, in any result()
    order; a pending batch dropped without result() leaves the sampler (and the pinned pool) usable; an index outside the
    table surfaces as IndexError from result() (deferred check), and the sampler keeps working after it."""
    import gc

    fields = toy_fields(77, ragged(77, 120, 4, 60), (21,), 5, np.float32)
    config = cfg()
    valid = np.nonzero(fields['valids'] > 0)[0]
    direct = device_sampler(fields, config, 'gc', seed=3, output='numpy')
    ahead = device_sampler(fields, config, 'gc', seed=3, output='numpy')
    pend = [ahead.sample_async(96) for _ in range(5)]
    want = [direct.sample(96) for _ in range(5)]
    for i in (3, 0, 4, 1, 2):
        _same(pend[i].result(), want[i], i)
    # dropped without result(): the copy is waited for before its pinned block is recycled
    dropped = ahead.sample_async(96, num_batches=40)
    del dropped
    gc.collect()
    direct.sample_many(40, 96)
    _same(ahead.sample(50), direct.sample(50), 'after drop')
    # deferred index check
    n = len(fields['terminals'])
    bad = np.array([valid[0], n + 5, valid[1]], dtype=np.int64)
    p = ahead.sample_async(3, idxs=bad)
    with pytest.raises(IndexError):
        p.result()
    direct.sample(3, idxs=valid[:3].astype(np.int64))      # (the failed launch advanced the Philox counter by one call)
    good = valid[:7].astype(np.int64)
    _same(ahead.sample_async(7, idxs=good).result(), direct.sample(7, idxs=good), 'after error')
