"""TEST INFRASTRUCTURE ONLY -- import the *unmodified* reference sampler file under stub modules.

The reference hot path lives in ``<reference>/impls/utils/datasets.py``.  That file imports ``jax``,
``jax.numpy`` and ``flax.core.frozen_dict`` (datasets.py:5-8), none of which is installed in this image.
Everything the sampler does with them on the GC/HGC path is: ``jax.tree_util.tree_map`` / ``tree_leaves``
over plain dicts, ``FrozenDict`` as an immutable mapping, and one jitted crop (datasets.py:17-33).  This
module registers minimal stand-ins for those three things in ``sys.modules`` *before* executing the reference
file.  The crop runs AS WRITTEN in the reference (``random_crop`` = ``jnp.pad(mode='edge')`` + ``lax.dynamic_slice``,
vmapped over the batch): ``jnp.pad`` is ``np.pad`` (same ``mode='edge'`` semantics), ``lax.dynamic_slice`` is a numpy
slice with XLA's documented start clamping (start = clip(start, 0, dim - size)), ``jax.vmap`` is a Python loop over the
mapped axis and ``jax.jit`` is the identity -- so the golden vectors' crops come from the reference's own function
bodies, and the oracle's closed form (``shifted_edge_crop``) is checked against them, not against itself.
``load_reference_datasets_module(fast_crop=True)`` swaps in the closed form instead (a Python loop per image is slow
for big batches); the golden generator and the crop tests use the literal path.

Used by ``tests/golden/make_golden.py`` (fixture generation, in the build container only) and by the
live cross-check tests, which skip when the reference tree is not mounted.  Nothing in the product package
imports this file, and nothing here is read at run time on the GPU box (``/root/reference`` does not exist
there).
"""

from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get('OGB_REFERENCE_ROOT', '/root/reference')
_REF_FILE = os.path.join(REFERENCE_ROOT, 'impls', 'utils', 'datasets.py')


def reference_available() -> bool:
    return os.path.isfile(_REF_FILE)


def _tree_map(fn, tree, *rest):
    """dict-only stand-in for jax.tree_util.tree_map (sorted keys, like jax's dict flattening)."""
    if isinstance(tree, dict):
        return {k: _tree_map(fn, tree[k], *[r[k] for r in rest]) for k in sorted(tree)}
    if hasattr(tree, '_dict') and isinstance(tree._dict, dict):
        return _tree_map(fn, tree._dict, *rest)
    return fn(tree, *rest)


def _tree_leaves(tree):
    if isinstance(tree, dict):
        out = []
        for k in sorted(tree):
            out.extend(_tree_leaves(tree[k]))
        return out
    if hasattr(tree, '_dict') and isinstance(tree._dict, dict):
        return _tree_leaves(tree._dict)
    return [tree]


class _FrozenDict:
    """The slice of flax.core.FrozenDict the reference sampler touches."""

    def __init__(self, *args, **kwargs):
        src = dict(*args, **kwargs) if not (len(args) == 1 and isinstance(args[0], _FrozenDict)) else dict(args[0]._dict)
        self._dict = src

    def __getitem__(self, key):
        return self._dict[key]

    def __contains__(self, key):
        return key in self._dict

    def __iter__(self):
        return iter(self._dict)

    def __len__(self):
        return len(self._dict)

    def keys(self):
        return self._dict.keys()

    def values(self):
        return self._dict.values()

    def items(self):
        return self._dict.items()

    def copy(self, add_or_replace=None):
        merged = dict(self._dict)
        merged.update(add_or_replace or {})
        return type(self)(merged)


def _identity_decorator(fn=None, **_kwargs):
    if fn is None:
        return lambda f: f
    return fn


def _vmap(fn, in_axes=0, out_axes=0):
    """jax.vmap for positional array arguments: apply `fn` per index of the mapped axes, stack the results."""
    assert out_axes == 0

    def mapped(*args):
        axes = tuple(in_axes) if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        assert len(axes) == len(args)
        n = next(np.shape(a)[ax] for a, ax in zip(args, axes) if ax is not None)
        outs = [fn(*[np.take(np.asarray(a), i, axis=ax) if ax is not None else a for a, ax in zip(args, axes)]) for i in range(n)]
        return np.stack(outs) if outs else np.zeros((0,) + tuple(np.shape(args[0])[1:]), dtype=np.asarray(args[0]).dtype)

    return mapped


def _dynamic_slice(operand, start_indices, slice_sizes):
    """jax.lax.dynamic_slice: start indices are clamped so that the slice stays inside the operand (XLA DynamicSlice)."""
    operand = np.asarray(operand)
    assert len(start_indices) == operand.ndim == len(slice_sizes)
    starts = [int(np.clip(int(s), 0, d - z)) for s, d, z in zip(start_indices, operand.shape, slice_sizes)]
    return operand[tuple(slice(s, s + z) for s, z in zip(starts, slice_sizes))]


def _install_stubs():
    if 'jax' in sys.modules and not getattr(sys.modules['jax'], '_ogb_stub', False):
        raise RuntimeError('a real jax is importable; refshim is only meant for images without it')
    jax = types.ModuleType('jax')
    jax._ogb_stub = True
    jax.tree_util = types.ModuleType('jax.tree_util')
    jax.tree_util.tree_map = _tree_map
    jax.tree_util.tree_leaves = _tree_leaves
    jax.jit = _identity_decorator
    jax.vmap = _vmap
    jax.lax = types.ModuleType('jax.lax')
    jax.lax.dynamic_slice = _dynamic_slice
    jnp = types.ModuleType('jax.numpy')
    jnp.pad = np.pad
    jax.numpy = jnp
    flax = types.ModuleType('flax')
    flax_core = types.ModuleType('flax.core')
    flax_fd = types.ModuleType('flax.core.frozen_dict')
    flax_fd.FrozenDict = _FrozenDict
    flax.core = flax_core
    flax_core.frozen_dict = flax_fd
    flax_core.FrozenDict = _FrozenDict
    for name, mod in [
        ('jax', jax),
        ('jax.tree_util', jax.tree_util),
        ('jax.lax', jax.lax),
        ('jax.numpy', jnp),
        ('flax', flax),
        ('flax.core', flax_core),
        ('flax.core.frozen_dict', flax_fd),
    ]:
        sys.modules.setdefault(name, mod)


_CACHED = None


def load_reference_datasets_module(fast_crop: bool = False):
    """Execute the reference's datasets.py (unmodified, read from its own location) and return the module.

    fast_crop=False: `random_crop` / `batched_random_crop` run as the reference wrote them (datasets.py:17-33) on the
    numpy stand-ins above.  fast_crop=True: `batched_random_crop` is the oracle's closed form (same results --
    tests/test_oracle_golden.py checks the two against each other -- without the per-image Python loop)."""
    global _CACHED
    if _CACHED is None:
        if not reference_available():
            raise FileNotFoundError(f'reference sampler not found at {_REF_FILE}')
        _install_stubs()
        spec = importlib.util.spec_from_file_location('ogb_reference_datasets', _REF_FILE)
        mod = importlib.util.module_from_spec(spec)
        sys.modules['ogb_reference_datasets'] = mod
        spec.loader.exec_module(mod)
        mod._literal_batched_random_crop = mod.batched_random_crop
        _CACHED = mod
    mod = _CACHED
    if fast_crop:
        from oracle.replay_oracle import shifted_edge_crop

        def batched_random_crop(imgs, crop_froms, padding):
            return shifted_edge_crop(np.asarray(imgs), np.asarray(crop_froms)[:, :2], padding)

        mod.batched_random_crop = batched_random_crop
    else:
        mod.batched_random_crop = mod._literal_batched_random_crop
    return mod


class DrawRecorder:
    """Context manager that logs every np.random.{randint,geometric,rand} call the reference makes.

    The log order is the validation-mode draw schema: one entry per numpy call,
    holding the already-transformed values.
    """

    def __init__(self):
        self.log = []
        self._saved = {}

    def __enter__(self):
        for name in ('randint', 'geometric', 'rand', 'choice'):
            self._saved[name] = getattr(np.random, name)

        def randint(*args, **kwargs):
            out = self._saved['randint'](*args, **kwargs)
            self.log.append(('randint', np.array(out, copy=True)))
            return out

        def geometric(*args, **kwargs):
            out = self._saved['geometric'](*args, **kwargs)
            self.log.append(('geometric', np.array(out, copy=True)))
            return out

        def rand(*args, **kwargs):
            out = self._saved['rand'](*args, **kwargs)
            self.log.append(('rand', np.array(out, copy=True)))
            return out

        def choice(a, size=None, **kwargs):
            # legacy RandomState.choice(a, size) without p is a[randint(0, len(a), size)]: log it as that randint
            assert not kwargs
            pos = self._saved['randint'](0, len(a), size=size)
            self.log.append(('randint', np.array(pos, copy=True)))
            return np.asarray(a)[pos]

        np.random.randint, np.random.geometric, np.random.rand, np.random.choice = randint, geometric, rand, choice
        return self

    def __exit__(self, *exc):
        for name, fn in self._saved.items():
            setattr(np.random, name, fn)
        return False
