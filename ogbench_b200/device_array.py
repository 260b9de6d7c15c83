"""Device-resident batch arrays and their hand-off protocols (DLPack, __cuda_array_interface__).

``sample()`` returns a dict of ``DeviceArray``.  A JAX consumer does ``jax.dlpack.from_dlpack(x)`` (or
``jnp.asarray``), a torch consumer ``torch.from_dlpack(x)``; ``np.asarray(x)`` / ``x.numpy()`` copies to host.
All arrays of one batch are views into one device block owned by the native ``ogb_batch``; the block is returned
to the stream-ordered pool when the last view (and every DLPack consumer) has dropped it.
"""

from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native

_PyCapsule_New = C.pythonapi.PyCapsule_New
_PyCapsule_New.restype = C.py_object
_PyCapsule_New.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
_PyCapsule_IsValid = C.pythonapi.PyCapsule_IsValid
_PyCapsule_IsValid.restype = C.c_int
_PyCapsule_IsValid.argtypes = [C.py_object, C.c_char_p]
_PyCapsule_GetPointer = C.pythonapi.PyCapsule_GetPointer
_PyCapsule_GetPointer.restype = C.c_void_p
_PyCapsule_GetPointer.argtypes = [C.py_object, C.c_char_p]


class _DLManagedTensorHead(C.Structure):
    # only the tail of DLManagedTensor matters here: we call ->deleter if the capsule was never consumed
    _fields_ = [('data', C.c_void_p), ('device_type', C.c_int32), ('device_id', C.c_int32), ('ndim', C.c_int32),
                ('code', C.c_uint8), ('bits', C.c_uint8), ('lanes', C.c_uint16), ('shape', C.c_void_p),
                ('strides', C.c_void_p), ('byte_offset', C.c_uint64), ('manager_ctx', C.c_void_p),
                ('deleter', C.CFUNCTYPE(None, C.c_void_p))]


@C.CFUNCTYPE(None, C.c_void_p)
def _capsule_destructor(capsule_ptr):
    # a consumer renames the capsule to "used_dltensor" and takes ownership; otherwise we must call the deleter
    capsule = C.cast(capsule_ptr, C.py_object)
    if _PyCapsule_IsValid(capsule, b'dltensor'):
        ptr = _PyCapsule_GetPointer(capsule, b'dltensor')
        head = _DLManagedTensorHead.from_address(ptr)
        if head.deleter:
            head.deleter(ptr)


class BatchHandle:
    """Owns one reference on a native ogb_batch."""

    __slots__ = ('ptr', 'device', 'stream', '__weakref__')

    def __init__(self, ptr, device, stream):
        self.ptr = ptr
        self.device = device
        self.stream = stream

    def __del__(self):
        ptr, self.ptr = self.ptr, None
        if ptr:
            try:
                _native.lib().ogb_batch_release(ptr)
            except Exception:  # interpreter shutdown
                pass

    def sync(self):
        _native.check(_native.lib().ogb_batch_sync(self.ptr))


class DeviceArray:
    """One key of a batch, resident in HBM."""

    __slots__ = ('_batch', '_index', 'shape', 'dtype', 'ptr', 'nbytes', 'name', '_slice')

    def __init__(self, batch: BatchHandle, index: int, name: str, dtype, shape, ptr: int, nbytes: int, batch_index: int = -1):
        self._batch = batch
        self._index = index
        self.shape = shape
        self.dtype = dtype
        self.ptr = ptr
        self.nbytes = nbytes
        self.name = name
        self._slice = batch_index      # >= 0: this array is ONE batch of a multi-batch launch (look-ahead sampling)

    # ---- array-ish surface ----
    @property
    def ndim(self):
        return len(self.shape)

    @property
    def size(self):
        return int(np.prod(self.shape, dtype=np.int64))

    def __len__(self):
        return self.shape[0]

    def __repr__(self):
        return f'DeviceArray({self.name!r}, shape={self.shape}, dtype={self.dtype}, device=cuda:{self._batch.device})'

    # ---- hand-off protocols ----
    def __dlpack_device__(self):
        return (2, self._batch.device)  # kDLCUDA

    def __dlpack__(self, stream=None, **_unused):
        lib = _native.lib()
        if stream != -1:  # -1: the consumer asks for no synchronisation
            handle = {None: 1, 1: 1, 2: 2}.get(stream, stream)  # legacy default / per-thread default / explicit handle
            _native.check(lib.ogb_batch_wait_on_stream(self._batch.ptr, C.c_void_p(handle)))
        out = C.c_void_p()
        if self._slice >= 0:
            _native.check(lib.ogb_batch_dlpack_slice(self._batch.ptr, self._index, self._slice, C.byref(out)))
        else:
            _native.check(lib.ogb_batch_dlpack(self._batch.ptr, self._index, C.byref(out)))
        return _PyCapsule_New(out, b'dltensor', C.cast(_capsule_destructor, C.c_void_p))

    @property
    def __cuda_array_interface__(self):
        self._batch.sync()  # the interface carries no stream: hand over only finished data ...
        _native.check(_native.lib().ogb_batch_mark_escaped(self._batch.ptr))  # ... and recycle the block conservatively
        return {'shape': self.shape, 'typestr': self.dtype.str, 'data': (self.ptr, False), 'version': 3, 'strides': None}

    def numpy(self) -> np.ndarray:
        """Synchronous device->host copy of this key."""
        out = np.empty(self.shape, dtype=self.dtype)
        if self._slice >= 0:
            _native.check(_native.lib().ogb_batch_copy_slice_to_host(self._batch.ptr, self._index, self._slice,
                                                                     out.ctypes.data_as(C.c_void_p), out.nbytes))
        else:
            _native.check(_native.lib().ogb_batch_copy_key_to_host(self._batch.ptr, self._index, out.ctypes.data_as(C.c_void_p),
                                                                   out.nbytes))
        return out

    def __array__(self, dtype=None, copy=None):
        arr = self.numpy()
        return arr if dtype is None else arr.astype(dtype)

    def torch(self):
        import torch

        return torch.from_dlpack(self)
