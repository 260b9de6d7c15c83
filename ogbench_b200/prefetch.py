"""Batches drawn ahead of the training loop.

The reference's loop draws its batch synchronously and then steps the agent (impls/main.py:202-203), so the sampler's
latency is on the critical path of every step.  Offline sampling does not depend on training state, so the next batch
can be drawn while the agent is stepping:

    batches = Prefetcher(train_dataset, batch_size=1024)        # any object with .sample(batch_size, **kwargs)
    for i in range(1, train_steps + 1):
        batch = next(batches)                                   # == train_dataset.sample(1024), already finished
        agent, info = agent.update(batch)

One worker thread owns the sampler (the library is called from one thread at a time); the C calls release the GIL,
so the launch, the device work and -- for output='numpy' -- the device-to-host copy of batch i+1 run under step i.
Measured on B200 with output='numpy' beside a consumer step of ~300 us that releases the GIL (as a jitted update
does; scratch/latency_public.py): C2 56 -> 13 us on the consumer's thread per batch of 1024, C4 724 -> 438 us per batch
of 256 stacked images (the rest is the PCIe copy that does not fit under the step).  With output='device' a direct
sample() is asynchronous already (17 us of host time) and needs no prefetching.

With rng='philox' the sequence of batches is exactly the sequence direct calls would return (the counter advances per
call).  rng='numpy' is refused: drawing ahead would interleave the global np.random stream differently from the
reference's call order, which that mode exists to reproduce.
"""

from __future__ import annotations

import atexit
import queue
import threading
import weakref
from typing import Any, Dict, Optional

_LIVE: 'weakref.WeakSet[Prefetcher]' = weakref.WeakSet()


@atexit.register
def _close_all():
    # a worker must not be inside the CUDA library while the interpreter (and the driver) shuts down
    for p in list(_LIVE):
        p.close()


class Prefetcher:
    """Iterator over `dataset.sample(batch_size, **sample_kwargs)`, `depth` batches ahead."""

    _STOP = object()

    def __init__(self, dataset, batch_size: int, depth: int = 2, idxs=None, **sample_kwargs):
        """`idxs`: optional iterable of index arrays, one per batch (the `idxs` argument of successive sample() calls); the
        iterator ends when it does.  `num_batches=K` in `sample_kwargs` draws sample_many(K, batch_size) per item."""
        if depth < 1:
            raise ValueError('depth must be at least 1')
        if getattr(dataset, 'rng', 'philox') == 'numpy':
            raise ValueError("rng='numpy' replays the reference's np.random call order and cannot draw ahead")
        self.dataset = dataset
        self.batch_size = int(batch_size)
        self.sample_kwargs = dict(sample_kwargs)
        self._idxs = iter(idxs) if idxs is not None else None
        self._queue: 'queue.Queue[Any]' = queue.Queue(maxsize=depth)
        self._closing = threading.Event()
        self._thread: Optional[threading.Thread] = threading.Thread(target=self._work, daemon=True)
        self.drawn = 0          # batches handed to the consumer
        _LIVE.add(self)
        self._thread.start()

    def _put(self, item):
        while not self._closing.is_set():
            try:
                self._queue.put(item, timeout=0.05)
                return
            except queue.Full:
                continue

    def _launch(self):
        """Start the next batch: a PendingBatch when the sampler can split launch and hand-over (sample_async), else the
        finished batch.  None when the `idxs` iterator is exhausted."""
        kwargs = self.sample_kwargs
        if self._idxs is not None:
            try:
                kwargs = dict(kwargs, idxs=next(self._idxs))
            except StopIteration:
                return None
        if hasattr(self.dataset, 'sample_async'):
            return self.dataset.sample_async(self.batch_size, **kwargs)
        if 'num_batches' in kwargs:
            kwargs = dict(kwargs)
            return self.dataset.sample_many(kwargs.pop('num_batches'), self.batch_size, **kwargs)
        return self.dataset.sample(self.batch_size, **kwargs)

    def _work(self):
        # Software-pipelined by one batch: batch k+1 is launched BEFORE batch k is finished, so with host output the
        # upload and the kernels of k+1 run under the device-to-host copy of k (which has a stream of its own).
        try:
            pending = self._launch()
            while pending is not None and not self._closing.is_set():
                failure = None
                try:
                    nxt = self._launch()
                except BaseException as exc:      # surfaces after the batch launched before it has been handed over
                    nxt, failure = None, exc
                self._put(pending.result() if hasattr(pending, 'result') else pending)
                if failure is not None:
                    raise failure
                pending = nxt
            if pending is None:
                self._put(self._STOP)
        except BaseException as exc:  # handed to the consumer at its next call
            self._queue.put(exc)

    def __iter__(self):
        return self

    def __next__(self) -> Dict[str, Any]:
        if self._thread is None:
            raise StopIteration
        item = self._queue.get()
        if item is self._STOP:
            self._thread = None
            raise StopIteration
        if isinstance(item, BaseException):
            self._thread = None
            raise item
        self.drawn += 1
        return item

    def close(self):
        """Stop the worker; batches already drawn are dropped (the sampler's counter has moved past them)."""
        thread, self._thread = self._thread, None
        if thread is None:
            return
        self._closing.set()
        while thread.is_alive():
            try:
                self._queue.get_nowait()
            except queue.Empty:
                pass
            thread.join(timeout=0.01)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
