import sys, os, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ogbench_b200 import Dataset, GCDataset, synthetic, _native
import ogbench_b200.datasets as D
key = 'c2'
w = synthetic.WORKLOADS[key]
ds = Dataset.create(**synthetic.device_fields(w))
s = GCDataset(ds, w.config, output='numpy')
lib = _native.lib()
Le = 127; rows = Le * w.batch
pinned = C.c_void_p(); _native.check(lib.ogb_host_alloc(rows * 8, C.byref(pinned)))
pool = np.frombuffer((C.c_ubyte * (rows * 8)).from_address(pinned.value), dtype=np.int64)
pos = np.random.default_rng(0).integers(0, 1000000, size=rows); pool[:] = pos + pos // 1000
for _ in range(3): out = s.sample_many(Le, w.batch, idxs=pool); del out
T = {}
def tick(name, t0): T[name] = T.get(name, 0.0) + time.perf_counter() - t0
N = 30
t_all = time.perf_counter()
for _ in range(N):
    t0 = time.perf_counter(); h = s._sampler.sample_native(w.batch, n_batches=Le, idxs=pool); tick('sample_native', t0)
    t0 = time.perf_counter(); _native.check(lib.ogb_batch_sync(h.ptr)); tick('kernel_sync', t0)
    nbytes = C.c_size_t(); lib.ogb_batch_nbytes(h.ptr, C.byref(nbytes))
    t0 = time.perf_counter(); block = D._PINNED.take(nbytes.value); tick('pinned_take', t0)
    t0 = time.perf_counter(); _native.check(lib.ogb_batch_copy_to_host(h.ptr, C.c_void_p(block.ptr), block.bucket)); tick('d2h', t0)
    t0 = time.perf_counter(); del h, block; tick('release', t0)
print('per-step ms:', {k: round(1e3 * v / N, 3) for k, v in T.items()}, 'total', round(1e3 * (time.perf_counter() - t_all) / N, 3), 'bytes', nbytes.value)
t0 = time.perf_counter()
for _ in range(N): out = s.sample_many(Le, w.batch, idxs=pool); del out
print('public API per-step ms', round(1e3 * (time.perf_counter() - t0) / N, 3))
t0 = time.perf_counter()
for _ in range(N):
    h = s._sampler.sample_native(w.batch, n_batches=Le, idxs=pool); d = s._sampler.wrap(h); del h, d
print('native+wrap per-step ms', round(1e3 * (time.perf_counter() - t0) / N, 3))
