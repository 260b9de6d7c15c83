import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ogbench_b200 import Dataset, GCDataset, HGCDataset, synthetic
key = sys.argv[1] if len(sys.argv) > 1 else 'c2'
w = synthetic.WORKLOADS[key]
ds = Dataset.create(**synthetic.device_fields(w))
cls = GCDataset if w.kind == 'gc' else HGCDataset
L = 1024 * 1024 // w.batch
def bench(n_streams, steps=60):
    ss = [cls(ds, w.config, stream_id=i) for i in range(n_streams)]
    sts = [torch.cuda.Stream() for _ in range(n_streams)]
    for s, st in zip(ss, sts): s._sampler.set_stream(st.cuda_stream)
    hs = [None] * n_streams
    for _ in range(3):
        for i, s in enumerate(ss): hs[i] = s._sampler.sample_native(w.batch, n_batches=L)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    t0 = time.perf_counter()
    for _ in range(steps):
        for i, s in enumerate(ss): hs[i] = s._sampler.sample_native(w.batch, n_batches=L)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    rows = steps * n_streams * L * w.batch
    print(f'{key} streams={n_streams}: {rows/dt:.4g} tr/s  ({1e3*dt/steps/n_streams:.3f} ms per launch-set)')
for n in (1, 2, 3):
    bench(n)
