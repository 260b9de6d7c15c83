import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ogbench_b200 import Dataset, GCDataset, HGCDataset, synthetic
key = sys.argv[1] if len(sys.argv) > 1 else 'c2'
w = synthetic.WORKLOADS[key]
ds = Dataset.create(**synthetic.device_fields(w))
cls = GCDataset if w.kind == 'gc' else HGCDataset
L = 1024 * 1024 // w.batch
def bench(own_stream, keep, steps=100):
    s = cls(ds, w.config)
    if not own_stream:
        st = torch.cuda.Stream(); s._sampler.set_stream(st.cuda_stream)
    hs = []
    for _ in range(5):
        hs.append(s._sampler.sample_native(w.batch, n_batches=L)); hs = hs[-keep:]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        hs.append(s._sampler.sample_native(w.batch, n_batches=L)); hs = hs[-keep:]
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f'{key} own_stream={own_stream} keep={keep} no_overlap={os.environ.get("OGB_NO_OVERLAP")}: {1e3*dt/steps:.3f} ms/step')
for own in (True, False):
    for keep in (1, 2, 3):
        bench(own, keep)
