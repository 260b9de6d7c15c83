import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ogbench_b200 import Dataset, GCDataset, HGCDataset, synthetic
key = sys.argv[1] if len(sys.argv) > 1 else 'c2'
w = synthetic.WORKLOADS[key]
ds = Dataset.create(**synthetic.device_fields(w))
cls = GCDataset if w.kind == 'gc' else HGCDataset
s = cls(ds, w.config)
st = torch.cuda.Stream()
s._sampler.set_stream(st.cuda_stream)
for L in (1, 64, 1024):
    for keep in (1, 2):
        hs = []
        for _ in range(4):
            hs.append(s._sampler.sample_native(w.batch, n_batches=L)); hs = hs[-keep:]
        torch.cuda.synchronize()
        evs = []; host = []
        for _ in range(10):
            a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
            a.record(st); t0 = time.perf_counter()
            hs.append(s._sampler.sample_native(w.batch, n_batches=L)); 
            t1 = time.perf_counter(); b.record(st)
            hs = hs[-keep:]
            host.append(t1 - t0); evs.append((a, b))
            torch.cuda.synchronize()
        print(key, 'L', L, 'keep', keep, 'host_us %.1f' % (1e6 * np.median(host)), 'gpu_ms %.3f' % np.median([a.elapsed_time(b) for a, b in evs]))
