"""single-call latency of sample(B) (device output, no sync between calls) and of a synced call"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ogbench_b200 import Dataset, GCDataset, HGCDataset, synthetic
for key in sys.argv[1:] or ['c2']:
    w = synthetic.WORKLOADS[key]
    ds = Dataset.create(**synthetic.device_fields(w))
    cls = GCDataset if w.kind == 'gc' else HGCDataset
    s = cls(ds, w.config)
    for K in (1, 8, 64):
        for _ in range(20): b = s._sampler.sample_native(w.batch, n_batches=K)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        N = 300
        for _ in range(N): b = s._sampler.sample_native(w.batch, n_batches=K)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / N
        # synced: includes launch latency + kernel + sync
        t0 = time.perf_counter()
        for _ in range(100):
            b = s._sampler.sample_native(w.batch, n_batches=K); torch.cuda.synchronize()
        ds_ = (time.perf_counter() - t0) / 100
        print(f'{key} K={K}: back-to-back {dt*1e6:.1f} us/call, synced {ds_*1e6:.1f} us/call ({w.batch*K/dt/1e6:.1f} M tr/s)')
