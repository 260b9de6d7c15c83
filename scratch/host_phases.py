"""Where the host time of a one-batch sample() goes (OGB_HOST_PHASES=1 prints the library's per-phase means at exit)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ogbench_b200 import Dataset, GCDataset, HGCDataset, synthetic
key = sys.argv[1] if len(sys.argv) > 1 else 'c2'
w = synthetic.WORKLOADS[key]
ds = Dataset.create(**synthetic.device_fields(w))
s = (GCDataset if w.kind == 'gc' else HGCDataset)(ds, w.config)
for _ in range(50): b = s._sampler.sample_native(w.batch)
torch.cuda.synchronize(); t0 = time.perf_counter()
N = 2000
for _ in range(N): b = s._sampler.sample_native(w.batch)
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f'{key}: native call {1e6*(t1-t0)/N:.2f} us host, {1e6*(t2-t0)/N:.2f} us per call incl. drain')
t0 = time.perf_counter()
for _ in range(N): b = s.sample(w.batch)
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f'{key}: public call {1e6*(t1-t0)/N:.2f} us host, {1e6*(t2-t0)/N:.2f} us per call incl. drain')
