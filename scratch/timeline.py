import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ogbench_b200 import Dataset, GCDataset, HGCDataset, synthetic, _native
key = sys.argv[1] if len(sys.argv) > 1 else 'c2'
w = synthetic.WORKLOADS[key]
ds = Dataset.create(**synthetic.device_fields(w))
cls = GCDataset if w.kind == 'gc' else HGCDataset
s = cls(ds, w.config)
L = (1 << 20) // w.batch
hs = []
for _ in range(12):
    hs.append(s._sampler.sample_native(w.batch, n_batches=L)); hs = hs[-1:]
out = (C.c_double * 4096)(); n = C.c_int32()
_native.check(_native.lib().ogb_debug_timeline(out, 4096, C.byref(n)))
t = np.array(out[:n.value]).reshape(-1, 4)
print('call: index_begin index_end gather_begin gather_end (ms)')
for i, r in enumerate(t): print(i, ' '.join(f'{x:8.3f}' for x in r))
