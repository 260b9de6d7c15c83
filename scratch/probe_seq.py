import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ogbench_b200 import Dataset, synthetic
key = sys.argv[1] if len(sys.argv) > 1 else 'c2'
w = synthetic.WORKLOADS[key]
ds = Dataset.create(**synthetic.device_fields(w))
s = ds._plain_sampler(0)
n = 999936
rng = np.random.default_rng(0)
seq = np.arange(n, dtype=np.int64)
rnd = rng.integers(0, w.rows - 2, size=n).astype(np.int64)
def bench(idxs, label, steps=30):
    hs = []
    for _ in range(3):
        hs.append(s.sample_native(n, idxs=idxs)); hs = hs[-1:]
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    # H2D of idxs is inside; measure kernel part by events around a burst (H2D 8 MB ~ 0.15 ms each, so report both)
    t0 = time.perf_counter()
    for _ in range(steps):
        hs.append(s.sample_native(n, idxs=idxs)); hs = hs[-1:]
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    print(f'{key} {label}: {dt*1e3:.3f} ms/launch (includes 8 MB idx H2D)')
bench(seq, 'sequential')
bench(rnd, 'random')
