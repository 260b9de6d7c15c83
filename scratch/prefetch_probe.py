import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ogbench_b200 import Dataset, GCDataset, Prefetcher, synthetic
w = synthetic.WORKLOADS['c2']
ds = Dataset.create(**synthetic.device_fields(w))
s = GCDataset(ds, w.config)
for depth in (1, 2, 4):
    with Prefetcher(s, w.batch, depth=depth) as batches:
        b = next(batches); t_del = t_next = 0.0; N = 300
        for _ in range(N):
            time.sleep(300e-6)
            t1 = time.perf_counter(); b = None; t2 = time.perf_counter(); b = next(batches); t3 = time.perf_counter()
            t_del += t2 - t1; t_next += t3 - t2
        print(f'depth {depth}: del {t_del/N*1e6:.1f} us, next {t_next/N*1e6:.1f} us, qsize {batches._queue.qsize()}')
