"""Long-running stability check: many launches, mixed call shapes, device memory must not grow."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ogbench_b200 import Dataset, GCDataset, HGCDataset, synthetic

def free_mb():
    torch.cuda.synchronize()
    return torch.cuda.mem_get_info()[0] / 1e6

for key, n_iter in (('c2', 6000), ('c3', 1500), ('c4', 1500)):
    w = synthetic.WORKLOADS[key]
    ds = Dataset.create(**synthetic.device_fields(w))
    cls = GCDataset if w.kind == 'gc' else HGCDataset
    dev = cls(ds, w.config, seed=1)
    host = cls(ds, w.config, seed=2, output='numpy')
    shapes = [(1, w.batch), (8, w.batch), (64 if key != 'c4' else 4, w.batch), (1, 37), (3, 1000)]
    for K, B in shapes:                        # warm every block size class
        dev.sample_many(K, B); host.sample_many(min(K, 4), B)
    before = free_mb()
    t0 = time.perf_counter()
    rng = np.random.default_rng(0)
    for it in range(n_iter):
        K, B = shapes[it % len(shapes)]
        out = dev.sample_many(K, B)
        if it % 50 == 0:
            h = host.sample_many(min(K, 4), B)
            assert h['observations'].shape[:2] == (min(K, 4), B)
        if it % 97 == 0:
            t = torch.from_dlpack(out['rewards'])
            assert torch.isfinite(t).all()
        del out
    after = free_mb()
    print(f'{key}: {n_iter} calls in {time.perf_counter() - t0:.1f} s, free device memory {before:.0f} -> {after:.0f} MB, counter {dev.state_dict()}')
    assert before - after < 64, 'device memory grew'
    del dev, host, ds
    torch.cuda.empty_cache()
print('soak ok')
