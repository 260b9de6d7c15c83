"""latency of the PUBLIC call, one batch per call as impls/main.py:202 makes it: GCDataset.sample(B) -> dict"""
import sys, os, time, cProfile, pstats, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ogbench_b200 import Dataset, GCDataset, HGCDataset, synthetic
prof = '--profile' in sys.argv
for key in [a for a in sys.argv[1:] if not a.startswith('--')] or ['c2']:
    w = synthetic.WORKLOADS[key]
    ds = Dataset.create(**synthetic.device_fields(w))
    cls = GCDataset if w.kind == 'gc' else HGCDataset
    for output in ('device', 'numpy'):
        s = cls(ds, w.config, output=output)
        for _ in range(50): b = s.sample(w.batch)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        N = 500
        for _ in range(N): b = s.sample(w.batch)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / N
        print(f'{key} output={output}: {dt*1e6:.1f} us per sample({w.batch}) through the public API ({len(b)} keys)')
        for K in (64, 1024):                       # the same unchanged loop with batches drawn K at a time behind sample()
            la = cls(ds, w.config, output=output, lookahead=K)
            n_la = 4 * K if output == 'device' else K + 8
            for _ in range(K + 2): b = la.sample(w.batch)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            for _ in range(n_la): b = la.sample(w.batch)
            torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / n_la
            print(f'{key} output={output} lookahead={K}: {dt*1e6:.2f} us per sample({w.batch}) amortised over {n_la} calls')
            del la, b
        from ogbench_b200 import Prefetcher
        with Prefetcher(s, w.batch) as batches:
            next(batches); waited = 0.0
            for _ in range(N):
                t1 = time.perf_counter()
                time.sleep(300e-6)                                     # the consumer's own work (an agent step; releases the GIL like a jitted update does)
                t1 = time.perf_counter(); b = next(batches); waited += time.perf_counter() - t1
        print(f'{key} output={output}: {waited/N*1e6:.1f} us per next(Prefetcher) beside a ~300 us consumer step that releases the GIL')
        if prof:
            pr = cProfile.Profile(); pr.enable()
            for _ in range(N): b = s.sample(w.batch)
            pr.disable(); st = io.StringIO(); pstats.Stats(pr, stream=st).sort_stats('tottime').print_stats(14); print(st.getvalue()[:3500])
