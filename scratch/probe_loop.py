import sys, time, os, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ogbench_b200 import Dataset, GCDataset, HGCDataset, synthetic
sys.path.insert(0, '.')
from bench import ClockSampler
key = sys.argv[1] if len(sys.argv) > 1 else 'c2'
w = synthetic.WORKLOADS[key]
ds = Dataset.create(**synthetic.device_fields(w))
cls = GCDataset if w.kind == 'gc' else HGCDataset
s = cls(ds, w.config)
st = torch.cuda.Stream()
s._sampler.set_stream(st.cuda_stream)
L = 1024
def run(steps, nvml, per_step_events, sync_each):
    for _ in range(3):
        h = s._sampler.sample_native(w.batch, n_batches=L)
    torch.cuda.synchronize()
    cs = None
    if nvml:
        cs = ClockSampler(0); cs.start()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    evs = []
    e0.record(st)
    t0 = time.perf_counter()
    for _ in range(steps):
        if per_step_events:
            a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True); a.record(st)
        h = s._sampler.sample_native(w.batch, n_batches=L)
        if per_step_events:
            b.record(st); evs.append((a, b))
        if sync_each:
            torch.cuda.synchronize()
    e1.record(st)
    t_host = time.perf_counter() - t0
    torch.cuda.synchronize()
    if cs:
        cs.stop_flag.set(); cs.join()
    tot = e0.elapsed_time(e1) / steps
    per = np.mean([a.elapsed_time(b) for a, b in evs]) if evs else float('nan')
    print(f'{key} steps={steps} nvml={nvml} events={per_step_events} sync={sync_each}: ms/step {tot:.3f} per-step-events {per:.3f} host_ms/step {1e3*t_host/steps:.3f}', cs.summary() if cs else '')
run(30, False, False, False)
run(30, False, True, False)
run(30, True, True, False)
run(30, False, True, True)
run(300, False, False, False)
run(300, True, False, False)
