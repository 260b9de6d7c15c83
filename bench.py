#!/usr/bin/env python
"""bench.py -- relabelled transitions/s of the replay sampler hot path on B200 (BASELINE.json's metric).

    python bench.py --gpus N --steps K --warmup W [--config c2] [--batches-per-launch L]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the CPU implementation of the same path on the host cores

A "step" is one launch of the hot path producing `batches_per_launch` successive GCDataset.sample(batch) calls
(config.batches_per_launch; 1 reproduces the reference's one-call-per-step usage and is launch-latency bound).
`value` = transitions all ranks produced / max-over-ranks device time with the dataset resident in HBM;
`e2e` = the same through the public Python API with host buffers (transition indices uploaded from pinned host
memory every step, the whole batch copied back into pinned host memory, both inside the timed region);
`roofline` = algorithmic bytes / average launch duration of the dominant kernel vs the measured
HBM copy bandwidth; `cpu_baseline` = the numpy oracle port timed on this box's host cores in the same run.
"""

from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = 'relabeled transitions/sec'
UNIT = 'transitions/s'
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md, used only when MEASURED_PEAKS.json is absent


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=300)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', choices=['ours', 'reference'], default='ours')
    ap.add_argument('--config', default='c2', help='workload key (c1..c5, c3b/c4b/c5b), SURVEY.md 8(d); c2 is the headline config')
    ap.add_argument('--batches-per-launch', type=int, default=None)
    ap.add_argument('--e2e-batches', type=int, default=None, help='batches per e2e step')
    ap.add_argument('--cpu-seconds', type=float, default=10.0, help='budget of the cpu_baseline leg')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    return ap.parse_args()


def hbm_peak():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        try:
            return float(json.load(open(path))['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
        except Exception:
            pass
    return FALLBACK_HBM_GBS, 'fallback (B200_PROFILING.md)'


# ------------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port (oracle/replay_oracle.py == the reference's numpy algorithm), timed on the host cores
# ------------------------------------------------------------------------------------------------------------------
def _cpu_fields(w):
    from ogbench_b200 import synthetic

    # the pixel workload's 12.3 GB dataset is not regenerated on the host: 100 episodes keep the per-call work
    # identical (same rows/frames per sample) while fitting in RAM
    episodes = min(w.episodes, 100) if w.obs_dtype == 'uint8' else w.episodes
    return synthetic.host_fields(w, episodes=episodes), episodes


def _cpu_worker(args):
    key, seconds, seed = args
    os.environ.setdefault('OMP_NUM_THREADS', '1')
    from ogbench_b200 import synthetic
    from oracle.replay_oracle import OracleSampler

    w = synthetic.WORKLOADS[key]
    fields, _ = _cpu_fields(w)
    sampler = OracleSampler(fields, w.config, w.kind)
    np.random.seed(seed)
    for _ in range(3):
        sampler.sample(w.batch)
    n, t0 = 0, time.perf_counter()
    while True:
        sampler.sample(w.batch)
        n += 1
        dt = time.perf_counter() - t0
        if dt >= seconds:
            return n, dt


def cpu_baseline(key, seconds, workers):
    """transitions/s of the oracle port with `workers` independent processes (1 = the reference's own threading)."""
    from ogbench_b200 import synthetic

    w = synthetic.WORKLOADS[key]
    if workers == 1:
        results = [_cpu_worker((key, seconds, 0))]
    else:
        import multiprocessing as mp

        with mp.get_context('spawn').Pool(workers) as pool:
            results = pool.map(_cpu_worker, [(key, seconds, i) for i in range(workers)])
    rate = sum(n * w.batch / dt for n, dt in results)
    calls = sum(n for n, _ in results)
    _, episodes = _cpu_fields(w)
    sample = (f'{calls} calls of sample({w.batch}) over {max(dt for _, dt in results):.1f} s on {workers} process(es), '
              f'numpy {np.__version__}, dataset {episodes}x{w.steps} rows')
    return rate, sample


def run_reference_arm(args):
    from ogbench_b200 import synthetic

    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    w = synthetic.WORKLOADS[args.config]
    workers = os.cpu_count() or 1
    # steps*warmup are honoured as a time budget: each "step" is a bounded slice of the same workload
    seconds = min(60.0, max(5.0, 0.05 * (args.steps + args.warmup)))
    t0 = time.perf_counter()
    rate, sample = cpu_baseline(args.config, seconds, workers)
    wall = time.perf_counter() - t0
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': rate, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': 1e3 * w.batch * workers / rate, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'u8' if w.obs_dtype == 'uint8' else 'f32', 'data': 'synthetic',
        'config': {'workload': w.name, 'batch': w.batch, 'rows': w.rows, 'note': 'reference numpy algorithm (oracle port; the '
                   'reference itself is Python and is not installed on the GPU box), one process per host core'},
        'cpu_baseline': {'value': rate, 'unit': UNIT, 'cores': workers, 'kind': 'port', 'sample': sample},
        'e2e': {'value': rate, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0, 'wall_s': wall,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons with NVML while the timed region runs."""

    def __init__(self, device_index, period=0.002):
        super().__init__(daemon=True)
        self.period = period
        self.samples = []
        self.reasons = set()
        self.stop_flag = threading.Event()
        self.ok = False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            visible = os.environ.get('CUDA_VISIBLE_DEVICES')
            phys = int(visible.split(',')[device_index]) if visible and visible.split(',')[device_index].isdigit() else device_index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.max_mhz = None

    def poll(self):
        nv = self.nv
        self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
        mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, 'nvmlDeviceGetCurrentClocksEventReasons') \
            else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        names = {
            'hw_slowdown': 0x8, 'sw_power_cap': 0x4, 'hw_thermal_slowdown': 0x40, 'sw_thermal_slowdown': 0x20,
            'hw_power_brake_slowdown': 0x80, 'sync_boost': 0x10, 'applications_clocks_setting': 0x2,
        }
        for name, bit in names.items():
            if mask & bit:
                self.reasons.add(name)

    def run(self):
        if not self.ok:
            return
        while not self.stop_flag.is_set():
            try:
                self.poll()
            except Exception:
                break
            time.sleep(self.period)

    def summary(self):
        if not self.ok:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvml unavailable']}
        if not self.samples:
            try:
                self.poll()
            except Exception:
                pass
        med = float(np.median(self.samples)) if self.samples else None
        return {'sm_mhz': med, 'sm_max_mhz': float(self.max_mhz), 'reasons': sorted(self.reasons), 'samples': len(self.samples)}


# ------------------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------------------
L2_BYTES = 126e6


def l2_note(out_bytes, resident_bytes):
    """How the timed loop relates to the 126 MB L2: outputs rotate through three blocks, inputs are random rows of the
    resident dataset."""
    out = (f'each step writes {out_bytes / 1e6:.0f} MB into one of three rotating output blocks '
           f'({"each larger than" if out_bytes > L2_BYTES else "together " + ("larger" if 3 * out_bytes > L2_BYTES else "smaller") + " than"} the 126 MB L2)')
    src = (f'and gathers random rows of a {resident_bytes / 1e6:.0f} MB resident dataset '
           f'({"larger than L2" if resident_bytes > L2_BYTES else "smaller than L2: it stays cache-resident, as the real dataset of this shape would"})')
    return out + ' ' + src


def default_batches_per_launch(w):
    # ~1M transitions per launch for vector workloads, 16 batches for the pixel workload: ~1.1 GB of output per launch
    # in both cases (>> 126 MB L2)
    return 16 if w.obs_dtype == 'uint8' else max(1, (1 << 20) // w.batch)


def run_gpu_arm(args):
    import ctypes as C

    import torch
    import torch.distributed as dist

    from ogbench_b200 import Dataset, GCDataset, HGCDataset, _native, dist_util, synthetic

    rank, world, local = dist_util.env_rank()
    if world > 1:   # NCCL prints its version / debug lines to stdout: keep stdout for the one JSON line
        os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')
        if os.environ.get('NCCL_DEBUG', '').upper() in ('', 'VERSION'):
            os.environ['NCCL_DEBUG'] = 'WARN'        # (an explicit INFO/TRACE request from the caller is left alone)
    torch.cuda.set_device(local)
    numa = dist_util.bind_to_gpu_numa(local) if world > 1 and not os.environ.get('OGB_NO_NUMA_BIND') else None
    dist_util.init('nccl', device=torch.device('cuda', local))
    w = synthetic.WORKLOADS[args.config]
    L = args.batches_per_launch or default_batches_per_launch(w)

    # each rank holds a replica (c1-c4) or its own trajectory-aligned shard (c5), generated directly in HBM
    fields = synthetic.device_fields(w, device=local, seed=w.seed + (rank if args.config.startswith('c5') else 0))
    dataset = Dataset.create(**fields)
    cls = GCDataset if w.kind == 'gc' else HGCDataset
    sampler = cls(dataset, w.config, device=local, seed=1234, stream_id=rank)
    del fields
    torch.cuda.empty_cache()
    stream = torch.cuda.Stream(device=local)
    sampler._sampler.set_stream(stream.cuda_stream)
    lib = _native.lib()
    _native.check(lib.ogb_sampler_set_profile(sampler._sampler.ptr, 1))   # CUDA events around the dominant kernel of each launch

    def barrier():
        dist_util.barrier()
        torch.cuda.synchronize(local)

    def launch():
        handle = sampler._sampler.sample_native(w.batch, n_batches=L)
        n = C.c_int32()
        lib.ogb_batch_launches(handle.ptr, C.byref(n))
        return handle, n.value

    # ---- device-resident throughput ----
    launches = 0
    dominant_ms = []
    dominant_name = C.c_char_p()

    def harvest(handle, keep):
        # device time of the launch's dominant kernel, from the events the library recorded on its own streams
        ms = C.c_float()
        _native.check(lib.ogb_batch_dominant_kernel(handle.ptr, C.byref(dominant_name), C.byref(ms)))
        if keep and ms.value >= 0:
            dominant_ms.append(ms.value)

    def run(n_steps, keep):
        # Up to three batches are alive at any time (the consumer holds two while the next is produced); a handle is
        # harvested two launches after its own, when its kernels have long finished, and dropping it recycles its block.
        nonlocal launches
        pending = []
        for _ in range(n_steps):
            h, n = launch()
            launches += n if keep else 0
            pending.append(h)
            if len(pending) > 2:
                harvest(pending.pop(0), keep)
        return pending

    tail = run(max(args.warmup, 3) + 2, False)   # same hand-over pattern as the timed loop: every recycled block exists
    del tail
    barrier()
    clocks = ClockSampler(local)
    clocks.start()
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        ev0.record(stream)
        tail = run(args.steps, True)
        ev1.record(stream)
    barrier()
    clocks.stop_flag.set()
    clocks.join()
    for h in tail:
        harvest(h, True)
    del tail
    elapsed_ms = ev0.elapsed_time(ev1)
    kernel_ms = float(np.mean(dominant_ms)) if dominant_ms else elapsed_ms / args.steps
    kernel_name = (dominant_name.value or b'').decode()
    elapsed_ms = dist_util.reduce_scalar(elapsed_ms, 'max', device=f'cuda:{local}')   # slowest rank
    per_step = w.batch * L
    value = dist_util.reduce_scalar(args.steps * per_step, 'sum', device=f'cuda:{local}') / (elapsed_ms * 1e-3)

    # ---- end to end through the public API with host buffers ----
    e2e = None
    if not args.no_e2e:
        Le = args.e2e_batches or max(1, min(L, (64 << 20) // (w.bytes_per_transition * w.batch // 2 + 1)))
        host_sampler = cls(dataset, w.config, device=local, seed=4321, stream_id=rank, output='numpy')
        rows = Le * w.batch
        n_valid = w.episodes * (w.steps - 1)
        pinned = C.c_void_p()
        _native.check(lib.ogb_host_alloc(rows * 8 * 4, C.byref(pinned)))
        pool = np.frombuffer((C.c_ubyte * (rows * 8 * 4)).from_address(pinned.value), dtype=np.int64).reshape(4, rows)
        rng = np.random.default_rng(rank)
        pos = rng.integers(0, n_valid, size=(4, rows))
        pool[:] = pos + pos // (w.steps - 1)  # valid_idxs[j] = j + j // (T-1) for fixed-length compact trajectories
        steps_e = max(3, min(args.steps, 50))
        d2h = 0
        for i in range(3):
            out = host_sampler.sample_many(Le, w.batch, idxs=pool[i % 4])
        d2h = sum(v.nbytes for k, v in out.items() if True)
        del out
        barrier()
        t0 = time.perf_counter()
        for i in range(steps_e):
            out = host_sampler.sample_many(Le, w.batch, idxs=pool[i % 4])  # returns after the D2H copy has completed
            del out
        torch.cuda.synchronize(local)
        dt = time.perf_counter() - t0
        dt = dist_util.reduce_scalar(dt, 'max', device=f'cuda:{local}')
        e2e = {'value': world * steps_e * rows / dt, 'unit': UNIT, 'numa_bound': numa is not None, 'h2d_bytes_per_step': rows * 8, 'd2h_bytes_per_step': int(d2h),
               'steps': steps_e, 'batches_per_step': Le, 'api': "GCDataset(..., output='numpy').sample_many(L, B, idxs=host)"}
        lib.ogb_host_free(pinned)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = hbm_peak()
    achieved = w.bytes_per_transition * per_step / (kernel_ms * 1e-3) / 1e9
    # DRAM bytes of one launch of the dominant kernel from the committed `ncu --set full` capture (profiles/traffic.json),
    # scaled to this run's launch size when the capture used another batches_per_launch
    traffic = None
    tpath = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(tpath):
        try:
            entry = json.load(open(tpath)).get(args.config)
            if entry and entry.get('kernel', '').find(kernel_name) >= 0:
                traffic = float(entry['dram_bytes_per_launch']) * per_step / float(entry.get('transitions_per_launch', per_step))
        except Exception:
            traffic = None
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
        'ms_per_step': elapsed_ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'u8' if w.obs_dtype == 'uint8' else 'f32', 'data': 'synthetic',
        'config': {
            'workload': w.name, 'key': w.key, 'rows_resident_per_gpu': w.rows, 'batch': w.batch, 'batches_per_launch': L,
            'transitions_per_step_per_gpu': per_step, 'rng': 'on-device Philox4x32-10',
            'placement': 'trajectory-aligned shard per GPU' if args.config.startswith('c5') else 'replica per GPU',
            'l2': l2_note(w.bytes_per_transition * per_step // 2, dataset.native(local).resident_bytes()),
        },
        'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak, 'traffic': traffic,
                     'kernel': kernel_name, 'kernel_ms': kernel_ms, 'bytes_per_transition': w.bytes_per_transition,
                     'bytes_per_launch': w.bytes_per_transition * per_step, 'peak_source': peak_src,
                     'step_frac': w.bytes_per_transition * per_step / (elapsed_ms / args.steps * 1e-3) / 1e9 / peak,
                     'note': 'achieved = algorithmic bytes of one launch / device time of the dominant kernel (CUDA events on its '
                             'stream); step_frac = the same bytes / whole step time (index kernel and launch gaps included)'},
        'clocks': clocks.summary(),
        'gpu_launches': launches,
    }
    if e2e is not None:
        line['e2e'] = e2e
    if world == 1 and not args.no_cpu_baseline:
        rate, sample = cpu_baseline(args.config, args.cpu_seconds, 1)
        line['cpu_baseline'] = {'value': rate, 'unit': UNIT, 'cores': 1, 'kind': 'port', 'sample': sample,
                                'host_cores_available': os.cpu_count()}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == 'reference':
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == '__main__':
    main()
