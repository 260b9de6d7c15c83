#!/bin/bash
timeout 600 python scratch/fuzz.py 400 2024 2>&1 | tail -2
timeout 300 python scratch/soak.py 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 > gpurun_out/r2c_bench_default.json 2> gpurun_out/r2c_bench_default.err
tail -c 300 gpurun_out/r2c_bench_default.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c_bench_default.json').read().strip().splitlines()[-1])
for k,v in d['configs'].items():
    r=v['roofline']; print(k, 'L=%d'%v['batches_per_launch'], '%.4g tr/s'%v['value'], '%.4f ms'%v['ms_per_step'], 'kernel %.4f'%r['kernel_ms'], 'frac %.3f step %.3f'%(r['frac'], r['step_frac']), 'traffic', r['traffic'], r['traffic_source'][:40], [(s['batches_per_launch'], round(s['frac'],3)) for s in v['launch_size_sweep']], 'e2e %.4g link %.3g'%(v['e2e']['value'], v['e2e']['frac_of_link']), 'cpu %.3g'%v['cpu_baseline']['value'])
print(d['clocks'], d['timed_region_ms'], d['repeats'])
PY
export OGB_BENCH_NO_SWEEP=1
bash scratch/profile_all.sh r2c c1 | cut -c1-120
