#!/bin/bash
# final artefacts of round 2: default bench line (all configs), reference arm, launch lists + ncu --set full per config, latency
python bench.py --steps 20 --warmup 5 > gpurun_out/r2c_bench_default.json 2> gpurun_out/r2c_bench_default.err
tail -c 300 gpurun_out/r2c_bench_default.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c_bench_default.json').read().strip().splitlines()[-1])
for k,v in d['configs'].items():
    r=v['roofline']; print(k, 'L=%d'%v['batches_per_launch'], '%.4g tr/s'%v['value'], '%.4f ms'%v['ms_per_step'], 'kernel %.4f'%r['kernel_ms'], 'frac %.3f step %.3f'%(r['frac'], r['step_frac']), 'traffic', r['traffic'], [(s['batches_per_launch'], round(s['frac'],3)) for s in v['launch_size_sweep']], 'e2e %.4g link %.3g'%(v['e2e']['value'], v['e2e']['frac_of_link']), 'cpu %.3g'%v['cpu_baseline']['value'])
print(d['clocks'], d['timed_region_ms'], d['repeats'])
PY
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2c_bench_reference.json 2>/dev/null; cut -c1-200 gpurun_out/r2c_bench_reference.json
export OGB_BENCH_NO_SWEEP=1
bash scratch/profile_all.sh r2c c1 c2 c3 c4 c5 | cut -c1-120
python scratch/latency_public.py > gpurun_out/r2c_latency.txt 2>&1; cat gpurun_out/r2c_latency.txt
