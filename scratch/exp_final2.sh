#!/bin/bash
# default bench line (all configs), reference arm, then per-config launch lists and ncu --set full captures
python bench.py --steps 20 --warmup 5 > gpurun_out/r2b_bench_default.json 2> gpurun_out/r2b_bench_default.err
tail -c 600 gpurun_out/r2b_bench_default.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2b_bench_default.json').read().strip().splitlines()[-1])
for k,v in d['configs'].items():
    r=v['roofline']; print(k, 'L=%d'%v['batches_per_launch'], '%.4g tr/s'%v['value'], '%.4f ms'%v['ms_per_step'], 'frac %.3f step %.3f'%(r['frac'], r['step_frac']), 'traffic', r['traffic'], r['traffic_source'][:60], [(s['batches_per_launch'], round(s['frac'],3)) for s in v['launch_size_sweep']], 'e2e %.3g'%v['e2e']['value'])
print(d['clocks'])
PY
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2b_bench_reference.json 2>/dev/null; cut -c1-300 gpurun_out/r2b_bench_reference.json
export OGB_BENCH_NO_SWEEP=1
bash scratch/profile_all.sh r2b c1 c2 c3 c4 c5
