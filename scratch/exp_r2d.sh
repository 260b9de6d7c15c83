#!/bin/bash
# round-2 batch D: parity of the new features first, then: where is the limit?  L2-resident datasets vs the full ones;
# one-CTA-per-SM shapes; the all-config bench line
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
bash scratch/ab.sh "c2" "A=1;OGB_GATHER_SHAPE=216;OGB_GATHER_SHAPE=316;OGB_GATHER_SHAPE=220;OGB_BENCH_EPISODES=100;OGB_BENCH_EPISODES=100 OGB_GATHER_SHAPE=216"
bash scratch/ab.sh "c5" "A=1;OGB_GATHER_SHAPE=316;OGB_GATHER_SHAPE=220;OGB_BENCH_EPISODES=200"
bash scratch/ab.sh "c1" "A=1;OGB_BENCH_EPISODES=10"
bash scratch/ab.sh "c3" "A=1;OGB_BENCH_EPISODES=30;OGB_NO_OVERLAP=1"
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_all_r2d.json 2> gpurun_out/bench_all_r2d.err; tail -c 600 gpurun_out/bench_all_r2d.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_all_r2d.json').read().strip().splitlines()[-1])
print('headline', d['value'], d['ms_per_step'], d['roofline']['frac'], 'repeats', d['repeats'], d['timed_region_ms'], d['clocks'], 'e2e', d['e2e']['value'], d['e2e'].get('link_gbs'), d['e2e'].get('frac_of_link'))
for k,c in d['configs'].items(): print(k, '%.4g'%c['value'], '%.4f ms'%c['ms_per_step'], 'frac %.3f step %.3f'%(c['roofline']['frac'], c['roofline']['step_frac']), c['roofline']['kernel'], 'traffic', c['roofline']['traffic'], c['roofline']['traffic_source'], 'e2e %.4g'%c['e2e']['value'], 'cpu %.4g'%c['cpu_baseline']['value'])
PY
python scratch/latency_public.py 2>&1 | tail -12
