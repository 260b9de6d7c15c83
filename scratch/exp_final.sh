#!/bin/bash
# final tree: the whole GPU suite, smoke(), the driver's two bench commands
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2
( time python bench.py --steps 20 --warmup 5 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err ) 2>&1 | grep real; tail -c 300 gpurun_out/final_bench.err
( time python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/final_bench_reference.json 2>/dev/null ) 2>&1 | grep real
python - <<'PY'
import json
d=json.loads(open('gpurun_out/final_bench.json').read().strip().splitlines()[-1])
r=json.loads(open('gpurun_out/final_bench_reference.json').read().strip().splitlines()[-1])
print('headline %.4g tr/s frac %.3f e2e %.4g (x%.2f vs ref %.4g) direct %.4g link %.1f' % (d['value'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['value']/r['value'], r['value'], d['e2e']['direct_call_value'], d['e2e']['link_gbs']), d['clocks'], 'same_config', d['config']==r['config'])
for k,c in d['configs'].items(): print(k, '%.4g'%c['value'], 'frac %.3f step %.3f'%(c['roofline']['frac'], c['roofline']['step_frac']), 'traffic', c['roofline']['traffic'], 'e2e %.4g'%c['e2e']['value'])
PY
