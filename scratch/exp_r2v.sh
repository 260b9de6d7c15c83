#!/bin/bash
for st in 100 20 100; do
python bench.py --config c2 --steps $st --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); e=d['e2e']; print('sweep on, steps $st: e2e %.4g tr/s  direct %.4g  link %.3g GB/s  frac %.3f  steps_e %d' % (e['value'], e['direct_call_value'], e['link_gbs'], e['frac_of_link'], e['steps']))"
done
python bench.py 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
for k,v in d['configs'].items():
    e=v['e2e']; print(k, 'e2e %.4g direct %.4g link %.3g frac %.3f' % (e['value'], e['direct_call_value'], e['link_gbs'], e['frac_of_link']))"
