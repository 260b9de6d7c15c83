#!/bin/bash
export OGB_BENCH_NO_SWEEP=1
for st in 20 100 20 100; do
python bench.py --config c2 --steps $st --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); e=d['e2e']; print('steps $st: e2e %.4g tr/s  direct %.4g  link %.3g GB/s  frac %.3f  steps_e %d' % (e['value'], e['direct_call_value'], e['link_gbs'], e['frac_of_link'], e['steps']))"
done
