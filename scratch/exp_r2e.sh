#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
bash scratch/ab.sh "c1" "A=1;OGB_NO_POINT=1;OGB_INDEX_GRID=8;OGB_INDEX_GRID=64;A=2;OGB_NO_POINT=1 OGB_NO_SHADOW=1"
bash scratch/ab.sh "c2 c5" "A=1;OGB_GATHER_SHAPE=208;OGB_GATHER_SHAPE=216;A=2;OGB_GATHER_SHAPE=208;OGB_GATHER_SHAPE=216"
for c in c2 c4; do python bench.py --config $c --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); e=d['e2e']; print('$c e2e %.4g direct %.4g link %.1f frac %.3f' % (e['value'], e['direct_call_value'], e['link_gbs'], e['frac_of_link']))"; done
