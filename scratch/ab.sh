#!/bin/bash
# usage: scratch/ab.sh "<configs>" "<ENV=VAL settings separated by ;>"
run() { python bench.py --config $1 --steps 100 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('%.4g tr/s  %.3f ms/step  frac %.3f  launches %d' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['gpu_launches']))"; }
IFS=';' read -ra SETTINGS <<< "$2"
for c in $1; do
  for s in "${SETTINGS[@]}"; do
    echo -n "$c [$s]: "; env $s bash -c "$(declare -f run); run $c"
  done
done
