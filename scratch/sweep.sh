#!/bin/bash
# usage: scratch/sweep.sh "<configs>" ; prints one line per (config, knob setting)
run() { python bench.py --config $1 --steps 100 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('%.4g tr/s  %.3f ms/step  frac %.3f  launches %d' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['gpu_launches']))"; }
for c in $1; do
  for stage in 2048 4096 8192; do for flat in 0 1; do
    echo -n "$c stage=$stage flat=$flat chunks=1: "; OGB_STAGE_BYTES=$stage OGB_DRAIN_FLAT=$flat OGB_CHUNKS=1 run $c
  done; done
  for ch in 2 4; do echo -n "$c stage=4096 flat=1 chunks=$ch: "; OGB_STAGE_BYTES=4096 OGB_DRAIN_FLAT=1 OGB_CHUNKS=$ch run $c; done
done
