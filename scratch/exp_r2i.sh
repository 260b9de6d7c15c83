#!/bin/bash
# launch-size sweep: how much of a launch is fixed cost (ramp-up, tail, launch gap)?
run() { python bench.py --config $1 --steps 50 --warmup 3 --no-cpu-baseline --no-e2e --batches-per-launch $2 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1 L=$2: %.4g tr/s  %.4f ms/step  kernel %.4f ms  frac %.3f step_frac %.3f' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['roofline'].get('step_frac', 0)))"; }
for L in 256 1024 4096 8192; do run c2 $L; done
for L in 64 256 1024; do run c5 $L; done
for L in 256 1024 2048; do run c3 $L; done
for L in 256 1024 4096 16384; do run c1 $L; done
for L in 8 16 64; do run c4 $L; done
