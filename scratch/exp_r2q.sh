#!/bin/bash
export OGB_BENCH_NO_SWEEP=1
bash scratch/ab.sh "c2" "A=1;OGB_GATHER_SHAPE=218;OGB_GATHER_SHAPE=220;OGB_GATHER_SHAPE=222"
bash scratch/ab.sh "c5" "A=1;OGB_GATHER_SHAPE=218;OGB_GATHER_SHAPE=220;OGB_GATHER_SHAPE=222;OGB_GATHER_SHAPE=308"
python bench.py --config c2 --steps 50 --warmup 3 --no-cpu-baseline --no-e2e --batches-per-launch 1024 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('c2 L=1024 216: frac %.3f' % d['roofline']['frac'])"
OGB_GATHER_SHAPE=220 python bench.py --config c2 --steps 50 --warmup 3 --no-cpu-baseline --no-e2e --batches-per-launch 1024 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('c2 L=1024 220: frac %.3f' % d['roofline']['frac'])"
