#!/bin/bash
export OGB_BENCH_NO_SWEEP=1
for c in c1 c2 c3 c4 c5 c5b; do
python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('$c', r['kernel'], '%.3f'%r['frac'], r['traffic'], r['traffic_source'])"
done
timeout 600 python -m pytest tests/test_gpu_bench.py -m gpu -x -q 2>&1 | tail -2
