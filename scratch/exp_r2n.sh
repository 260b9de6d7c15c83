#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_lookahead.py tests/test_gpu_philox.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4
export OGB_BENCH_NO_SWEEP=1
for c in c2 c5 c3 c4 c1; do
python bench.py --config $c --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); e=d['e2e']; print('$c e2e %.4g tr/s  direct %.4g  link %.3g GB/s  frac %.3f  d2h %d' % (e['value'], e['direct_call_value'], e['link_gbs'], e['frac_of_link'], e['d2h_bytes_per_step']))"
done
