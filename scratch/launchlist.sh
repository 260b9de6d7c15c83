#!/bin/bash
# per-kernel durations (ncu launch list) for the given configs; output gpurun_out/launches_<cfg>.csv
for c in $1; do
  python bench.py --config $c --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/plain_$c.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"relabel|gather" -c 12 --csv --log-file gpurun_out/launches_$c.csv \
      python bench.py --config $c --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_list_$c.log 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_$c.csv')) if len(r)>10 and r[0].isdigit()]
for r in rows[:12]: print('$c', r[4][:60], r[-1], r[-2])
PY
done
