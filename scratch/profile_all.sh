#!/bin/bash
# per config: plain bench run, ncu launch list, ncu --set full of the dominant kernels (one capture each)
tag=$1; shift
for c in "$@"; do
  python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${tag}_plain_$c.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"relabel|gather" -c 24 --csv --log-file gpurun_out/${tag}_launches_$c.csv \
      python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:"relabel|gather" -s 4 -c 3 -f -o gpurun_out/${tag}_$c \
      python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${tag}_ncu_$c.log 2>&1
  tail -1 gpurun_out/${tag}_plain_$c.log | cut -c1-400
done
