#!/bin/bash
# usage: scratch/ncu_full.sh <tag> "<configs>" [ENV=VAL ...]   -> gpurun_out/<tag>_<cfg>.ncu-rep
tag=$1; cfgs=$2; shift 2
for c in $cfgs; do
  env "$@" python bench.py --config $c --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/plain_$c.log 2>&1 &&
  env "$@" ncu --set full --clock-control none --import-source on -k regex:"relabel|gather" -s ${NCU_SKIP:-6} -c 2 -f -o gpurun_out/${tag}_$c \
      python bench.py --config $c --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_$c.log 2>&1
  tail -2 gpurun_out/ncu_$c.log | cut -c1-200
done
