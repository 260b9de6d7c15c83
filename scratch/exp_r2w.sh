#!/bin/bash
export OGB_BENCH_NO_SWEEP=1
bash scratch/ab.sh "c1" "A=1;OGB_POINT_NA=1;A=2;OGB_POINT_NA=1"
