#!/bin/bash
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
export OGB_BENCH_NO_SWEEP=1
bash scratch/ab.sh "c1" "A=1;A=2;OGB_NO_POINT=1"
