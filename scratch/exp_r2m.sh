#!/bin/bash
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python scratch/fuzz.py 300 777 2>&1 | tail -3
export OGB_BENCH_NO_SWEEP=1
bash scratch/ab.sh "c1" "A=1;OGB_NO_WIDE_RECORD=1;OGB_INDEX_GRID=16;A=2"
bash scratch/ab.sh "c2 c4" "A=1"
