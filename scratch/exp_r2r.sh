#!/bin/bash
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
export OGB_BENCH_NO_SWEEP=1
bash scratch/ab.sh "c2 c5 c3 c5b c3b" "A=1"
