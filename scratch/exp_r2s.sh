#!/bin/bash
export OGB_BENCH_NO_SWEEP=1
bash scratch/ab.sh "c2 c5" "A=1;OGB_GATHER_SHAPE=224;A=2"
bash scratch/ab.sh "c4" "A=1;OGB_BAND_ROWS=16;OGB_BAND_ROWS=64"
bash scratch/ab.sh "c1" "A=1;OGB_INDEX_GRID=5;OGB_INDEX_GRID=20"
