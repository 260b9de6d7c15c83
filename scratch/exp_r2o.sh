#!/bin/bash
export OGB_BENCH_NO_SWEEP=1
bash scratch/ab.sh "c3" "A=1;OGB_FUSE=1;OGB_NO_OVERLAP=1"
bash scratch/ab.sh "c5b" "A=1;OGB_FUSE=1"
bash scratch/ab.sh "c2" "A=1;OGB_GATHER_SHAPE=308;OGB_GATHER_SHAPE=316;OGB_GATHER_SHAPE=220;OGB_FUSE=0"
bash scratch/ab.sh "c5" "A=1;OGB_GATHER_SHAPE=308;OGB_GATHER_SHAPE=316;OGB_FUSE=0"
