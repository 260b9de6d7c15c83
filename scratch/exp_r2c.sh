#!/bin/bash
# round-2 experiment batch: parity first, then A/B of the new switches
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
bash scratch/ab.sh "c1" "A=1;OGB_NO_SHADOW=1"
bash scratch/ab.sh "c2 c5" "A=1;OGB_NO_PREFETCH=1;OGB_GATHER_SHAPE=212;OGB_GATHER_SHAPE=216;OGB_GATHER_SHAPE=312 OGB_STAGE_BYTES=2048;OGB_GATHER_SHAPE=408 OGB_STAGE_BYTES=2048"
bash scratch/ab.sh "c3" "A=1;OGB_GATHER_SHAPE=212;OGB_GATHER_SHAPE=216;OGB_GATHER_SHAPE=212 OGB_STAGE_BYTES=4096"
