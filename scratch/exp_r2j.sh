#!/bin/bash
# dynamic tile scheduling (tickets) vs static round-robin tiles, same box
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
bash scratch/ab.sh "c2 c5 c3" "A=1;OGB_STATIC_TILES=1;A=2;OGB_STATIC_TILES=1"
bash scratch/ab.sh "c5b c3b" "A=1;OGB_STATIC_TILES=1"
