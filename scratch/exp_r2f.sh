#!/bin/bash
# batch F: occupancy variants of the point kernel, stage size with the 2x16 shape, then the profile captures of the final build
bash scratch/ab.sh "c1" "A=1;OGB_POINT_BLOCKS=5;OGB_POINT_BLOCKS=6;A=2;OGB_POINT_BLOCKS=5"
bash scratch/ab.sh "c2" "A=1;OGB_STAGE_BYTES=6144;OGB_GATHER_SHAPE=308;A=2"
bash scratch/ab.sh "c5" "A=1;OGB_STAGE_BYTES=8192;OGB_GATHER_SHAPE=308"
bash scratch/ab.sh "c3 c4" "A=1"
