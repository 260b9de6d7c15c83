#!/bin/bash
# batch G: the final build -- whole GPU suite, record-stride A/B, profile captures, the all-config bench line, latency
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
bash scratch/ab.sh "c2" "A=1;OGB_RECORD_ALIGN=64;OGB_RECORD_ALIGN=16"
bash scratch/profile_all.sh r2 c1 c2 c3 c4 c5
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; tail -c 300 gpurun_out/r2_bench_default.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench_reference.json 2>/dev/null
python scratch/latency_public.py c2 c3 > gpurun_out/r2_latency.txt 2>&1; tail -20 gpurun_out/r2_latency.txt
