#!/bin/bash
# Dev: launch list of one short bench run + full capture of the dominant kernels. Output in gpurun_out/.
set -x
mkdir -p gpurun_out
CAND=${1:-256}
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --candidates $CAND --plans 1 > gpurun_out/launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_penalty -s 60 -c 2 -f -o gpurun_out/prof_penalty \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --candidates $CAND --plans 1 > gpurun_out/prof_penalty.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_cand -s 60 -c 2 -f -o gpurun_out/prof_cand \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --candidates $CAND --plans 1 > gpurun_out/prof_cand.log 2>&1
ls -la gpurun_out
