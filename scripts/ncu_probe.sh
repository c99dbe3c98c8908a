#!/bin/bash
# Launch list of one short bench run + full captures of the two dominant kernels mid-solve (L-BFGS
# history full). Output in gpurun_out/; summaries are copied to profiles/ by scripts/ncu_summary.py.
set -x
mkdir -p gpurun_out
CAND=${1:-256}
ARGS="--steps 1 --warmup 0 --no-cpu-baseline --no-extras --candidates $CAND --plans 1"
ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py $ARGS > gpurun_out/launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_penalty -s 1000 -c 1 -f -o gpurun_out/prof_penalty \
    python bench.py $ARGS > gpurun_out/prof_penalty.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_cand -s 1000 -c 1 -f -o gpurun_out/prof_cand \
    python bench.py $ARGS > gpurun_out/prof_cand.log 2>&1
ls -la gpurun_out
