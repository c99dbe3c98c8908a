#!/bin/bash
# ncu launch list of the bench command (one lone plan, 100 ticks mid-solve): per-launch durations are
# cold-cache and serialised, use the SHARES.
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-extras --candidates 256 --plans 1 > gpurun_out/launches_bench.log 2>&1
tail -2 gpurun_out/launches_bench.log | cut -c1-300
