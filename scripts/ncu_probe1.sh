#!/bin/bash
# Dev: single-candidate latency anatomy (per-kernel durations + k_cand source hot spots)
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 400 --csv --log-file gpurun_out/launches1.csv \
    python scripts/quick_cfg3.py 1 > gpurun_out/launches1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_cand -s 400 -c 1 -f -o gpurun_out/prof_cand1 \
    python scripts/quick_cfg3.py 1 > gpurun_out/prof_cand1.log 2>&1
ls -la gpurun_out | tail -5
