#!/bin/bash
mkdir -p gpurun_out
export KEEP_SQ=0
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/field_launches.csv \
    python scripts/field_probe.py > gpurun_out/field_launches.log 2>&1
tail -2 gpurun_out/field_launches.log
