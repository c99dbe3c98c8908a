#!/bin/bash
# per-kernel durations of one 800x800x80 rebuild + full capture of the two strided passes
mkdir -p gpurun_out
export KEEP_SQ=0
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/field_launches.csv \
    python scripts/field_probe.py > gpurun_out/field_launches.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_edt_strided -s 10 -c 2 -f -o gpurun_out/prof_edt \
    python scripts/field_probe.py > gpurun_out/prof_edt.log 2>&1
ls -la gpurun_out | grep -E "field|edt"
