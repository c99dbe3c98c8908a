#!/bin/bash
# full capture of the int32 strided pass of one 800x800x80 rebuild
mkdir -p gpurun_out
export KEEP_SQ=0
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_edt_strided32 -s 2 -c 1 -f -o gpurun_out/prof_edt32 \
    python scripts/field_probe.py > gpurun_out/prof_edt32.log 2>&1
ls -la gpurun_out | grep -E "edt32"
