#!/bin/bash
# Round-2 profile captures: EDT kernels (full sets), launch list of the pool with two lanes.
mkdir -p gpurun_out
export PYTHONPATH=.
KEEP_SQ=0 timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_edt -s 13 -c 13 -f -o gpurun_out/r02_prof_edt2 \
    python scripts/field_probe.py > gpurun_out/prof_edt2.log 2>&1
tail -2 gpurun_out/prof_edt2.log
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 480 --csv --log-file gpurun_out/r02_launches_pool2.csv \
    python scripts/pool_probe.py --skip-small --plans 8 --slots 2048 --no-timed > gpurun_out/r02_launches_pool2.log 2>&1
python scripts/ncu_summary.py gpurun_out/r02_launches_pool2.csv | head -12
