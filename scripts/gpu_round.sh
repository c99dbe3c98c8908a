#!/bin/bash
# Dev: whole GPU suite, field timing + launch list, small-plan latency, pool throughput with per-kernel times.
mkdir -p gpurun_out
export PYTHONPATH=.
python -m pytest tests -m gpu -x -q > gpurun_out/gputest.log 2>&1; tail -4 gpurun_out/gputest.log
KEEP_SQ=0 python scripts/field_probe.py 2>&1 | tail -2
KEEP_SQ=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_field.csv \
    python scripts/field_probe.py > gpurun_out/field_launches.log 2>&1
python scripts/ncu_summary.py gpurun_out/r02_launches_field.csv | head -8
python scripts/latency_probe.py 2>&1 | tail -1
python scripts/pool_probe.py --skip-small --plans 6 --slots 1024 2>&1 | tail -2
