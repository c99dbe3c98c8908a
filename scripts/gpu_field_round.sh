#!/bin/bash
# Dev: field parity tests, 800x800x80 rebuild timing, launch list and full captures of the EDT kernels.
mkdir -p gpurun_out
export PYTHONPATH=.
python -m pytest tests/test_gpu_parity.py tests/test_gpu_rog.py tests/test_gpu_headline.py -m gpu -x -q -k "field or rog or baseline or raster" > gpurun_out/field_tests.log 2>&1
tail -4 gpurun_out/field_tests.log
KEEP_SQ=0 python scripts/field_probe.py 2>&1 | tail -3
for cfg in $SWEEP; do IFS=x read tz ty <<< "$cfg"; echo "TZ=$tz TY=$ty"; TOPAY_EDT_TZ=$tz TOPAY_EDT_TY=$ty KEEP_SQ=0 python scripts/field_probe.py 2>&1 | tail -2; done
KEEP_SQ=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_field.csv \
    python scripts/field_probe.py > gpurun_out/field_launches.log 2>&1
python scripts/ncu_summary.py gpurun_out/r02_launches_field.csv | head -8
if [ "$1" != "nofull" ]; then
KEEP_SQ=0 timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_edt -s 8 -c 3 -f -o gpurun_out/r02_prof_edt \
    python scripts/field_probe.py > gpurun_out/prof_edt.log 2>&1
fi
