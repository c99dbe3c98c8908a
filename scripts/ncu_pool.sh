#!/bin/bash
# Full ncu capture of one mid-run launch of each kernel of a tick, with the slot pool full (bench geometry).
mkdir -p gpurun_out
SKIP=${1:-300}
KERNELS=${2:-"k_penalty k_cand k_chain"}
export PYTHONPATH=.
for K in $KERNELS; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c ${3:-1} -f -o gpurun_out/r02_prof_$K \
      python scripts/pool_probe.py --skip-small --plans 6 --slots 1024 --no-timed > gpurun_out/r02_prof_$K.log 2>&1
done
ls -la gpurun_out | tail -6
