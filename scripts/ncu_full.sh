#!/bin/bash
# Full captures of k_cand and k_penalty at tick ~SKIP of a lone 256-candidate plan (history full).
mkdir -p gpurun_out
SKIP=${1:-600}
export PYTHONPATH=.
for K in k_cand k_penalty k_chain; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c 1 -f -o gpurun_out/prof_$K \
      python scripts/ncu_driver.py 256 > gpurun_out/prof_$K.log 2>&1
done
ls -la gpurun_out | tail -8
