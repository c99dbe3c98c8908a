#!/bin/bash
# compute-sanitizer over the kernels added this round (thread-per-line EDT kernels, visibility rays, two-lane pools).
mkdir -p gpurun_out
export PYTHONPATH=.
for tool in memcheck racecheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python -m pytest tests/test_gpu_parity.py tests/test_gpu_headline.py -m gpu -x -q \
      -k "thread_per_line or visibility or two_lanes" > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool"; grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY" gpurun_out/sanitize_$tool.log | tail -3
done
