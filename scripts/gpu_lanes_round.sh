#!/bin/bash
# Dev: two-lane pools — parity tests, then throughput with and without lanes.
mkdir -p gpurun_out
export PYTHONPATH=.
python -m pytest tests/test_gpu_headline.py tests/test_gpu_parity.py -m gpu -x -q -k "lanes or pool or solve or selection or capacity" > gpurun_out/lanes_tests.log 2>&1
tail -3 gpurun_out/lanes_tests.log
for lm in 0 512; do
  echo "TOPAY_LANE_MIN_SLOTS=$lm"
  TOPAY_LANE_MIN_SLOTS=$lm python scripts/pool_probe.py --skip-small --no-timed --slots 1024 2048 2>&1 | grep "pool:"
done
