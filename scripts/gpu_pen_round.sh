#!/bin/bash
# Dev: solver parity tests + pool throughput + small-plan breakdown after a k_penalty change.
mkdir -p gpurun_out
export PYTHONPATH=.
python -m pytest tests/test_gpu_parity.py tests/test_gpu_headline.py tests/test_gpu_rog.py tests/test_gpu_traj.py -m gpu -x -q -k "not field and not baseline_size" > gpurun_out/pen_tests.log 2>&1
tail -3 gpurun_out/pen_tests.log
python scripts/pool_probe.py --skip-small --slots 2048 2>&1 | grep "pool:\|penalty\|k_pen" | head -5
python scripts/small_breakdown.py 2>&1 | tail -2
