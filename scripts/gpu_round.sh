#!/bin/bash
# Round-2 measurement round: whole GPU suite, default bench line, lone-plan lanes probe.
mkdir -p gpurun_out
export PYTHONPATH=.
python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputest_b.log 2>&1; tail -3 gpurun_out/r02_gputest_b.log
python bench.py > gpurun_out/r02_bench_d.json 2> gpurun_out/r02_bench_d.err; tail -c 600 gpurun_out/r02_bench_d.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_d.json").read().strip().splitlines()[-1])
for k in ("value", "e2e", "single_plan", "latency", "field", "sweep"):
    print(k, d.get(k))
print("shares", d["roofline"]["kernel_shares"], d["roofline"]["frac"], d["roofline_k_lbfgs"]["frac"])
PY
for lm in 0 128; do echo "lone plan, TOPAY_LANE_MIN_SLOTS=$lm"; TOPAY_LANE_MIN_SLOTS=$lm python scripts/pool_probe.py --skip-small --no-timed --plans 1 --slots 256 2>&1 | grep "pool:"; done
