#!/bin/bash
# Launch list (durations) of a reference-scale plan mid-solve: kernel time vs tick time = launch-gap share.
mkdir -p gpurun_out
export PYTHONPATH=.
cat > /tmp/small_one.py <<'PY'
import topay_b200 as tp
from topay_b200 import scenes
pts, _ = scenes.cuboids_scene(42)
gm = tp.GridMap(tp.grid_desc()); gm.regenerateMap(pts)
s = tp.MomaTrajOpt(gm, max_cand=8, max_pieces=16)
paths, bv, ba = scenes.short_candidates(8, 5001)
r = s.optimizeTrajBatch(paths, bv, ba)
print(r["evals"], s.stats()["ms_total"], s.stats()["ticks"])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 400 --csv --log-file gpurun_out/r02_launches_small.csv \
    python /tmp/small_one.py > gpurun_out/r02_launches_small.log 2>&1
python scripts/ncu_summary.py gpurun_out/r02_launches_small.csv | head -8
python /tmp/small_one.py
