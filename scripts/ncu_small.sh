#!/bin/bash
# Full capture of one mid-solve k_cand launch of a reference-scale plan (8 candidates, fused launch).
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
print(r["evals"])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_cand -s ${1:-600} -c 2 -f -o gpurun_out/r02_prof_small_cand \
    python /tmp/small_one.py > gpurun_out/r02_prof_small_cand.log 2>&1
tail -2 gpurun_out/r02_prof_small_cand.log
