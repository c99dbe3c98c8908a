#!/bin/bash
# Dev: k_penalty occupancy / register variants, timed per-kernel breakdown from the pool probe (run on the GPU box).
for mb in 6 5 4; do
  touch topay_b200/csrc/solver.cu
  make -s -C topay_b200/csrc PTXAS="-DTP_PEN_MIN_BLOCKS=$mb" > /dev/null 2>&1
  echo "== TP_PEN_MIN_BLOCKS=$mb"
  python scripts/pool_probe.py --skip-small --plans 6 --slots 1024 2>&1 | grep timed
done
touch topay_b200/csrc/solver.cu; make -s -C topay_b200/csrc > /dev/null 2>&1
