#!/bin/bash
# Development aid: timing-only ablations of k_pbs_pipe (results are wrong by construction): which exchange costs what.
# FHE_PBS_DBG bits: 1 = rotated-accumulator pushes stay local, 2 = forward copies stay local, 4 = inverse copies stay local
for d in 0 1 2 4 7; do
  echo "== FHE_PBS_DBG=$d"
  FHE_PBS_DBG=$d timeout 200 python scripts/quick_bench.py p8 2>&1 | tail -1
done
