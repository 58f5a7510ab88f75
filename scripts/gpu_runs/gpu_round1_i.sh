#!/bin/bash
mkdir -p gpurun_out
for m in 0 1 2 4; do
  echo "== SCORE_WGRID_MULT=$m" >> gpurun_out/sweep_i.log
  SCORE_WGRID_MULT=$m timeout 600 python scripts/sweep_params.py 1024 "" "cg_per_cycle=3" >> gpurun_out/sweep_i.log 2>&1
done
SCORE_WGRID_MULT=0 timeout 600 python scripts/profile_solve.py 1024 gpurun_out/profile_solve_i.json > gpurun_out/profile_solve_i.log 2>&1
cat gpurun_out/sweep_i.log
