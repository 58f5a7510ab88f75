#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_c.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_c.log
timeout 900 python scripts/sweep_params.py 1024 "" "cg_per_cycle=1" "cg_per_cycle=2" "cg_per_cycle=4" "cg_per_cycle=5" "cg_per_cycle=7" \
  "mu_factor=0.2" "mu_factor=0.3" "cg_per_cycle=2 mu_factor=0.2" "cg_per_cycle=4 mu_factor=0.2" "center_tol=1.0" "center_tol=2.0" "center_tol=8.0" \
  "cg_forcing=0.2" "cg_forcing=0.05" "cg_per_cycle=2 cg_forcing=0.2" > gpurun_out/sweep_c.log 2>&1
timeout 600 python scripts/profile_solve.py 1024 gpurun_out/profile_solve_c.json > gpurun_out/profile_solve_c.log 2>&1
tail -3 gpurun_out/pytest_gpu_c.log; cat gpurun_out/sweep_c.log
