#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_b.log
timeout 900 python scripts/sweep_params.py 1024 "" "cg_per_cycle=2" "cg_per_cycle=4" "cg_per_cycle=5" "cg_per_cycle=6" \
  "mu_factor=0.2" "mu_factor=0.05" "mu_factor=0.03" "coarse_every=2" "coarse_every=3" "coarse_every=5" \
  "center_tol=1.0" "center_tol=10.0" "center_tol=30.0" "cg_forcing=0.3" "cg_forcing=0.03" \
  "cg_per_cycle=4 mu_factor=0.05" "cg_per_cycle=4 coarse_every=3" "cg_per_cycle=2 cg_forcing=0.3" \
  "cg_per_cycle=4 mu_factor=0.05 coarse_every=3 center_tol=10.0" > gpurun_out/sweep_b.log 2>&1
timeout 600 python scripts/profile_solve.py 1024 gpurun_out/profile_solve_b.json > gpurun_out/profile_solve_b.log 2>&1
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b.log 2>&1
tail -3 gpurun_out/pytest_gpu_b.log; cat gpurun_out/sweep_b.log; tail -2 gpurun_out/bench_b.log | cut -c1-400
