#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_d.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_d.log
timeout 900 python scripts/sweep_params.py 1024 "" "cg_per_cycle=5" "cg_per_cycle=7" "cg_forcing=0.2" "cg_per_cycle=5 cg_forcing=0.2" > gpurun_out/sweep_d.log 2>&1
timeout 600 python scripts/profile_solve.py 1024 gpurun_out/profile_solve_d.json > gpurun_out/profile_solve_d.log 2>&1
tail -30 gpurun_out/pytest_gpu_d.log; cat gpurun_out/sweep_d.log
