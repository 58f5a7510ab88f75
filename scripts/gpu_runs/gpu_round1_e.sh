#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_e.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_e.log
timeout 900 python scripts/sweep_params.py 1024 "" "cg_per_cycle=5" "cg_per_cycle=5 cg_forcing=0.2" > gpurun_out/sweep_e.log 2>&1
timeout 600 python scripts/profile_solve.py 1024 gpurun_out/profile_solve_e.json > gpurun_out/profile_solve_e.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_coarse_build|k_linesearch|k_rowupdate" -s 9 -c 6 -o gpurun_out/prof_ls python scripts/profile_solve.py 1024 x ncu=1 > gpurun_out/ncu_ls.log 2>&1
tail -5 gpurun_out/pytest_gpu_e.log; cat gpurun_out/sweep_e.log
