#!/bin/bash
# session 3, call d: look-ahead pipelined coarse sweep + paired-thread line search — parity subset, sweep timing,
# per-kernel profile, A/B of the line-search occupancy hint, source-level ncu of the LS-tick kernels
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "coarse or bitwise or solution_parity or edge_case" ) > gpurun_out/pytest_gpu_s2d.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_s2d.log
timeout 200 python scripts/sweep_params.py 1024 "" 2>&1 | grep -v "    inst" > gpurun_out/sweep_s2d.log
timeout 200 python scripts/profile_solve.py 1024 gpurun_out/profile_solve_s2d.json > gpurun_out/profile_solve_s2d.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_coarse_build|k_linesearch" -s 2 -c 2 -o gpurun_out/prof_s2d python scripts/profile_solve.py 1024 x ncu=1 > gpurun_out/ncu_s2d.log 2>&1
# variant: 2 CTAs / SM for the line search (no spills, 126 registers)
sed -i 's/__launch_bounds__(kThreads, 3) k_linesearch/__launch_bounds__(kThreads, 2) k_linesearch/' score_b200/csrc/solver.cuh
timeout 200 python scripts/profile_solve.py 1024 gpurun_out/profile_solve_s2d_lb2.json > gpurun_out/profile_solve_s2d_lb2.log 2>&1
tail -5 gpurun_out/pytest_gpu_s2d.log; cat gpurun_out/sweep_s2d.log; grep -A12 kernel_ms_total gpurun_out/profile_solve_s2d.json;  grep -A3 kernel_ms_total gpurun_out/profile_solve_s2d_lb2.json
