#!/bin/bash
# session 3, call e: pipelined coarse sweep with sleeping waiters
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "coarse or bitwise or solution_parity or edge_case" ) > gpurun_out/pytest_gpu_s2e.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_s2e.log
timeout 200 python scripts/profile_solve.py 1024 gpurun_out/profile_solve_s2e.json > gpurun_out/profile_solve_s2e.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_coarse_build" -s 1 -c 1 -o gpurun_out/prof_s2e python scripts/profile_solve.py 1024 x ncu=1 > gpurun_out/ncu_s2e.log 2>&1
tail -5 gpurun_out/pytest_gpu_s2e.log; grep -A12 kernel_ms_total gpurun_out/profile_solve_s2e.json; grep "solve_ms\|cycles" gpurun_out/profile_solve_s2e.json
