#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_u.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_u.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_u.log 2>&1
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_u.log 2>&1
timeout 600 python scripts/profile_solve.py 1024 gpurun_out/profile_solve_u.json > gpurun_out/profile_solve_u.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 400 --csv --log-file gpurun_out/launches_u.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --streams 1 > gpurun_out/ncu_launch_u.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_rowpass|k_colpass|k_precond|k_coarse|k_linesearch|k_rowupdate|k_pupdate" -s 0 -c 40 -o gpurun_out/prof_u python scripts/profile_solve.py 1024 x ncu=1 > gpurun_out/ncu_full_u.log 2>&1
tail -3 gpurun_out/pytest_gpu_u.log; tail -2 gpurun_out/smoke_u.log; tail -1 gpurun_out/bench_u.log | cut -c1-300
