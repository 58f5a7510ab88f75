#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_f.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_f.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_f.log 2>&1
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_f.log 2>&1
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref_f.log 2>&1
timeout 600 python scripts/profile_solve.py 1024 gpurun_out/profile_solve_f.json > gpurun_out/profile_solve_f.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 400 --csv --log-file gpurun_out/launches_f.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch_f.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_rowpass|k_colpass|k_precond|k_coarse|k_linesearch|k_rowupdate|k_pupdate" -s 60 -c 30 -o gpurun_out/prof_f python scripts/profile_solve.py 1024 x ncu=1 > gpurun_out/ncu_full_f.log 2>&1
tail -3 gpurun_out/pytest_gpu_f.log; tail -2 gpurun_out/smoke_f.log; tail -1 gpurun_out/bench_f.log | cut -c1-300; tail -1 gpurun_out/bench_ref_f.log | cut -c1-300
