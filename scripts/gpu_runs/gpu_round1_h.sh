#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_h.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_h.log
timeout 900 python scripts/sweep_params.py 1024 "" "cg_per_cycle=3" "cg_per_cycle=5" "cg_per_cycle=2" > gpurun_out/sweep_h.log 2>&1
timeout 600 python scripts/profile_solve.py 1024 gpurun_out/profile_solve_h.json > gpurun_out/profile_solve_h.log 2>&1
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_h.log 2>&1
tail -5 gpurun_out/pytest_gpu_h.log; cat gpurun_out/sweep_h.log; tail -1 gpurun_out/bench_h.log | cut -c1-200
