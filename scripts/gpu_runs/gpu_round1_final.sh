#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_final.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_final.log 2>&1
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_final.log 2>&1
timeout 600 python scripts/profile_solve.py 1024 gpurun_out/profile_solve_final.json > gpurun_out/profile_solve_final.log 2>&1
tail -3 gpurun_out/pytest_gpu_final.log; tail -1 gpurun_out/smoke_final.log; tail -1 gpurun_out/bench_final.log | cut -c1-250
