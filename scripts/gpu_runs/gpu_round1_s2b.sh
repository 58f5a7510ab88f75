#!/bin/bash
# session 3, call b: GPU tests (incl. the new evaluation tests) + sweep timing + per-kernel profile after the line-search rewrite
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_s2b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_s2b.log
timeout 500 python scripts/sweep_params.py 1024 "" 2>&1 | grep -v "    inst" > gpurun_out/sweep_s2b.log
timeout 600 python scripts/profile_solve.py 1024 gpurun_out/profile_solve_s2b.json > gpurun_out/profile_solve_s2b.log 2>&1
tail -8 gpurun_out/pytest_gpu_s2b.log; cat gpurun_out/sweep_s2b.log; grep -A14 kernel_ms_total gpurun_out/profile_solve_s2b.json
