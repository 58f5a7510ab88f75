#!/bin/bash
# first GPU call of this session: tests, whole-solve breakdown, bench, ncu launch list + full capture
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python scripts/profile_solve.py 1024 gpurun_out/profile_solve.json > gpurun_out/profile_solve.log 2>&1
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 440 --csv --log-file gpurun_out/launches_ipm.csv python scripts/profile_solve.py 1024 x ncu=1 > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_rowpass|k_colpass|k_precond|k_coarse|k_linesearch|k_rowupdate|k_pupdate" -s 270 -c 27 -o gpurun_out/prof_ipm python scripts/profile_solve.py 1024 x ncu=1 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
tail -3 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/bench.log | cut -c1-600
