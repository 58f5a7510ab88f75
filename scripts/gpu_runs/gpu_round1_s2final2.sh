#!/bin/bash
# session 3 measurement pass (final code): full GPU tests, smoke, bench, whole-solve kernel profile, ncu launch list, ncu full capture
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_s2final2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_s2final2.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_s2final2.log 2>&1
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_s2final2.log 2>&1
timeout 300 python scripts/profile_solve.py 1024 gpurun_out/profile_solve_s2final2.json > gpurun_out/profile_solve_s2final2.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 400 --csv --log-file gpurun_out/launches_s2final2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --streams 1 > gpurun_out/ncu_launch_s2final2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_rowpass|k_colpass|k_precond|k_coarse|k_linesearch|k_rowupdate|k_pupdate" -s 0 -c 20 -o gpurun_out/prof_s2final2 python scripts/profile_solve.py 1024 x ncu=1 > gpurun_out/ncu_full_s2final2.log 2>&1
timeout 300 python scripts/e2e_trace.py 1024 2 3 1 > gpurun_out/e2e_trace_final2.log 2>&1
tail -3 gpurun_out/pytest_gpu_s2final2.log; tail -1 gpurun_out/smoke_s2final2.log; tail -1 gpurun_out/bench_s2final2.log | cut -c1-300; tail -8 gpurun_out/e2e_trace_final2.log
