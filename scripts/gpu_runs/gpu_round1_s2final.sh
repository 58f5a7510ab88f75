#!/bin/bash
# session 3 measurement pass: full GPU tests, smoke, bench (both arms), whole-solve kernel profile, ncu launch list, ncu full capture
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_s2final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_s2final.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_s2final.log 2>&1
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_s2final.log 2>&1
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_s2final.log 2>&1
timeout 300 python scripts/profile_solve.py 1024 gpurun_out/profile_solve_s2final.json > gpurun_out/profile_solve_s2final.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 400 --csv --log-file gpurun_out/launches_s2final.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --streams 1 > gpurun_out/ncu_launch_s2final.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_rowpass|k_colpass|k_precond|k_coarse|k_linesearch|k_rowupdate|k_pupdate" -s 0 -c 22 -o gpurun_out/prof_s2final python scripts/profile_solve.py 1024 x ncu=1 > gpurun_out/ncu_full_s2final.log 2>&1
tail -3 gpurun_out/pytest_gpu_s2final.log; tail -1 gpurun_out/smoke_s2final.log; tail -1 gpurun_out/bench_s2final.log | cut -c1-400; tail -1 gpurun_out/bench_ref_s2final.log | cut -c1-300
# tail experiments: growing PCG ticks per cycle, sub-batch stream count
timeout 300 python scripts/sweep_params.py 1024 "" "cg_grow_after=40 cg_grow_every=8" "cg_grow_after=60 cg_grow_every=8" "cg_grow_after=60 cg_grow_every=16" "cg_per_cycle=6" "cg_per_cycle=3" 2>&1 | grep -v "    inst" > gpurun_out/sweep_s2final.log
for sp in "1 1" "4 4" "2 4" "2 8" "3 6" "4 8" "4 16"; do set -- $sp; timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --streams $1 --parts $2 > gpurun_out/bench_s${1}_p${2}_s2final.log 2>&1; done
SCORE_SPLIT_COARSE_APPLY=1 timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/bench_split_s2final.log 2>&1
cat gpurun_out/sweep_s2final.log; for f in gpurun_out/bench_s*_p*_s2final.log gpurun_out/bench_split_s2final.log; do echo $f; tail -1 $f | python -c "
import json,sys
try:
    l=json.loads(sys.stdin.read()); print(round(l['value'],1), 'e2e', l.get('e2e') and round(l['e2e']['value'],1), 'ms', round(l['ms_per_step'],1))
except Exception as e: print('fail', e)
"; done
