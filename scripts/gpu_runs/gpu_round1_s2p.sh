#!/bin/bash
# session 3, call p: tail graphs (small grids for the last cycles of a batch) — parity subset + A/B
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bitwise or solution_parity or edge_case" ) > gpurun_out/pytest_gpu_s2p.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_s2p.log
for v in "16 128" "0 128" "32 128" "64 256" "16 32" "128 512"; do set -- $v; echo "TAIL_INST=$1 TAIL_GRID=$2"; SCORE_TAIL_INST=$1 SCORE_TAIL_GRID=$2 timeout 200 python scripts/sweep_params.py 1024 "" 2>&1 | grep -v "    inst"; done > gpurun_out/sweep_s2p.log 2>&1
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_s2p.log 2>&1
SCORE_TAIL_INST=0 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_s2p_off.log 2>&1
tail -4 gpurun_out/pytest_gpu_s2p.log; cat gpurun_out/sweep_s2p.log | cut -c1-200; for f in gpurun_out/bench_s2p.log gpurun_out/bench_s2p_off.log; do tail -1 $f | python -c "
import json,sys
try:
    l=json.loads(sys.stdin.read()); print(round(l['value'],1), 'ms', round(l['ms_per_step'],1))
except Exception as e: print('fail', e)
"; done
