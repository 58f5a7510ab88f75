#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_j.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_j.log
for s in 1 2 4 8; do
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --streams $s > gpurun_out/bench_j_s$s.log 2>&1
done
tail -5 gpurun_out/pytest_gpu_j.log
for s in 1 2 4 8; do tail -1 gpurun_out/bench_j_s$s.log | python -c "
import json,sys
try:
    l=json.loads(sys.stdin.read()); print('streams', l['e2e']['streams'], 'value', round(l['value'],1), 'e2e', round(l['e2e']['value'],1), 'ms', round(l['ms_per_step'],1))
except Exception as e: print('fail', e)
"; done
