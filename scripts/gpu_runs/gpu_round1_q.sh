#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_q.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_q.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --first-instance 1024 > gpurun_out/bench_q_shard1.log 2>&1
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --first-instance 2048 > gpurun_out/bench_q_shard2.log 2>&1
tail -4 gpurun_out/pytest_gpu_q.log
for f in gpurun_out/bench_q_shard1.log gpurun_out/bench_q_shard2.log; do tail -1 $f | python -c "
import json,sys
try:
    l=json.loads(sys.stdin.read()); print('value', round(l['value'],1), 'ms', round(l['ms_per_step'],1), 'solved', l['solved'], l['instances'], 'cycles', l['cycles_per_step'])
except Exception as e: print('fail', e)
"; done
