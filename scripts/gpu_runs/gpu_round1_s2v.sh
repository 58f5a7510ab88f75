#!/bin/bash
# session 3, call v: create without driver round trips for host inputs — full GPU tests, e2e spread, bench
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_s2v.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_s2v.log
for r in 1 2 3 4 5 6; do SCORE_TRACE_CREATE=1 timeout 200 python scripts/e2e_trace.py 1024 2 4 1 > gpurun_out/e2e_trace_v$r.log 2>&1; tail -1 gpurun_out/e2e_trace_v$r.log; done
timeout 900 python bench.py > gpurun_out/bench_s2v.log 2>&1
tail -4 gpurun_out/pytest_gpu_s2v.log; tail -1 gpurun_out/bench_s2v.log | python -c "
import json,sys
try:
    l=json.loads(sys.stdin.read()); print(round(l['value'],1), 'e2e', l.get('e2e') and round(l['e2e']['value'],1), 'ms', round(l['ms_per_step'],1), 'launches', l['gpu_launches'])
except Exception as e: print('fail', e)
"
