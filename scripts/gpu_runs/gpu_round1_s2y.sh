#!/bin/bash
# session 3, last call: GPU tests + smoke + default bench on the final tree
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_s2y.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_s2y.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke_s2y.log 2>&1
timeout 600 python bench.py > gpurun_out/bench_s2y.log 2>&1
tail -4 gpurun_out/pytest_gpu_s2y.log; tail -1 gpurun_out/smoke_s2y.log; tail -1 gpurun_out/bench_s2y.log | python -c "
import json,sys
try:
    l=json.loads(sys.stdin.read()); print(round(l['value'],1), 'e2e', l.get('e2e') and round(l['e2e']['value'],1), 'ms', round(l['ms_per_step'],1), 'launches', l['gpu_launches'], l['roofline']['kernel'], round(l['roofline']['frac'],3))
except Exception as e: print('fail', e)
"
