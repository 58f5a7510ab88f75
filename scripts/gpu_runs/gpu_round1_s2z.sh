#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "stream_group" > gpurun_out/pytest_gpu_s2z.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_s2z.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_s2z.log 2>&1
tail -3 gpurun_out/pytest_gpu_s2z.log; tail -1 gpurun_out/bench_s2z.log | python -c "
import json,sys
try:
    l=json.loads(sys.stdin.read()); print(round(l['value'],1), 'e2e', l.get('e2e') and round(l['e2e']['value'],1), 'ms', round(l['ms_per_step'],1))
except Exception as e: print('fail', e)
"
