#!/bin/bash
# session 3, call k: create / destroy without device-wide synchronisation — parity, e2e timeline, bench
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_evaluate.py -m gpu -x -q ) > gpurun_out/pytest_gpu_s2k.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_s2k.log
timeout 300 python scripts/e2e_trace.py 1024 2 3 1 > gpurun_out/e2e_trace_k.log 2>&1
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_s2k.log 2>&1
tail -4 gpurun_out/pytest_gpu_s2k.log; tail -8 gpurun_out/e2e_trace_k.log; for f in gpurun_out/bench_s2k.log; do tail -1 $f | python -c "
import json,sys
try:
    l=json.loads(sys.stdin.read()); print(round(l['value'],1), 'e2e', l.get('e2e') and round(l['e2e']['value'],1), 'ms', round(l['ms_per_step'],1), l['roofline']['kernel'], round(l['roofline']['frac'],3))
except Exception as e: print('fail', e)
"; done
