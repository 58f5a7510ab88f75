#!/bin/bash
# session 3, call m: arena / resource caches — full GPU tests, e2e timeline with create phases, bench
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_s2m.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_s2m.log
SCORE_TRACE_CREATE=1 timeout 300 python scripts/e2e_trace.py 1024 2 4 1 > gpurun_out/e2e_trace_m.log 2>&1
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_s2m.log 2>&1
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_s2m_k6.log 2>&1
tail -4 gpurun_out/pytest_gpu_s2m.log; grep score_create gpurun_out/e2e_trace_m.log | tail -8 | cut -c1-200; tail -9 gpurun_out/e2e_trace_m.log; for f in gpurun_out/bench_s2m.log gpurun_out/bench_s2m_k6.log; do tail -1 $f | python -c "
import json,sys
try:
    l=json.loads(sys.stdin.read()); print(round(l['value'],1), 'e2e', l.get('e2e') and round(l['e2e']['value'],1), 'ms', round(l['ms_per_step'],1), l['roofline']['kernel'], round(l['roofline']['frac'],3))
except Exception as e: print('fail', e)
"; done
