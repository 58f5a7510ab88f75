#!/bin/bash
# session 3, call s: ctrl_a fused into the row pass (last CTA of an instance) — full GPU tests, A/B
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_s2s.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_s2s.log
timeout 200 python scripts/sweep_params.py 1024 "" 2>&1 | grep -v "    inst" > gpurun_out/sweep_s2s_fused.log
SCORE_SPLIT_CTRL_A=1 timeout 200 python scripts/sweep_params.py 1024 "" 2>&1 | grep -v "    inst" > gpurun_out/sweep_s2s_split.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_s2s.log 2>&1
SCORE_SPLIT_CTRL_A=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_s2s_split.log 2>&1
tail -4 gpurun_out/pytest_gpu_s2s.log; cat gpurun_out/sweep_s2s_fused.log gpurun_out/sweep_s2s_split.log | cut -c1-200; for f in gpurun_out/bench_s2s.log gpurun_out/bench_s2s_split.log; do tail -1 $f | python -c "
import json,sys
try:
    l=json.loads(sys.stdin.read()); print(round(l['value'],1), 'e2e', l.get('e2e') and round(l['e2e']['value'],1), 'ms', round(l['ms_per_step'],1), 'launches', l['gpu_launches'])
except Exception as e: print('fail', e)
"; done
