#!/bin/bash
# session 3, call n: programmatic dependent launch of the tick kernels — parity, A/B bench
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_evaluate.py -m gpu -x -q ) > gpurun_out/pytest_gpu_s2n.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_s2n.log
timeout 300 python scripts/sweep_params.py 1024 "" 2>&1 | grep -v "    inst" > gpurun_out/sweep_s2n_pdl1.log
SCORE_PDL=0 timeout 300 python scripts/sweep_params.py 1024 "" 2>&1 | grep -v "    inst" > gpurun_out/sweep_s2n_pdl0.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_s2n_pdl1.log 2>&1
SCORE_PDL=0 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_s2n_pdl0.log 2>&1
tail -4 gpurun_out/pytest_gpu_s2n.log; cat gpurun_out/sweep_s2n_pdl1.log gpurun_out/sweep_s2n_pdl0.log; for f in gpurun_out/bench_s2n_pdl1.log gpurun_out/bench_s2n_pdl0.log; do tail -1 $f | python -c "
import json,sys
try:
    l=json.loads(sys.stdin.read()); print(round(l['value'],1), 'e2e', l.get('e2e') and round(l['e2e']['value'],1), 'ms', round(l['ms_per_step'],1), l['roofline']['kernel'], round(l['roofline']['frac'],3))
except Exception as e: print('fail', e)
"; done
