#!/bin/bash
# session 3, call r: throughput vs instances per GPU (tail amortisation), ncu of the controller kernels
mkdir -p gpurun_out
for n in 2048 4096; do timeout 600 python bench.py --instances $n --steps 2 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/bench_s2r_n$n.log 2>&1; done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_ctrl_a|k_ctrl_b" -s 8 -c 4 -o gpurun_out/prof_s2r_ctrl python scripts/profile_solve.py 1024 x ncu=1 > gpurun_out/ncu_s2r.log 2>&1
for n in 2048 4096; do tail -1 gpurun_out/bench_s2r_n$n.log | python -c "
import json,sys
try:
    l=json.loads(sys.stdin.read()); print(l['config']['instances_per_gpu'], round(l['value'],1), 'ms', round(l['ms_per_step'],1), l['roofline']['kernel'], round(l['roofline']['frac'],3), 'cycles', l['cycles_per_step'])
except Exception as e: print('fail', e)
"; done
