#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_o.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_o.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_o_1gpu.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_o_2gpu.log 2>&1
timeout 600 python scripts/large3d.py 100 1000 1000 1000000 100 repeat=2 > gpurun_out/l3d_c_1gpu.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 scripts/large3d.py 100 1000 1000 1000000 100 repeat=2 > gpurun_out/l3d_c_2gpu.log 2>&1
tail -4 gpurun_out/pytest_gpu_o.log
for f in gpurun_out/bench_o_1gpu.log gpurun_out/bench_o_2gpu.log; do tail -1 $f | python -c "
import json,sys
try:
    l=json.loads(sys.stdin.read()); print('gpus', l['n_gpus'], 'value', round(l['value'],1), 'e2e', round(l['e2e']['value'],1), 'ms', round(l['ms_per_step'],1), 'solved', l['solved'], l['instances'], 'roof', l['roofline']['kernel'], round(l['roofline']['frac'],3))
except Exception as e: print('fail', e)
"; done
for f in gpurun_out/l3d_c_?gpu.log; do tail -n 1 $f | cut -c1-300; done
