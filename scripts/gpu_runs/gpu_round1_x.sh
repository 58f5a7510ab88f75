#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l > gpurun_out/ngpu_x.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/bench_x_8gpu.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 scripts/large3d.py 100 1000 1000 1000000 100 repeat=2 > gpurun_out/l3d_c_8gpu.log 2>&1
tail -1 gpurun_out/bench_x_8gpu.log | python -c "
import json,sys
try:
    l=json.loads(sys.stdin.read()); print('gpus', l['n_gpus'], 'value', round(l['value'],1), 'e2e', round(l['e2e']['value'],1), 'ms', round(l['ms_per_step'],1), 'solved', l['solved'], l['instances'], l['per_rank_ms_and_cycles_per_step'])
except Exception as e: print('fail', e)
"
tail -n 3 gpurun_out/bench_x_8gpu.log | cut -c1-300
tail -n 2 gpurun_out/l3d_c_8gpu.log | cut -c1-300
