#!/bin/bash
# session 3: 2-GPU bench under torchrun (weak scaling, no data-path collective) + 2-GPU tests
mkdir -p gpurun_out
nvidia-smi -L | wc -l > gpurun_out/ngpu_s2.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_s2_2gpu.log 2>&1
( timeout 600 python -m pytest tests/test_row_partition.py tests/test_sharding.py -m gpu -x -q ) > gpurun_out/pytest_gpu_s2_2gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu_s2_2gpu.log
tail -1 gpurun_out/bench_s2_2gpu.log | python -c "
import json,sys
try:
    l=json.loads(sys.stdin.read()); print('gpus', l['n_gpus'], 'value', round(l['value'],1), 'e2e', round(l['e2e']['value'],1), 'ms', round(l['ms_per_step'],1), 'solved', l['solved'], l['instances'], l['per_rank_ms_and_cycles_per_step'])
except Exception as e: print('fail', e)
"
tail -n 3 gpurun_out/bench_s2_2gpu.log | cut -c1-300
