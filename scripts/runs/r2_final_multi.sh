# usage: bash scripts/runs/r2_final_multi.sh N   (under gpurun --gpus N)
N=$1
timeout 1000 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_r2_final_${N}gpu.json 2> gpurun_out/bench_r2_final_${N}gpu.err
tail -2 gpurun_out/bench_r2_final_${N}gpu.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_r2_final_${N}gpu.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms/step', d['ms_per_step'], 'solved', d['solved'], d['instances'], 'e2e', d['e2e']['value'], 'strong', d['strong_scaling']['value'], d['strong_scaling']['ms_per_step'], 'per_rank', d['per_rank_ms_and_cycles_per_step'])
"
