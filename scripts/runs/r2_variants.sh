timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 200 python scripts/kernel_full.py 1024 2>&1 | grep "solve_ms"
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-per-config --no-e2e --no-refine > gpurun_out/bench_var.json 2> gpurun_out/bench_var.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_var.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms/step', d['ms_per_step'], 'solved', d['solved'], d['instances'], 'colpass share', d['roofline']['kernel_share_of_step']['k_colpass'], 'whole-step gbs', d['roofline']['kernel_gbs_whole_step']['k_colpass'])
"
