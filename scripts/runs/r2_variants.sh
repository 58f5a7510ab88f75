for v in default spin0 spinall; do
  echo "== $v"
  lib=$PWD/variants/$v.so; [ $v = default ] && lib=$PWD/score_b200/libscore_b200.so
  SCORE_B200_LIB=$lib timeout 200 python scripts/per_config.py 2>&1 | tail -3
done
for v in default spin0; do
  echo "== bench $v"
  lib=$PWD/variants/$v.so; [ $v = default ] && lib=$PWD/score_b200/libscore_b200.so
  SCORE_B200_LIB=$lib timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-per-config > gpurun_out/bench_var_$v.json 2> gpurun_out/bench_var_$v.err
  python -c "
import json
d=json.loads(open('gpurun_out/bench_var_$v.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'])
"
done
