# A/B of compile-time variants (variants/*.so built with -DSCORE_*=...) against the in-tree library
for v in default nostage; do
  echo "== $v"
  lib=$PWD/variants/$v.so; [ $v = default ] && lib=$PWD/score_b200/libscore_b200.so
  SCORE_B200_LIB=$lib timeout 200 python scripts/kernel_full.py 1024 2>&1 | grep "solve_ms\|precond\|full PCG"
done
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
