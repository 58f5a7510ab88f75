for v in pf16 pf12 pf10 pf8; do
  echo "== $v"
  SCORE_B200_LIB=$PWD/variants/$v.so timeout 200 python scripts/kernel_full.py 1024 2>&1 | grep "solve_ms\|precond\|full PCG"
done
