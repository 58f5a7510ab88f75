SCORE_SPLIT_COARSE_APPLY=1 timeout 200 python scripts/kernel_full.py 1024 2>&1 | tail -12
