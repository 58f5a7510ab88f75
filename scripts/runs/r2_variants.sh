timeout 300 python scripts/inst_explore.py 4136 "kkt_tol=1e-7" "kkt_tol=1e-8" 2>&1 | tail -4
timeout 300 python scripts/inst_explore.py 462 "kkt_tol=1e-7" 2>&1 | tail -4
timeout 600 python scripts/unsolved_probe.py 0 10 2>&1 | tail -30
timeout 200 python scripts/kernel_full.py 1024 2>&1 | grep "solve_ms"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
