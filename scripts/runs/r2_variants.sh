# in-tree library: per-kernel full-occupancy times + the GPU test suite
timeout 200 python scripts/kernel_full.py 1024 2>&1 | tail -12
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
