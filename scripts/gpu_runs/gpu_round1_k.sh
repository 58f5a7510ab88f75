#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_k.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_k.log
timeout 300 python scripts/large3d.py 10 100 20 10000 20 max_ticks=20000 > gpurun_out/large3d_a.log 2>&1
timeout 300 python scripts/large3d.py 20 200 50 40000 30 max_ticks=20000 > gpurun_out/large3d_b.log 2>&1
timeout 600 python scripts/large3d.py 100 1000 1000 1000000 100 max_ticks=20000 > gpurun_out/large3d_c.log 2>&1
tail -3 gpurun_out/pytest_gpu_k.log; for f in gpurun_out/large3d_?.log; do tail -n 3 $f; done
