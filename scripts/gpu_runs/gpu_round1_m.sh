#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_row_partition.py -m gpu -x -q > gpurun_out/pytest_gpu_m.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_m.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 scripts/large3d.py 20 200 50 40000 30 repeat=2 > gpurun_out/l3d_b_2gpu.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 scripts/large3d.py 100 1000 1000 1000000 100 repeat=2 > gpurun_out/l3d_c_2gpu.log 2>&1
tail -15 gpurun_out/pytest_gpu_m.log; for f in gpurun_out/l3d_?_2gpu.log; do echo "== $f"; tail -n 3 $f | cut -c1-400; done
