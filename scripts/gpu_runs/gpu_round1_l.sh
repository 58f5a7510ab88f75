#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/smi_l.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_l.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_l.log
timeout 300 python scripts/large3d.py 20 200 50 40000 30 repeat=2 > gpurun_out/l3d_b_1gpu.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 scripts/large3d.py 20 200 50 40000 30 repeat=2 > gpurun_out/l3d_b_2gpu.log 2>&1
timeout 600 python scripts/large3d.py 100 1000 1000 1000000 100 repeat=2 > gpurun_out/l3d_c_1gpu.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 scripts/large3d.py 100 1000 1000 1000000 100 repeat=2 > gpurun_out/l3d_c_2gpu.log 2>&1
tail -4 gpurun_out/pytest_gpu_l.log; for f in gpurun_out/l3d_*.log; do echo "== $f"; tail -n 3 $f | cut -c1-400; done
