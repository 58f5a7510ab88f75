#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_s2q.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_s2q.log
tail -6 gpurun_out/pytest_gpu_s2q.log
