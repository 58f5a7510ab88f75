#!/bin/bash
mkdir -p gpurun_out
for f in 4096 7168 3072 5120 6144; do
FIRST=$f timeout 600 python scripts/sweep_params.py 1024 "" 2>&1 | head -5 > gpurun_out/sweep_y$f.log
done
cat gpurun_out/sweep_y*.log
