#!/bin/bash
mkdir -p gpurun_out
for f in 7168 4096 0 2048; do
FIRST=$f timeout 600 python scripts/sweep_params.py 1024 "" 2>&1 | head -4 > gpurun_out/sweep_z$f.log
done
cat gpurun_out/sweep_z*.log
