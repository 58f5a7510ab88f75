#!/bin/bash
mkdir -p gpurun_out
for f in 0 1024 2048 3072; do
FIRST=$f timeout 600 python scripts/sweep_params.py 1024 "" "cg_forcing=0.05" "cg_forcing=0.2" 2>&1 | grep -v "    inst" > gpurun_out/sweep_s$f.log
done
cat gpurun_out/sweep_s*.log
