#!/bin/bash
mkdir -p gpurun_out
for f in 0 1024; do
FIRST=$f timeout 600 python scripts/sweep_params.py 1024 "cg_forcing=0.2" "cg_forcing=0.3" "cg_forcing=0.5" "cg_forcing=0.2 cg_per_cycle=3" "cg_forcing=0.2 cg_per_cycle=5" "cg_forcing=0.3 cg_per_cycle=3" "cg_forcing=0.2 mu_factor=0.15" "cg_forcing=0.2 center_tol=6.0" 2>&1 | grep -v "    inst" > gpurun_out/sweep_t$f.log
done
cat gpurun_out/sweep_t*.log
