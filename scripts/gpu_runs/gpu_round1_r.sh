#!/bin/bash
mkdir -p gpurun_out
FIRST=1024 timeout 600 python scripts/sweep_params.py 1024 "" "mu_factor=0.2" "center_tol=1.0" "cg_forcing=0.05" > gpurun_out/sweep_r1.log 2>&1
FIRST=2048 timeout 600 python scripts/sweep_params.py 1024 "" "mu_factor=0.2" "center_tol=1.0" "cg_forcing=0.05" > gpurun_out/sweep_r2.log 2>&1
cat gpurun_out/sweep_r1.log gpurun_out/sweep_r2.log
