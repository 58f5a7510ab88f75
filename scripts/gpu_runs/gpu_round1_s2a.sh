#!/bin/bash
# session 3, call a: lagged-coarse sweep + source-level ncu capture of the two non-streaming line-search kernels
mkdir -p gpurun_out
timeout 500 python scripts/sweep_params.py 1024 "" "coarse_every=2" "coarse_every=3" "coarse_every=4" 2>&1 | grep -v "    inst" > gpurun_out/sweep_s2a.log
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"k_coarse_build|k_linesearch" -s 2 -c 4 -o gpurun_out/prof_s2a python scripts/profile_solve.py 1024 x ncu=1 > gpurun_out/ncu_s2a.log 2>&1
cat gpurun_out/sweep_s2a.log; tail -3 gpurun_out/ncu_s2a.log; ls -la gpurun_out
