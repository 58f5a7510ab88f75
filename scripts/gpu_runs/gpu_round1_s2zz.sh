#!/bin/bash
# session 3, last call: work-list grid multiplier (CTAs per SM x mult) A/B
mkdir -p gpurun_out
for m in 1 3; do echo "SCORE_WGRID_MULT=$m"; SCORE_WGRID_MULT=$m timeout 100 python scripts/sweep_params.py 1024 "" 2>&1 | grep -v "    inst"; done > gpurun_out/sweep_s2zz.log 2>&1
cat gpurun_out/sweep_s2zz.log | cut -c1-120
