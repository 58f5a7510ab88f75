#!/bin/bash
# session 3, call t: run-to-run spread of the streamed end-to-end path vs the number of table-building threads
mkdir -p gpurun_out
nproc > gpurun_out/nproc_s2t.txt
for nt in 16 8 4; do for r in 1 2 3; do SCORE_CREATE_THREADS=$nt timeout 200 python scripts/e2e_trace.py 1024 2 4 1 2>&1 | tail -1 | sed "s/^/threads=$nt run=$r: /"; done; done > gpurun_out/e2e_spread_s2t.log 2>&1
cat gpurun_out/nproc_s2t.txt gpurun_out/e2e_spread_s2t.log
