#!/bin/bash
mkdir -p gpurun_out
for r in 1 2 3 4 5; do SCORE_TRACE_CREATE=1 timeout 200 python scripts/e2e_trace.py 1024 2 4 1 > gpurun_out/e2e_trace_u$r.log 2>&1; tail -1 gpurun_out/e2e_trace_u$r.log; done
