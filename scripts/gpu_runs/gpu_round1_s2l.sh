#!/bin/bash
mkdir -p gpurun_out
SCORE_TRACE_CREATE=1 timeout 300 python scripts/e2e_trace.py 1024 2 3 1 > gpurun_out/e2e_trace_l.log 2>&1
grep score_create gpurun_out/e2e_trace_l.log | tail -9; tail -7 gpurun_out/e2e_trace_l.log
