#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/e2e_trace.py 1024 2 3 1 > gpurun_out/e2e_trace_a.log 2>&1
timeout 300 python scripts/e2e_trace.py 1024 2 3 0 > gpurun_out/e2e_trace_b.log 2>&1
tail -8 gpurun_out/e2e_trace_a.log; tail -8 gpurun_out/e2e_trace_b.log
