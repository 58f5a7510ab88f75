#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/e2e_stages.py 512 3 > gpurun_out/e2e_stages_512.log 2>&1
timeout 300 python scripts/e2e_stages.py 1024 3 > gpurun_out/e2e_stages_1024.log 2>&1
tail -3 gpurun_out/e2e_stages_512.log; tail -3 gpurun_out/e2e_stages_1024.log
