#!/bin/bash
# session 3, call o: the bench lines for profiles/ (default flags, both arms)
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_s2o.log 2>&1
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_s2o.log 2>&1
tail -1 gpurun_out/bench_s2o.log | cut -c1-300; tail -1 gpurun_out/bench_ref_s2o.log | cut -c1-300
