#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scripts/large3d.py 100 1000 1000 1000000 100 repeat=2 profile_cycles=100000 > gpurun_out/l3d_c_prof.log 2>&1
tail -n 16 gpurun_out/l3d_c_prof.log | cut -c1-300
