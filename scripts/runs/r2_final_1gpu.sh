timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r2_final_1gpu.json 2> gpurun_out/bench_r2_final_1gpu.err; tail -2 gpurun_out/bench_r2_final_1gpu.err
timeout 1200 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_r2_reference_arm.json 2> gpurun_out/bench_r2_reference_arm.err; tail -c 600 gpurun_out/bench_r2_reference_arm.json
timeout 300 python scripts/large3d.py 100 1000 1000 1000000 profile_cycles=100000 > gpurun_out/large3d_r2_final_prof.log 2>&1; tail -16 gpurun_out/large3d_r2_final_prof.log
timeout 200 python scripts/profile_solve.py 1024 gpurun_out/profile_solve_r2_final.json > /dev/null 2>&1
timeout 200 python scripts/kernel_full.py 1024 > gpurun_out/kernel_full_r2_final.txt 2>&1; tail -14 gpurun_out/kernel_full_r2_final.txt
