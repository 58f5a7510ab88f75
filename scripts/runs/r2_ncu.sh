# ncu evidence for the bench command (never a bench value): launch list + one full capture of the first solver cycles
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r2.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-per-config > gpurun_out/launches_r2_bench.log 2>&1
wc -l gpurun_out/launches_r2.csv
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:'k_(hessvec|precond_rev|precond_fwd|cg_update|pupdate_vec|rowupdate|linesearch|coarse_build|rows_mf|grad_mf|ctrl_b)' \
  -c 36 -f -o gpurun_out/prof_r2_full python scripts/kernel_full.py 1024 > gpurun_out/prof_r2_full.log 2>&1
ls -la gpurun_out/prof_r2_full.ncu-rep
