# ncu evidence for the bench command (never a bench value): launch list + one full capture of a full-occupancy cycle.
# The .ncu-rep stays on the box (gpurun copies back at most 64 MiB): it is condensed there.
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r2.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-per-config > gpurun_out/launches_r2_bench.log 2>&1
wc -l gpurun_out/launches_r2.csv
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:'k_(hessvec|precond_rev|precond_fwd|cg_update|pupdate_vec|rowupdate|linesearch|coarse_build|rows_mf|grad_mf|ctrl_b)' \
  -s 28 -c 30 -f -o /tmp/prof_r2_full python scripts/kernel_full.py 1024 > gpurun_out/prof_r2_full.log 2>&1
ls -la /tmp/prof_r2_full.ncu-rep
python scripts/ncu_summary.py /tmp/prof_r2_full.ncu-rep gpurun_out/ncu_r2_full.csv gpurun_out/traffic.json
ncu -i /tmp/prof_r2_full.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv, sys
rows = list(csv.reader(sys.stdin)); h = rows[0]
want = ['Kernel Name', 'gpu__time_duration.sum', 'l1tex__throughput.avg.pct_of_peak_sustained_active', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active']
idx = [h.index(w) for w in want if w in h]
print(','.join(h[i] for i in idx))
for r in rows[2:]: print(','.join(r[i].split('(')[0][:40] for i in idx))
" > gpurun_out/ncu_r2_l1.csv
du -sh gpurun_out
