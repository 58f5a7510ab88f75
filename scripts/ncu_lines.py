"""Per-source-line stall-sample totals from `ncu --page source --print-source cuda,sass --csv` output.

    ncu -i rep.ncu-rep --page source --print-source cuda,sass --csv --kernel-name regex:NAME --launch-count 1 > src.csv
    python scripts/ncu_lines.py src.csv [min_pct]
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
minpct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
cur_file, hdr = None, None
agg = {}
for r in rows:
    if len(r) >= 2 and r[0] in ("File Path", "File Name"):
        cur_file = r[1].split("/")[-1]
        continue
    if len(r) > 6 and r[0] == "Line No":
        hdr = r
        si = hdr.index("# Samples")
        continue
    if hdr is None or len(r) <= si:
        continue
    if r[0]:  # a CUDA source line: following sass rows belong to it
        cur = (cur_file, int(r[0]), r[1])
        agg.setdefault(cur, [0, 0])
        continue
    if r[si].isdigit():
        agg[cur][0] += int(r[si])
        ie = hdr.index("Instructions Executed")
        if r[ie].isdigit():
            agg[cur][1] += int(r[ie])
tot = sum(v[0] for v in agg.values())
print("total samples", tot)
for (f, ln, src), (s, ie) in sorted(agg.items(), key=lambda kv: (kv[0][0], kv[0][1])):
    if s >= tot * minpct / 100:
        print(f"{100 * s / tot:5.1f}% {ie:>10d}  {f}:{ln:<4d} {src.strip()[:110]}")
