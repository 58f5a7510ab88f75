"""Per-kernel launch count / summed device time / share from an ncu launch list
(`ncu --metrics gpu__time_duration.sum --clock-control none -c N --csv --log-file launches.csv <command>`).

    python scripts/launch_summary.py gpurun_out/launches_r2.csv > profiles/launches_r2_summary.txt
"""
import csv
import re
import sys
from collections import defaultdict

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows[1:]:
    if r[mi] != "gpu__time_duration.sum":
        continue
    name = r[ki]
    m = re.search(r"(k_[a-z_0-9]+)", name)
    key = m.group(1) if m else ("cub radix sort" if "RadixSort" in name or "radix" in name.lower() else name.split("(")[0][:40])
    v = float(r[vi].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6)
    tot[key] += v
    cnt[key] += 1
allt = sum(tot.values())
print(f"# ncu launch list: {sum(cnt.values())} launches, {allt:.2f} ms of device time in the captured window")
print("# per kernel: launches, summed device time, share of the captured window (cold-cache, serialised: compare SHARES)\n")
for k in sorted(tot, key=lambda k: -tot[k]):
    print(f"{k:28s} {cnt[k]:4d} launches {tot[k]:9.2f} ms  {100 * tot[k] / allt:5.1f} %")
