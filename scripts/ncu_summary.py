"""Condense an .ncu-rep (ncu --set full) into a small CSV of the metrics the roofline uses, plus
profiles/traffic.json = DRAM bytes (read + write) of one launch per tick kernel.

    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep profiles/ncu_rNN_name.csv [profiles/traffic.json]
"""
import csv
import io
import json
import re
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
traffic_path = sys.argv[3] if len(sys.argv) > 3 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, body = rows[0], rows[1], rows[2:]
want = [
    "ID", "Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "l1tex__t_sector_hit_rate.pct",
    "lts__t_sector_hit_rate.pct", "smsp__cycles_active.avg", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
]
idx = [hdr.index(w) for w in want if w in hdr]


def to_bytes(val, unit):
    v = float(val.replace(",", ""))
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]
    return v * mult


def to_ms(val, unit):
    v = float(val.replace(",", ""))
    return v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1.0)


with open(out, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow([hdr[i] for i in idx] + ["dram_bytes_total", "dram_GBps"])
    w.writerow([units[i] for i in idx] + ["byte", "GB/s"])
    traffic = {}
    ir, iw, it, ik = (hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"),
                      hdr.index("gpu__time_duration.sum"), hdr.index("Kernel Name"))
    for r in body:
        tot = to_bytes(r[ir], units[ir]) + to_bytes(r[iw], units[iw])
        ms = to_ms(r[it], units[it])
        w.writerow([r[i] for i in idx] + [f"{tot:.0f}", f"{tot / ms / 1e6:.1f}"])
        name = re.sub(r"^void ", "", r[ik]).split("<")[0].split("(")[0]
        traffic.setdefault(name, []).append(tot)
if traffic_path:
    json.dump({k: max(v) for k, v in traffic.items()}, open(traffic_path, "w"), indent=1)
print("wrote", out)
