"""Per-launch time and achieved GB/s of every tick kernel at full occupancy (launches whose work list is the whole
batch: cycles 2-5, line-search tick for the line-search-only kernels, first PCG tick for the others).

    python scripts/kernel_full.py [n_instances] [key=value solver params ...]
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
kw = {}
for a in sys.argv[2:]:
    k, v = a.split("=")
    kw[k] = float(v) if "." in v or "e" in v else int(v)
prob = bench.make_batch(0, n, 20, 100)
from score_b200 import build
build.build()
from score_b200.solver import KERNEL_NAMES, ScoreSolver
with ScoreSolver(prob) as s:
    s.solve(**kw)
    st = s.solve(**kw)
    stf = s.solve(profile_cycles=4, profile_skip=2, **kw)
print(f"solve_ms {st.solve_ms:.1f} cycles {st.cycles} solved {st.n_solved}")
tot = 0.0
for k, ms, c, b in zip(KERNEL_NAMES, stf.kernel_ms_full, stf.kernel_count_full, stf.kernel_bytes):
    if c > 0:
        per = ms / c
        tot += per if k not in ("k_linesearch", "k_rowupdate", "k_coarse_build") else 0
        print(f"  {k:16s} {1e3 * per:8.1f} us/launch  {b / 1e9:6.3f} GB  {b / per / 1e6 if per > 0 else 0:8.0f} GB/s  ({int(c)} launches)")
print(f"  one full PCG tick: {1e3 * tot:.0f} us")
