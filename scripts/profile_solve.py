"""Whole-solve kernel-time breakdown of the Monte-Carlo sweep batch (CUDA events between every kernel
of every tick, un-graphed) plus the per-instance iteration statistics.

    python scripts/profile_solve.py [n_instances] [out.json] [key=value solver params ...]

With `ncu=1` it only runs one plain (graphed) solve — the command to wrap in ncu.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
out = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out", "profile_solve.json")
kw = {}
for a in sys.argv[3:]:
    k, v = a.split("=")
    kw[k] = float(v) if "." in v or "e" in v else int(v)
ncu = kw.pop("ncu", 0)
prob = bench.make_batch(0, n, 20, 100)
from score_b200 import build

build.build()
from score_b200.solver import KERNEL_NAMES, ScoreSolver

with ScoreSolver(prob) as s:
    if ncu:
        st = s.solve(**kw)
        print("ncu solve: solved", st.n_solved, "ticks", st.ticks)
        sys.exit(0)
    st = s.solve(**kw)  # warm
    st = s.solve(**kw)
    stp = s.solve(profile_cycles=int(st.cycles), profile_skip=0, **kw)
I = st.instances
tot = I["cg_iters"] + I["newton_iters"]
rep = {
    "n": n,
    "solved": int(st.n_solved),
    "ticks": int(st.ticks),
    "cycles": int(st.cycles),
    "solve_ms": st.solve_ms,
    "assemble_ms": st.assemble_ms,
    "setup_ms": st.setup_ms,
    "extract_ms": st.extract_ms,
    "whole_solve_gbs": st.algorithmic_bytes / st.solve_ms / 1e6,
    "ticks_per_instance_pct": dict(zip(["p0", "p10", "p50", "p90", "p99", "p100"],
                                       np.percentile(tot, [0, 10, 50, 90, 99, 100]).tolist())),
    "newton_pct": np.percentile(I["newton_iters"], [0, 10, 50, 90, 99, 100]).tolist(),
    "cg_pct": np.percentile(I["cg_iters"], [0, 10, 50, 90, 99, 100]).tolist(),
    "mean_active_fraction": float(tot.mean() / st.ticks),
    "profiled_solve_ms": stp.solve_ms,
    "kernel_ms_total": {k: float(v) for k, v in zip(KERNEL_NAMES, stp.kernel_ms)},
    "kernel_share": {k: float(v / stp.kernel_ms.sum()) for k, v in zip(KERNEL_NAMES, stp.kernel_ms)},
    "kernel_launches": {k: int(v) for k, v in zip(KERNEL_NAMES, stp.kernel_count)},
    "kernel_gbs_whole_solve": {k: (float(b / (v * 1e6)) if v > 0 else None)
                               for k, v, b in zip(KERNEL_NAMES, stp.kernel_ms, stp.kernel_bytes_total)},
}
print(json.dumps(rep, indent=1))
with open(out, "w") as f:
    json.dump(rep, f, indent=1)
