"""Solver-parameter sweep on the Monte-Carlo batch: one line per setting.

    python scripts/sweep_params.py N "cg_per_cycle=3" "cg_per_cycle=4 mu_factor=0.05" ...
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

import bench

n = int(sys.argv[1])
first = int(os.environ.get("FIRST", "0"))
prob = bench.make_batch(first, n, 20, 100)
from score_b200 import build

build.build()
from score_b200.solver import ScoreSolver

with ScoreSolver(prob) as s:
    s.solve()
    for spec in sys.argv[2:] or [""]:
        kw = {}
        for a in spec.split():
            k, v = a.split("=")
            kw[k] = float(v) if "." in v or "e" in v else int(v)
        st = s.solve(**kw)
        st = s.solve(**kw)
        I = st.instances
        print(f"[{spec}] solved {st.n_solved}/{n} solve_ms {st.solve_ms:.1f} cycles {st.cycles} ticks {st.ticks} "
              f"newton p50/p99/max {np.percentile(I['newton_iters'], 50):.0f}/{np.percentile(I['newton_iters'], 99):.0f}/{I['newton_iters'].max()} "
              f"cg mean {I['cg_iters'].mean():.1f} max {I['cg_iters'].max()} lsfail {int((I['ls_failures'] > 0).sum())} "
              f"maxkkt {I['rel_kkt'].max():.2e}", flush=True)
        worst = np.argsort(-(I["newton_iters"] + 1000 * (1 - I["solved"])))[:6]
        for w in worst:
            print("    inst", first + int(w), {k: (float(I[w][k]) if I[w][k].dtype.kind == "f" else int(I[w][k])) for k in I.dtype.names}, flush=True)
