"""Experiment driver for the sweep's slow instances: solve chosen sweep instances with different path-following
parameters and print Newton / PCG counts.

    python scripts/tail_exp.py "838,201,173,276,462,963,181,889,551" "mu_factor=0.1" "mu_factor=0.05" ...
Each further argument is one parameter set: comma-separated key=value pairs.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

from score_b200 import build, generators
from score_b200.lowering import concat, lower_manhattan_arrays

ids = [int(a) for a in sys.argv[1].split(",")]
sets = sys.argv[2:] or [""]
probs = [lower_manhattan_arrays(generators.manhattan_2d_arrays(generators.MC_BASE_SEED + i, n_robots=20, n_steps=100), "QCQP",
                                with_names=False) for i in ids]
build.build()
from score_b200.solver import ScoreSolver

with ScoreSolver(concat(probs)) as s:
    for ps in sets:
        kw = {}
        for a in ps.split(","):
            if a:
                k, v = a.split("=")
                kw[k] = float(v) if "." in v or "e" in v else int(v)
        st = s.solve(**kw)
        I = st.instances
        print(f"[{ps}] solved {st.n_solved}/{len(ids)} cycles {st.cycles} solve_ms {st.solve_ms:.1f}")
        print("   newton", I["newton_iters"].tolist(), "sum", int(I["newton_iters"].sum()))
        print("   cg    ", I["cg_iters"].tolist(), "sum", int(I["cg_iters"].sum()))
        print("   kkt   ", [f"{v:.1e}" for v in I["rel_kkt"]], "lsfail", I["ls_failures"].tolist())
